// emu_harness.cpp -- TEST INFRASTRUCTURE.  Sequential CPU replay of the CUDA kernels' tile loops,
// built from the same rod_core.h / rod_tables.h the kernels use, so the index math, the resize
// tables and the per-chunk arithmetic can be checked against the oracle in a container that
// has no GPU.  Compiled by tests/test_emulation.py with g++; never linked into the product.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../robust-object-detection_b200/csrc/rod_core.h"
#include "../../robust-object-detection_b200/csrc/rod_tables.h"

using namespace rod;

static const int kBlurLeft = 64;
static const int kLowresTH = 32, kLowresTWB = 480;

// Replays blur_rows_kernel<9> / <0> for one image.  `dst_phase` shifts the destination (and
// source) start address phase inside a 16-byte block, as an unaligned device buffer would.
extern "C" int emu_blur(const uint8_t* src, uint8_t* dst, int h, int w, long src_pitch, long dst_pitch, int k,
                        int dst_phase) {
    const int n = 3 * w;
    const int halo = 3 * (k >> 1);
    std::vector<uint8_t> bufv(kBlurLeft + ((15 + n + 15) & ~15) + 64 + 16, 0xAB);
    // 16-byte aligned scratch
    uint8_t* buf = bufv.data() + ((16 - ((uintptr_t)bufv.data() & 15)) & 15);
    for (int y = 0; y < h; ++y) {
        const uint8_t* srow = src + (long)y * src_pitch;
        uint8_t* drow = dst + (long)y * dst_pitch;
        const int shift = (int)(((long)y * dst_pitch + dst_phase) & 15);  // emulated address phase
        uint8_t* row = buf + kBlurLeft + shift;
        const int nchunks = (shift + n + 15) >> 4;
        memset(buf, 0xCD, kBlurLeft + ((15 + n + 15) & ~15) + 64);  // garbage, like uninitialised smem
        for (int j = 0; j < nchunks; ++j) {
            const int lo = 16 * j - shift;
            for (int b = 0; b < 16; ++b) {
                const int i = lo + b;
                if (i >= 0 && i < n) row[i] = srow[i];
            }
        }
        for (int q = 0; q < 2 * halo; ++q) {
            const int i = (q < halo) ? (q - halo) : (n + q - halo);
            const int px = (i >= 0) ? i / 3 : -((-i + 2) / 3);
            const int c = i - 3 * px;
            row[i] = row[3 * reflect101(px, w) + c];
        }
        for (int j = 0; j < nchunks; ++j) {
            const int lo = 16 * j - shift;
            uint32_t out[4];
            if (k == 9) {
                uint32_t wv[12];
                memcpy(wv, buf + kBlurLeft + 16 * j - 16, 48);
                blur9_chunk16(wv, out);
            } else {
                for (int g = 0; g < 4; ++g) {
                    uint32_t o = 0;
                    for (int b = 0; b < 4; ++b) {
                        const int i = lo + 4 * g + b;
                        const uint32_t v = (i >= 0 && i < n) ? blur_byte_generic(row, i, k) : 0u;
                        o |= v << (8 * b);
                    }
                    out[g] = o;
                }
            }
            for (int b = 0; b < 16; ++b) {
                const int i = lo + b;
                if (i >= 0 && i < n) drow[i] = (uint8_t)(out[b >> 2] >> (8 * (b & 3)));
            }
        }
    }
    return 0;
}

// Replays lowres_kernel for one image: same tiling (kLowresTH rows x kLowresTWB byte columns), same
// phase B variants (vector exact-2x units / generic), C1 (closed form / table) and C2 (marching).
// `src_phase`: emulated address phase of the source (the vector path needs 4-byte alignment).
static int emu_lowres_generic(const uint8_t* src, uint8_t* dst, int h, int w, long src_pitch, long dst_pitch,
                              double factor, int src_phase) {
    (void)src_phase;
    std::vector<uint32_t> blob;
    DevShape sh;
    if (!build_lowres_shape(h, w, factor, 8, blob, &sh)) return 2;
    if (blob.empty()) blob.push_back(0);
    const uint32_t* tab = blob.data();
    if (sh.lin_identity) {
        for (int y = 0; y < h; ++y) memcpy(dst + (long)y * dst_pitch, src + (long)y * src_pitch, 3 * w);
        return 0;
    }
    const int32_t* lx_s0 = (const int32_t*)(tab + sh.lx_s0);
    const uint32_t* lx_a = tab + sh.lx_a;
    const uint32_t* ly_s = tab + sh.ly_s;
    const uint32_t* ly_b = tab + sh.ly_b;
    const int32_t* xfirst = (const int32_t*)(tab + sh.ax_first);
    const int32_t* xcount = (const int32_t*)(tab + sh.ax_count);
    const float* xalpha = (const float*)(tab + sh.ax_alpha);
    const int32_t* yfirst = (const int32_t*)(tab + sh.ay_first);
    const int32_t* ycount = (const int32_t*)(tab + sh.ay_count);
    const float* yalpha = (const float*)(tab + sh.ay_alpha);
    int half_rows, half_cols, src_rows;
    lowres_tile_footprint(sh, tab, kLowresTH, kLowresTWB, &half_rows, &half_cols, &src_rows);
    const int p_pitch = (3 * (half_cols + 3) + 15) & ~15;
    const int hx_pitch = kLowresTWB + 8;
    const int hb_pitch = 3 * half_cols + 1;
    std::vector<uint8_t> P((size_t)half_rows * p_pitch);
    std::vector<uint16_t> hx((size_t)half_rows * hx_pitch);
    std::vector<float> hbuf((size_t)src_rows * hb_pitch);
    const int n = 3 * w;
    const bool general = sh.area_mode == AREA_GENERAL;
    for (int y0 = 0; y0 < h; y0 += kLowresTH)
        for (int b0 = 0; b0 < n; b0 += kLowresTWB) {
            const int th = std::min(kLowresTH, h - y0), twb = std::min(kLowresTWB, n - b0);
            const int x_first = b0 / 3, x_last = (b0 + twb - 1) / 3;
            const int j_lo = (int)(ly_s[y0] & 0xFFFFu), j_hi = (int)(ly_s[y0 + th - 1] >> 16);
            const int nj = j_hi - j_lo + 1;
            if (nj > half_rows) return 3;
            const int i_lo = lx_s0[x_first], i_hi = std::min(lx_s0[x_last] + 1, sh.nw - 1);
            const int i_base = i_lo;
            if (i_hi - i_lo + 1 > half_cols) return 4;
            const int ncol = 3 * (i_hi - i_lo + 1);
            const int sy_lo = general ? yfirst[j_lo] : j_lo * sh.iy;
            const int sy_hi = general ? yfirst[j_hi] + ycount[j_hi] - 1 : j_hi * sh.iy + sh.iy - 1;
            const int nsr = sy_hi - sy_lo + 1;
            if (nsr > src_rows || sy_hi >= h) return 5;
            std::fill(P.begin(), P.end(), 0xEE);
            std::fill(hx.begin(), hx.end(), 0xEEEE);
            std::fill(hbuf.begin(), hbuf.end(), -1e30f);
            for (int o = 0; o < ncol; ++o) {  // B1
                const int ir = o / 3, cch = o - 3 * ir, dx = i_lo + ir;
                const int sx0 = general ? xfirst[dx] : dx * sh.ix;
                const int nx = general ? xcount[dx] : sh.ix;
                const uint8_t* sp = src + (long)sy_lo * src_pitch + 3 * sx0 + cch;
                for (int sr = 0; sr < nsr; ++sr, sp += src_pitch) {
                    float buf;
                    if (general && nx <= 4) {
                        float al[4];
                        for (int q = 0; q < 4; ++q) al[q] = (q < nx) ? xalpha[dx * sh.xt + q] : 0.f;
                        const int o1 = 3 * std::min(1, nx - 1), o2 = 3 * std::min(2, nx - 1), o3 = 3 * std::min(3, nx - 1);
                        buf = fmul((float)sp[0], al[0]);
                        buf = fadd(buf, fmul((float)sp[o1], al[1]));
                        buf = fadd(buf, fmul((float)sp[o2], al[2]));
                        buf = fadd(buf, fmul((float)sp[o3], al[3]));
                    } else if (general) {
                        buf = 0.f;
                        for (int q = 0; q < nx; ++q) buf = fadd(buf, fmul((float)sp[3 * q], xalpha[dx * sh.xt + q]));
                    } else {
                        uint32_t sum = 0;
                        for (int q = 0; q < nx; ++q) sum += sp[3 * q];
                        buf = (float)sum;
                    }
                    hbuf[(size_t)sr * hb_pitch + o] = buf;
                }
            }
            for (int o = 0; o < ncol; ++o)  // B2
                for (int jr = 0; jr < nj; ++jr) {
                    const int j = j_lo + jr;
                    uint32_t v;
                    if (general) {
                        const float* be = yalpha + j * sh.yt;
                        const float* hp = hbuf.data() + (size_t)(yfirst[j] - sy_lo) * hb_pitch + o;
                        float sum = fmul(be[0], hp[0]);
                        for (int q = 1; q < ycount[j]; ++q) sum = fadd(sum, fmul(be[q], hp[(size_t)q * hb_pitch]));
                        float r = frint(sum);
                        r = r < 0.f ? 0.f : (r > 255.f ? 255.f : r);
                        v = (uint32_t)(int)r;
                    } else {
                        const float* hp = hbuf.data() + (size_t)(j * sh.iy - sy_lo) * hb_pitch + o;
                        float sum = 0.f;
                        for (int q = 0; q < sh.iy; ++q) sum += hp[(size_t)q * hb_pitch];
                        if (sh.area_mode == AREA_FAST2) {
                            v = ((uint32_t)sum + 2u) >> 2;
                        } else {
                            float r = frint(fmul(sum, sh.inv_area));
                            r = r < 0.f ? 0.f : (r > 255.f ? 255.f : r);
                            v = (uint32_t)(int)r;
                        }
                    }
                    P[jr * p_pitch + o] = (uint8_t)v;
                }
            for (int jr = 0; jr < nj; ++jr)  // C1
                for (int ob = 0; ob < twb; ++ob) {
                    const int o = b0 + ob;
                    const int x = o / 3, c = o - 3 * x;
                    const int s0 = lx_s0[x];
                    const int s1 = std::min(s0 + 1, sh.nw - 1);
                    const uint8_t* prow = P.data() + jr * p_pitch;
                    hx[jr * hx_pitch + ob] = (uint16_t)linear_h4(prow[(s0 - i_base) * 3 + c], prow[(s1 - i_base) * 3 + c], lx_a[x]);
                }
            for (int tid = 0; tid < 256; ++tid) {
                const int rg = tid >> 6, cc = tid & 63;
                const int col = 8 * cc;
                const int nvalid = std::min(8, twb - col);
                if (nvalid <= 0) continue;
                uint32_t ha[8], hb[8];
                int ja = -1, jb = -1;
                const int r_end = std::min(th, 8 * rg + 8);
                for (int r = 8 * rg; r < r_end; ++r) {
                    const uint32_t ys = ly_s[y0 + r], yb = ly_b[y0 + r];
                    const int s0 = (int)(ys & 0xFFFFu), s1 = (int)(ys >> 16);
                    if (s0 != ja) {
                        if (s0 == jb) { for (int q = 0; q < 8; ++q) ha[q] = hb[q]; }
                        else for (int q = 0; q < 8; ++q) ha[q] = hx[(s0 - j_lo) * hx_pitch + col + q];
                        ja = s0;
                    }
                    if (s1 != jb) {
                        if (s1 == ja) { for (int q = 0; q < 8; ++q) hb[q] = ha[q]; }
                        else for (int q = 0; q < 8; ++q) hb[q] = hx[(s1 - j_lo) * hx_pitch + col + q];
                        jb = s1;
                    }
                    const uint32_t c0 = yb & 0xFFFFu, c1 = yb >> 16;
                    for (int q = 0; q < nvalid; ++q)
                        dst[(long)(y0 + r) * dst_pitch + b0 + col + q] =
                            (uint8_t)((((c0 * ha[q]) >> 16) + ((c1 * hb[q]) >> 16) + 2u) >> 2);
                }
            }
        }
    return 0;
}

// Replays lowres_x2_kernel (full-width strips of exact-2x shapes).
static int emu_lowres_x2(const uint8_t* src, uint8_t* dst, int h, int w, long src_pitch, long dst_pitch,
                         const DevShape& sh, const uint32_t* tab, int src_phase) {
    const int n = 3 * w, nw = sh.nw;
    const int p_pitch = (3 * nw + 24 + 15) & ~15;
    const uint32_t* ly_s = tab + sh.ly_s;
    const uint32_t* ly_b = tab + sh.ly_b;
    std::vector<uint32_t> smem_words((size_t)sh.strip_half_rows * p_pitch / 4 + 8);
    uint8_t* smem = (uint8_t*)smem_words.data();
    for (int y0 = 0; y0 < h; y0 += sh.strip_rows) {
        const int th = std::min(sh.strip_rows, h - y0);
        const int j_lo = (int)(ly_s[y0] & 0xFFFFu), j_hi = (int)(ly_s[y0 + th - 1] >> 16);
        const int nj = j_hi - j_lo + 1;
        if (nj > sh.strip_half_rows) return 6;
        memset(smem, 0xEE, (size_t)sh.strip_half_rows * p_pitch);
        const bool vec = (w & 3) == 0 && (((long)src_phase | src_pitch) & 3) == 0 &&
                         (sh.area_mode == AREA_FAST2 || (sh.area_mode == AREA_GENERAL && sh.ay_packed));
        if (vec) {
            const int n_units = nw >> 1;
            const uint32_t magic_div = 0xFFFFFFFFu / (uint32_t)n_units + 1u;
            const int total = nj * n_units;
            for (int idx = 0; idx < total; ++idx) {
                const int jr = (n_units == 1) ? idx : (int)mulhi32((uint32_t)idx, magic_div);
                const int u = idx - jr * n_units;
                if (jr != idx / n_units) return 7;
                const int dy = j_lo + jr, sb = 12 * u;
                uint32_t o6[6];
                if (sh.area_mode == AREA_FAST2) {
                    uint32_t ra[3], rb[3];
                    memcpy(ra, src + (long)(2 * dy) * src_pitch + sb, 12);
                    memcpy(rb, src + (long)(2 * dy + 1) * src_pitch + sb, 12);
                    area_fast2_unit(ra, rb, o6);
                } else {
                    const uint32_t* pk = tab + sh.ay_pack + 4 * dy;
                    const int sy0 = (int)pk[0];
                    float acc[6];
                    for (int ty = 0; ty < 3; ++ty) {
                        uint32_t rw[3];
                        const int row = (ty < 2) ? sy0 + ty : std::min(sy0 + 2, h - 1);
                        memcpy(rw, src + (long)row * src_pitch + sb, 12);
                        area_x2f_accumulate(rw, bitsf(pk[1 + ty]), ty == 0, acc);
                    }
                    area_x2f_finish(acc, o6);
                }
                uint8_t* prow = smem + jr * p_pitch;
                for (int q = 0; q < 6; ++q) prow[4 + 6 * u + q] = (uint8_t)o6[q];
            }
        } else {
            const int total = nj * 3 * nw;
            for (int idx = 0; idx < total; ++idx) {
                const int jr = idx / (3 * nw), o = idx - jr * 3 * nw;
                const int i = o / 3, c = o - 3 * i;
                smem[jr * p_pitch + 4 + o] = (uint8_t)area_value(src, src_pitch, sh, tab, j_lo + jr, i, c);
            }
        }
        for (int q = 0; q < nj * 6; ++q) {
            const int jr = q / 6, kk = q - 6 * jr;
            uint8_t* prow = smem + jr * p_pitch;
            if (kk < 3) prow[1 + kk] = prow[4 + kk];
            else prow[4 + 3 * nw + (kk - 3)] = prow[4 + 3 * (nw - 1) + (kk - 3)];
        }
        const int nchunks = (w + 7) >> 3, ngroups = (th + 7) >> 3;
        const uint32_t magic_div = 0xFFFFFFFFu / (uint32_t)nchunks + 1u;
        for (int idx = 0; idx < ngroups * nchunks; ++idx) {
            const int rg = (nchunks == 1) ? idx : (int)mulhi32((uint32_t)idx, magic_div);
            const int ch = idx - rg * nchunks;
            if (rg != idx / nchunks) return 8;
            const int nvalid = std::min(24, n - 24 * ch);
            float xe[24], xo[24];
            int je = -1, jo = -1;
            auto load_row = [&](int j, float* x) {
                const uint32_t* wq = (const uint32_t*)(smem + (j - j_lo) * p_pitch) + 3 * ch;
                uint32_t win[5];
                for (int k = 0; k < 5; ++k) win[k] = funnel_r(wq[k], wq[k + 1], 8);
                x2_expand24(win, x);
            };
            const int r_end = std::min(th, 8 * rg + 8);
            for (int r = 8 * rg; r < r_end; ++r) {
                const uint32_t ys = ly_s[y0 + r];
                const int s0 = (int)(ys & 0xFFFFu), s1 = (int)(ys >> 16);
                const X2Row rc = x2_row_consts(ly_b[y0 + r]);
                if ((s0 & 1) ? (jo != s0) : (je != s0)) { if (s0 & 1) { load_row(s0, xo); jo = s0; } else { load_row(s0, xe); je = s0; } }
                if ((s1 & 1) ? (jo != s1) : (je != s1)) { if (s1 & 1) { load_row(s1, xo); jo = s1; } else { load_row(s1, xe); je = s1; } }
                const float* xlo = (s0 & 1) ? xo : xe;
                const float* xhi = (s1 & 1) ? xo : xe;
                for (int t = 0; t < nvalid; ++t)
                    dst[(long)(y0 + r) * dst_pitch + 24 * ch + t] = (uint8_t)(x2_vertical(xlo[t], xhi[t], rc) & 0xFFu);
            }
        }
    }
    return 0;
}

// Replays lowres_x2w_kernel (warp-marching: 32 lanes x (band, strip) tiles, neighbour pixels by shuffle).
extern "C" int emu_lowres_x2w(const uint8_t* src, uint8_t* dst, int h, int w, long src_pitch, long dst_pitch,
                              double factor, int band_rows) {
    std::vector<uint32_t> blob;
    DevShape sh;
    if (!build_lowres_shape(h, w, factor, 8, blob, &sh)) return 2;
    if (blob.empty()) blob.push_back(0);
    if (!sh.x2w) return 3;
    const uint32_t* tab = blob.data();
    const int n = 3 * w, nw = sh.nw;
    const bool fast2 = sh.area_mode == AREA_FAST2;
    const uint32_t* ypack = tab + sh.ay_pack;
    const uint32_t* ly_s = tab + sh.ly_s;
    const float* ly_rc = (const float*)(tab + sh.ly_rc);
    const int nchunks = (w + 7) >> 3, nstrips = (nchunks + 29) / 30;
    for (int Y0 = 0; Y0 < h; Y0 += band_rows)
        for (int st = 0; st < nstrips; ++st) {
            const int Y1 = std::min(h, Y0 + band_rows);
            const int j_first = (int)(ly_s[Y0] & 0xFFFFu);
            float xe[32][24], xo[32][24];
            for (int l = 0; l < 32; ++l) for (int q = 0; q < 24; ++q) xe[l][q] = xo[l][q] = -1e30f;  // poison
            int have = j_first - 1;
            for (int r = Y0; r < Y1; ++r) {
                const uint32_t ys = ly_s[r];
                const int s0 = (int)(ys & 0xFFFFu), s1 = (int)(ys >> 16);
                while (have < s1) {
                    ++have;
                    const int j = have;
                    uint32_t own[32][3];
                    for (int l = 0; l < 32; ++l) {
                        const int ch = 30 * st - 1 + l;
                        const int cc = std::min(std::max(ch, 0), nchunks - 1);
                        const bool second = (nw - 4 * cc) >= 4;
                        const int sy0 = fast2 ? 2 * j : (int)ypack[4 * j];
                        const int rows[3] = {sy0, sy0 + 1, std::min(sy0 + 2, h - 1)};
                        uint32_t rw[3][6];
                        for (int t = 0; t < 3; ++t) {
                            const uint8_t* rp = src + (long)rows[t] * src_pitch + 24 * cc;
                            memcpy(&rw[t][0], rp, 12);
                            memcpy(&rw[t][3], rp + (second ? 12 : 0), 12);
                        }
                        uint32_t o6[2][6];
                        for (int hf = 0; hf < 2; ++hf) {
                            if (fast2) {
                                area_fast2_unit(&rw[0][3 * hf], &rw[1][3 * hf], o6[hf]);
                            } else {
                                float acc[6];
                                for (int t = 0; t < 3; ++t) area_x2f_accumulate(&rw[t][3 * hf], bitsf(ypack[4 * j + 1 + t]), t == 0, acc);
                                area_x2f_finish(acc, o6[hf]);
                            }
                        }
                        uint8_t b[12];
                        for (int q = 0; q < 6; ++q) { b[q] = (uint8_t)o6[0][q]; b[6 + q] = (uint8_t)o6[1][q]; }
                        if (!second) { b[6] = b[3]; b[7] = b[4]; b[8] = b[5]; b[9] = b[10] = b[11] = 0; }
                        memcpy(own[l], b, 12);
                    }
                    for (int l = 0; l < 32; ++l) {
                        const int ch = 30 * st - 1 + l;
                        const int cc = std::min(std::max(ch, 0), nchunks - 1);
                        const uint32_t from_left = own[l > 0 ? l - 1 : l][2], from_right = own[l < 31 ? l + 1 : l][0];
                        const uint32_t w0 = (cc == 0) ? (own[l][0] << 8) : (from_left & 0xFFFFFF00u);
                        const uint32_t w4 = (cc == nchunks - 1) ? (own[l][2] >> 8) : (from_right & 0x00FFFFFFu);
                        const uint32_t win[5] = {funnel_r(w0, own[l][0], 8), funnel_r(own[l][0], own[l][1], 8),
                                                 funnel_r(own[l][1], own[l][2], 8), funnel_r(own[l][2], w4, 8), w4 >> 8};
                        x2_expand24(win, (j & 1) ? xo[l] : xe[l]);
                    }
                }
                X2Row rc;
                rc.c0s = ly_rc[4 * r]; rc.c1s = ly_rc[4 * r + 1]; rc.k0 = ly_rc[4 * r + 2]; rc.k2 = ly_rc[4 * r + 3];
                for (int l = 1; l <= 30; ++l) {
                    const int ch = 30 * st - 1 + l;
                    if (ch < 0 || ch >= nchunks) continue;
                    const int nvalid = std::min(24, n - 24 * ch);
                    const float* xlo = (s0 & 1) ? xo[l] : xe[l];
                    const float* xhi = (s1 & 1) ? xo[l] : xe[l];
                    for (int q = 0; q < nvalid; ++q)
                        dst[(long)r * dst_pitch + 24 * ch + q] = (uint8_t)(x2_vertical(xlo[q], xhi[q], rc) & 0xFFu);
                }
            }
        }
    return 0;
}

// Replays lowres_x2f_kernel: the strips / bands / halo lanes of lowres_x2w_kernel with its own arithmetic pieces -- pair sums
// formed once per source row and carried from the last tap row of one low-res row to the first of the next
// (x2f_pairsums / x2f_mac), and the paired vertical stage (x2_vertical_pair).
extern "C" int emu_lowres_x2f(const uint8_t* src, uint8_t* dst, int h, int w, long src_pitch, long dst_pitch,
                              double factor, int band_rows) {
    std::vector<uint32_t> blob;
    DevShape sh;
    if (!build_lowres_shape(h, w, factor, 8, blob, &sh)) return 2;
    if (blob.empty()) blob.push_back(0);
    if (!sh.x2w) return 3;
    const uint32_t* tab = blob.data();
    const int n = 3 * w, nw = sh.nw;
    const bool fast2 = sh.area_mode == AREA_FAST2;
    const uint32_t* ypack = tab + sh.ay_pack;
    const uint32_t* ly_s = tab + sh.ly_s;
    const float* ly_rc = (const float*)(tab + sh.ly_rc);
    const uint32_t* ly_b = tab + sh.ly_b;
    const int nchunks = (w + 7) >> 3, nstrips = (nchunks + 29) / 30;
    for (int Y0 = 0; Y0 < h; Y0 += band_rows)
        for (int st = 0; st < nstrips; ++st) {
            const int Y1 = std::min(h, Y0 + band_rows);
            const int j_first = (int)(ly_s[Y0] & 0xFFFFu);
            float xe[32][24], xo[32][24];
            for (int l = 0; l < 32; ++l) for (int q = 0; q < 24; ++q) xe[l][q] = xo[l][q] = -1e30f;  // poison
            int have = j_first - 1;
            uint32_t carry[32][12];
            int carry_row[32];
            for (int l = 0; l < 32; ++l) carry_row[l] = -1;
            for (int r = Y0; r < Y1; ++r) {
                const uint32_t ys = ly_s[r];
                const int s0 = (int)(ys & 0xFFFFu), s1 = (int)(ys >> 16);
                while (have < s1) {
                    ++have;
                    const int j = have;
                    uint32_t own[32][3];
                    for (int l = 0; l < 32; ++l) {
                        const int ch = 30 * st - 1 + l;
                        const int cc = std::min(std::max(ch, 0), nchunks - 1);
                        const bool second = (nw - 4 * cc) >= 4;
                        const int sy0 = fast2 ? 2 * j : (int)ypack[4 * j];
                        const int rows[3] = {sy0, sy0 + 1, std::min(sy0 + 2, h - 1)};
                        uint32_t rw[3][6];
                        for (int t = 0; t < 3; ++t) {
                            const uint8_t* rp = src + (long)rows[t] * src_pitch + 24 * cc;
                            memcpy(&rw[t][0], rp, 12);
                            memcpy(&rw[t][3], rp + (second ? 12 : 0), 12);
                        }
                        uint32_t o6[2][6];
                        if (fast2) return 4;
                        float acc[12];
                        for (int t = 0; t < 3; ++t) {
                            if (!(t == 0 && rows[0] == carry_row[l])) x2f_pairsums(rw[t], carry[l]);
                            x2f_mac(carry[l], bitsf(ypack[4 * j + 1 + t]), t == 0, acc);
                            carry_row[l] = rows[t];
                        }
                        area_x2f_finish(acc, o6[0]);
                        area_x2f_finish(acc + 6, o6[1]);
                        uint8_t b[12];
                        for (int q = 0; q < 6; ++q) { b[q] = (uint8_t)o6[0][q]; b[6 + q] = (uint8_t)o6[1][q]; }
                        if (!second) { b[6] = b[3]; b[7] = b[4]; b[8] = b[5]; b[9] = b[10] = b[11] = 0; }
                        memcpy(own[l], b, 12);
                    }
                    for (int l = 0; l < 32; ++l) {
                        const int ch = 30 * st - 1 + l;
                        const int cc = std::min(std::max(ch, 0), nchunks - 1);
                        const uint32_t from_left = own[l > 0 ? l - 1 : l][2], from_right = own[l < 31 ? l + 1 : l][0];
                        const uint32_t w0 = (cc == 0) ? (own[l][0] << 8) : (from_left & 0xFFFFFF00u);
                        const uint32_t w4 = (cc == nchunks - 1) ? (own[l][2] >> 8) : (from_right & 0x00FFFFFFu);
                        const uint32_t win[5] = {funnel_r(w0, own[l][0], 8), funnel_r(own[l][0], own[l][1], 8),
                                                 funnel_r(own[l][1], own[l][2], 8), funnel_r(own[l][2], w4, 8), w4 >> 8};
                        x2_expand24(win, (j & 1) ? xo[l] : xe[l]);
                    }
                }
                X2Row rc;
                rc.c0s = ly_rc[4 * r]; rc.c1s = ly_rc[4 * r + 1]; rc.k0 = ly_rc[4 * r + 2]; rc.k2 = ly_rc[4 * r + 3];
                for (int l = 1; l <= 30; ++l) {
                    const int ch = 30 * st - 1 + l;
                    if (ch < 0 || ch >= nchunks) continue;
                    const int nvalid = std::min(24, n - 24 * ch);
                    const float* xlo = (s0 & 1) ? xo[l] : xe[l];
                    const float* xhi = (s1 & 1) ? xo[l] : xe[l];
                    const uint32_t cfix = x2_vertical_cfix(ly_b[r]);
                    for (int q = 0; q < nvalid; q += 2) {
                        const uint32_t pr = x2_vertical_pair(xlo[q], xhi[q], xlo[q + 1], xhi[q + 1], rc.c0s, rc.c1s, rc.k0 + 2.0f, cfix);
                        dst[(long)r * dst_pitch + 24 * ch + q] = (uint8_t)(pr >> 8);
                        dst[(long)r * dst_pitch + 24 * ch + q + 1] = (uint8_t)(pr >> 24);
                    }
                }
            }
        }
    return 0;
}

// Replays lowres_x2h_kernel (exact-2x widths with h = 2 nh + 1): the loop over LOW-RES rows of a band, source rows 2j+1 and
// 2j+2 read as a pair with row 2j's pair sums carried, and the per-low-res-row emission schedule of DevShape::hy_pack
// (rows that blend (j-1, j), then rows that blend (j, j)), clipped to the band.  Returns 3 if the shape is not eligible.
extern "C" int emu_lowres_x2h(const uint8_t* src, uint8_t* dst, int h, int w, long src_pitch, long dst_pitch,
                              double factor, int band_rows) {
    std::vector<uint32_t> blob;
    DevShape sh;
    if (!build_lowres_shape(h, w, factor, 8, blob, &sh)) return 2;
    if (blob.empty()) blob.push_back(0);
    if (!sh.x2h) return 3;
    const uint32_t* tab = blob.data();
    const int n = 3 * w, nw = sh.nw;
    const uint32_t* hyp = tab + sh.hy_pack;
    const uint32_t* ly_s = tab + sh.ly_s;
    const uint32_t* rc2 = tab + sh.ly_rc2;
    const int nchunks = (w + 7) >> 3, nstrips = (nchunks + 29) / 30;
    for (int Y0 = 0; Y0 < h; Y0 += band_rows)
        for (int st = 0; st < nstrips; ++st) {
            const int Y1 = std::min(h, Y0 + band_rows);
            const int j_first = (int)(ly_s[Y0] & 0xFFFFu), j_last = (int)(ly_s[Y1 - 1] >> 16);
            float xe[32][24], xo[32][24];
            for (int l = 0; l < 32; ++l) for (int q = 0; q < 24; ++q) xe[l][q] = xo[l][q] = -1e30f;  // poison
            uint32_t carry[32][12];
            auto lane_words = [&](int l, int row, uint32_t rw[6]) {
                const int cc = std::min(std::max(30 * st - 1 + l, 0), nchunks - 1);
                const bool second = (nw - 4 * cc) >= 4;
                const uint8_t* rp = src + (long)row * src_pitch + 24 * cc;
                memcpy(&rw[0], rp, 12);
                memcpy(&rw[3], rp + (second ? 12 : 0), 12);
            };
            for (int l = 0; l < 32; ++l) { uint32_t rw[6]; lane_words(l, 2 * j_first, rw); x2f_pairsums(rw, carry[l]); }
            int r = Y0;
            for (int j = j_first; j <= j_last; ++j) {
                if (2 * j + 2 > h - 1) return 5;
                const uint32_t* hp = hyp + 4 * j;
                uint32_t own[32][3];
                for (int l = 0; l < 32; ++l) {
                    const int cc = std::min(std::max(30 * st - 1 + l, 0), nchunks - 1);
                    const bool second = (nw - 4 * cc) >= 4;
                    uint32_t ra[6], rb[6], sa[12];
                    lane_words(l, 2 * j + 1, ra);
                    lane_words(l, 2 * j + 2, rb);
                    float acc[12];
                    x2f_mac(carry[l], bitsf(hp[0]), true, acc);
                    x2f_pairsums(ra, sa);
                    x2f_mac(sa, bitsf(hp[1]), false, acc);
                    x2f_pairsums(rb, carry[l]);
                    x2f_mac(carry[l], bitsf(hp[2]), false, acc);
                    uint32_t o6[2][6];
                    area_x2f_finish(acc, o6[0]);
                    area_x2f_finish(acc + 6, o6[1]);
                    uint8_t b[12];
                    for (int q = 0; q < 6; ++q) { b[q] = (uint8_t)o6[0][q]; b[6 + q] = (uint8_t)o6[1][q]; }
                    if (!second) { b[6] = b[3]; b[7] = b[4]; b[8] = b[5]; b[9] = b[10] = b[11] = 0; }
                    memcpy(own[l], b, 12);
                }
                float (*xnew)[24] = (j & 1) ? xo : xe;
                float (*xprev)[24] = (j & 1) ? xe : xo;
                for (int l = 0; l < 32; ++l) {
                    const int cc = std::min(std::max(30 * st - 1 + l, 0), nchunks - 1);
                    const uint32_t from_left = own[l > 0 ? l - 1 : l][2], from_right = own[l < 31 ? l + 1 : l][0];
                    const uint32_t w0 = (cc == 0) ? (own[l][0] << 8) : (from_left & 0xFFFFFF00u);
                    const uint32_t w4 = (cc == nchunks - 1) ? (own[l][2] >> 8) : (from_right & 0x00FFFFFFu);
                    const uint32_t win[5] = {funnel_r(w0, own[l][0], 8), funnel_r(own[l][0], own[l][1], 8),
                                             funnel_r(own[l][1], own[l][2], 8), funnel_r(own[l][2], w4, 8), w4 >> 8};
                    x2_expand24(win, xnew[l]);
                }
                const int r0 = (int)(hp[3] & 0xFFFFu), ra_end = r0 + (int)((hp[3] >> 16) & 0xFFu), rb_end = ra_end + (int)(hp[3] >> 24);
                const int a_end = std::min(ra_end, Y1), b_end = std::min(rb_end, Y1);
                for (int pass = 0; pass < 2; ++pass) {
                    const int end = pass == 0 ? a_end : b_end;
                    for (; r < end; ++r) {
                        if (r < r0) return 6;   // the schedule must never skip a row
                        float c0s, c1s, k0p;
                        memcpy(&c0s, &rc2[4 * r], 4); memcpy(&c1s, &rc2[4 * r + 1], 4); memcpy(&k0p, &rc2[4 * r + 2], 4);
                        const uint32_t cfix = rc2[4 * r + 3];
                        for (int l = 1; l <= 30; ++l) {
                            const int ch = 30 * st - 1 + l;
                            if (ch < 0 || ch >= nchunks) continue;
                            const int nvalid = std::min(24, n - 24 * ch);
                            const float* xlo = pass == 0 ? xprev[l] : xnew[l];
                            const float* xhi = xnew[l];
                            for (int q = 0; q < nvalid; q += 2) {
                                const uint32_t pr = x2_vertical_pair(xlo[q], xhi[q], xlo[q + 1], xhi[q + 1], c0s, c1s, k0p, cfix);
                                dst[(long)r * dst_pitch + 24 * ch + q] = (uint8_t)(pr >> 8);
                                dst[(long)r * dst_pitch + 24 * ch + q + 1] = (uint8_t)(pr >> 24);
                            }
                        }
                    }
                }
            }
            if (r != Y1) return 7;
        }
    return 0;
}

// Replays lowres_x2g_kernel (odd widths at factor 0.5): strips of 30 chunks with halo lanes; per lane the 27-byte source
// window, the three-tap float INTER_AREA x pass (x2g_hrow) on each tap row with the last tap row carried to the next
// low-res row, the float y pass, border replication of missing low-res pixels (x2g_replicate), the coefficient /
// slip-driven INTER_LINEAR x stage (x2g_expand24) and the general three-FMA y stage.  (The shared-memory staging and the
// byte-phase shifts of the kernel move bytes only; they are covered by the GPU tests.)
extern "C" int emu_lowres_x2g(const uint8_t* src, uint8_t* dst, int h, int w, long src_pitch, long dst_pitch, int band_rows) {
    std::vector<uint32_t> blob;
    DevShape sh;
    if (!build_lowres_shape(h, w, 0.5, 8, blob, &sh)) return 2;
    if (!sh.x2g) return 3;
    const uint32_t* tab = blob.data();
    const int n = 3 * w, nw = sh.nw;
    const uint32_t* ypack = tab + sh.ay_pack;
    const uint32_t* ly_s = tab + sh.ly_s;
    const float* ly_rc = (const float*)(tab + sh.ly_rc3);
    const float* xalpha = (const float*)(tab + sh.ax_alpha);
    const int32_t* lx_s0 = (const int32_t*)(tab + sh.lx_s0);
    const uint32_t* lx_a = tab + sh.lx_a;
    const int nchunks = (w + 7) >> 3, nstrips = (nchunks + 29) / 30;
    for (int Y0 = 0; Y0 < h; Y0 += band_rows)
        for (int st = 0; st < nstrips; ++st) {
            const int Y1 = std::min(h, Y0 + band_rows);
            const int j_first = (int)(ly_s[Y0] & 0xFFFFu);
            float xe[32][24], xo[32][24], carry[32][12];
            int carry_row[32];
            for (int l = 0; l < 32; ++l) { carry_row[l] = -1; for (int q = 0; q < 24; ++q) xe[l][q] = xo[l][q] = -1e30f; }
            int have = j_first - 1;
            for (int r = Y0; r < Y1; ++r) {
                const uint32_t ys = ly_s[r];
                const int s0 = (int)(ys & 0xFFFFu), s1 = (int)(ys >> 16);
                while (have < s1) {
                    ++have;
                    const int j = have;
                    uint32_t own[32][3];
                    for (int l = 0; l < 32; ++l) {
                        const int ch = 30 * st - 1 + l;
                        const int cc = std::min(std::max(ch, 0), nchunks - 1);
                        float al[12];
                        for (int q = 0; q < 4; ++q)
                            for (int tp = 0; tp < 3; ++tp) al[3 * q + tp] = xalpha[3 * std::min(4 * cc + q, nw - 1) + tp];
                        const int sy0 = (int)ypack[4 * j];
                        float acc[12];
                        for (int tp = 0; tp < 3; ++tp) {
                            const int row = (tp == 2) ? std::min(sy0 + 2, h - 1) : sy0 + tp;
                            if (!(tp == 0 && row == carry_row[l])) {
                                uint8_t win[28];
                                for (int k = 0; k < 28; ++k) {
                                    const int col = 24 * cc + k;
                                    win[k] = col < n ? src[(long)row * src_pitch + col] : (uint8_t)0xEE;  // beyond the row: garbage
                                }
                                uint32_t wn[7];
                                memcpy(wn, win, 28);
                                x2g_hrow(wn, al, carry[l]);
                            }
                            x2g_vmac(carry[l], bitsf(ypack[4 * j + 1 + tp]), tp == 0, acc);
                            carry_row[l] = row;
                        }
                        x2g_round12(acc, own[l]);
                    }
                    // shfl_up before the replication, shfl_down after it (x2g_expand)
                    uint32_t from_left[32];
                    for (int l = 0; l < 32; ++l) from_left[l] = own[l > 0 ? l - 1 : l][2];
                    for (int l = 0; l < 32; ++l) {
                        const int cc = std::min(std::max(30 * st - 1 + l, 0), nchunks - 1);
                        const int valid = nw - 4 * cc;
                        if (valid < 4) x2g_replicate(own[l], std::max(valid, 0), from_left[l]);
                    }
                    for (int l = 0; l < 32; ++l) {
                        const int cc = std::min(std::max(30 * st - 1 + l, 0), nchunks - 1);
                        uint32_t coef[8], slip = 0;
                        for (int x = 0; x < 8; ++x) {
                            const int xa = std::min(8 * cc + x, w - 1);
                            coef[x] = lx_a[xa];
                            if ((x & 1) && lx_s0[xa] == ((xa - 1) >> 1) - 1) slip |= 1u << (x >> 1);
                        }
                        const uint32_t from_right = own[l < 31 ? l + 1 : l][0];
                        const uint32_t w0 = (cc == 0) ? (own[l][0] << 8) : (from_left[l] & 0xFFFFFF00u);
                        const uint32_t w4 = (cc == nchunks - 1) ? (own[l][2] >> 8) : (from_right & 0x00FFFFFFu);
                        const uint32_t win[5] = {funnel_r(w0, own[l][0], 8), funnel_r(own[l][0], own[l][1], 8),
                                                 funnel_r(own[l][1], own[l][2], 8), funnel_r(own[l][2], w4, 8), w4 >> 8};
                        x2g_expand24(win, coef, slip, (j & 1) ? xo[l] : xe[l]);
                    }
                }
                X2Row rc;
                rc.c0s = ly_rc[4 * r]; rc.c1s = ly_rc[4 * r + 1]; rc.k0 = ly_rc[4 * r + 2]; rc.k2 = ly_rc[4 * r + 3];
                for (int l = 1; l <= 30; ++l) {
                    const int ch = 30 * st - 1 + l;
                    if (ch < 0 || ch >= nchunks) continue;
                    const int nvalid = std::min(24, n - 24 * ch);
                    const float* xlo = (s0 & 1) ? xo[l] : xe[l];
                    const float* xhi = (s1 & 1) ? xo[l] : xe[l];
                    for (int q = 0; q < nvalid; ++q)
                        dst[(long)r * dst_pitch + 24 * ch + q] = (uint8_t)(x2_vertical(xlo[q], xhi[q], rc) & 0xFFu);
                }
            }
        }
    return 0;
}

// Replays lowres_x2i_kernel (odd widths with a regular y axis): lowres_x2g_kernel's per-lane arithmetic inside the loop over
// LOW-RES rows -- fixed source-row pairs (odd h: rows 2j+1, 2j+2 with row 2j's horizontal pass carried; even h: rows 2j,
// 2j+1) and the emission schedule of DevShape::hy_pack clipped to the band.  Returns 3 if the shape is not eligible.
extern "C" int emu_lowres_x2i(const uint8_t* src, uint8_t* dst, int h, int w, long src_pitch, long dst_pitch, int band_rows) {
    std::vector<uint32_t> blob;
    DevShape sh;
    if (!build_lowres_shape(h, w, 0.5, 8, blob, &sh)) return 2;
    if (!sh.x2g || !sh.x2i) return 3;
    const bool carry_mode = sh.x2i == 1;
    const uint32_t* tab = blob.data();
    const int n = 3 * w, nw = sh.nw;
    const uint32_t* hyp = tab + sh.hy_pack;
    const uint32_t* ly_s = tab + sh.ly_s;
    const float* ly_rc = (const float*)(tab + sh.ly_rc3);
    const float* xalpha = (const float*)(tab + sh.ax_alpha);
    const int32_t* lx_s0 = (const int32_t*)(tab + sh.lx_s0);
    const uint32_t* lx_a = tab + sh.lx_a;
    const int nchunks = (w + 7) >> 3, nstrips = (nchunks + 29) / 30;
    for (int Y0 = 0; Y0 < h; Y0 += band_rows)
        for (int st = 0; st < nstrips; ++st) {
            const int Y1 = std::min(h, Y0 + band_rows);
            const int j_first = (int)(ly_s[Y0] & 0xFFFFu), j_last = (int)(ly_s[Y1 - 1] >> 16);
            float xe[32][24], xo[32][24], carry[32][12];
            for (int l = 0; l < 32; ++l) for (int q = 0; q < 24; ++q) xe[l][q] = xo[l][q] = -1e30f;
            auto hrow = [&](int l, int row, float out[12]) {
                const int cc = std::min(std::max(30 * st - 1 + l, 0), nchunks - 1);
                float al[12];
                for (int q = 0; q < 4; ++q)
                    for (int tp = 0; tp < 3; ++tp) al[3 * q + tp] = xalpha[3 * std::min(4 * cc + q, nw - 1) + tp];
                uint8_t win[28];
                for (int k = 0; k < 28; ++k) {
                    const int col = 24 * cc + k;
                    win[k] = col < n ? src[(long)row * src_pitch + col] : (uint8_t)0xEE;  // beyond the row: garbage
                }
                uint32_t wn[7];
                memcpy(wn, win, 28);
                x2g_hrow(wn, al, out);
            };
            if (carry_mode)
                for (int l = 0; l < 32; ++l) hrow(l, 2 * j_first, carry[l]);
            int r = Y0;
            for (int j = j_first; j <= j_last; ++j) {
                const uint32_t* hp = hyp + 4 * j;
                if ((carry_mode ? 2 * j + 2 : 2 * j + 1) > h - 1) return 5;
                uint32_t own[32][3];
                for (int l = 0; l < 32; ++l) {
                    float acc[12], ha[12];
                    if (carry_mode) {
                        x2g_vmac(carry[l], bitsf(hp[0]), true, acc);
                        hrow(l, 2 * j + 1, ha);
                        x2g_vmac(ha, bitsf(hp[1]), false, acc);
                        hrow(l, 2 * j + 2, carry[l]);
                        x2g_vmac(carry[l], bitsf(hp[2]), false, acc);
                    } else {
                        hrow(l, 2 * j, ha);
                        x2g_vmac(ha, bitsf(hp[0]), true, acc);
                        hrow(l, 2 * j + 1, ha);
                        x2g_vmac(ha, bitsf(hp[1]), false, acc);
                    }
                    x2g_round12(acc, own[l]);
                }
                uint32_t from_left[32];
                for (int l = 0; l < 32; ++l) from_left[l] = own[l > 0 ? l - 1 : l][2];
                for (int l = 0; l < 32; ++l) {
                    const int cc = std::min(std::max(30 * st - 1 + l, 0), nchunks - 1);
                    const int valid = nw - 4 * cc;
                    if (valid < 4) x2g_replicate(own[l], std::max(valid, 0), from_left[l]);
                }
                float (*xnew)[24] = (j & 1) ? xo : xe;
                float (*xprev)[24] = (j & 1) ? xe : xo;
                for (int l = 0; l < 32; ++l) {
                    const int cc = std::min(std::max(30 * st - 1 + l, 0), nchunks - 1);
                    uint32_t coef[8], slip = 0;
                    for (int x = 0; x < 8; ++x) {
                        const int xa = std::min(8 * cc + x, w - 1);
                        coef[x] = lx_a[xa];
                        if ((x & 1) && lx_s0[xa] == ((xa - 1) >> 1) - 1) slip |= 1u << (x >> 1);
                    }
                    const uint32_t from_right = own[l < 31 ? l + 1 : l][0];
                    const uint32_t w0 = (cc == 0) ? (own[l][0] << 8) : (from_left[l] & 0xFFFFFF00u);
                    const uint32_t w4 = (cc == nchunks - 1) ? (own[l][2] >> 8) : (from_right & 0x00FFFFFFu);
                    const uint32_t win[5] = {funnel_r(w0, own[l][0], 8), funnel_r(own[l][0], own[l][1], 8),
                                             funnel_r(own[l][1], own[l][2], 8), funnel_r(own[l][2], w4, 8), w4 >> 8};
                    x2g_expand24(win, coef, slip, xnew[l]);
                }
                const int r0 = (int)(hp[3] & 0xFFFFu), ra_end = r0 + (int)((hp[3] >> 16) & 0xFFu), rb_end = ra_end + (int)(hp[3] >> 24);
                const int a_end = std::min(ra_end, Y1), b_end = std::min(rb_end, Y1);
                for (int pass = 0; pass < 2; ++pass) {
                    const int end = pass == 0 ? a_end : b_end;
                    for (; r < end; ++r) {
                        if (r < r0) return 6;
                        X2Row rc;
                        rc.c0s = ly_rc[4 * r]; rc.c1s = ly_rc[4 * r + 1]; rc.k0 = ly_rc[4 * r + 2]; rc.k2 = ly_rc[4 * r + 3];
                        for (int l = 1; l <= 30; ++l) {
                            const int ch = 30 * st - 1 + l;
                            if (ch < 0 || ch >= nchunks) continue;
                            const int nvalid = std::min(24, n - 24 * ch);
                            const float* xlo = pass == 0 ? xprev[l] : xnew[l];
                            for (int q = 0; q < nvalid; ++q)
                                dst[(long)r * dst_pitch + 24 * ch + q] = (uint8_t)(x2_vertical(xlo[q], xnew[l][q], rc) & 0xFFu);
                        }
                    }
                }
            }
            if (r != Y1) return 7;
        }
    return 0;
}

// Replays lowres_x2p_kernel (packed-integer pipeline for shapes that are exact 2x in both axes): the same band x strip
// tiles and 32-lane strips as lowres_x2w_kernel; per low-res row every lane forms its packed words (rod_core.h x2p_*),
// the neighbour words arrive by "shuffle" (lane 0 / 31 get their own value back, like shfl.up / shfl.down).
extern "C" int emu_lowres_x2p(const uint8_t* src, uint8_t* dst, int h, int w, long src_pitch, long dst_pitch, int band_rows) {
    std::vector<uint32_t> blob;
    DevShape sh;
    if (!build_lowres_shape(h, w, 0.5, 8, blob, &sh)) return 2;
    if (!sh.x2p) return 3;
    const int n = 3 * w, nw = w >> 1, nh = h >> 1;
    const int nchunks = (w + 7) >> 3, nstrips = (nchunks + 29) / 30;
    struct Slot { uint32_t A[12], Bp[12]; };
    for (int Y0 = 0; Y0 < h; Y0 += band_rows)
        for (int st = 0; st < nstrips; ++st) {
            const int Y1 = std::min(h, Y0 + band_rows);
            const int j0 = Y0 >> 1, jfirst = std::max(j0 - 1, 0), jlast = std::min(Y1 >> 1, nh - 1);
            std::vector<Slot> even(32), odd(32);
            memset(even.data(), 0xCD, sizeof(Slot) * 32);  // poison
            memset(odd.data(), 0xCD, sizeof(Slot) * 32);
            auto store = [&](int y, const std::vector<Slot>& nearS, const std::vector<Slot>& farS) {
                for (int l = 1; l <= 30; ++l) {
                    const int ch = 30 * st - 1 + l;
                    if (ch < 0 || ch >= nchunks) continue;
                    const int nvalid = std::min(24, n - 24 * ch);
                    uint32_t wds[6];
                    x2p_emit(nearS[l].Bp, farS[l].A, wds);
                    memcpy(dst + (long)y * dst_pitch + 24 * ch, wds, nvalid);
                }
            };
            for (int j = jfirst; j <= jlast; ++j) {
                uint32_t b[32][6];
                for (int l = 0; l < 32; ++l) {
                    const int ch = 30 * st - 1 + l;
                    const int cc = std::min(std::max(ch, 0), nchunks - 1);
                    const bool second = (nw - 4 * cc) >= 4;
                    uint32_t rw[2][6];
                    for (int t = 0; t < 2; ++t) {
                        const uint8_t* rp = src + (long)(2 * j + t) * src_pitch + 24 * cc;
                        memcpy(&rw[t][0], rp, 12);
                        memcpy(&rw[t][3], rp + (second ? 12 : 0), 12);
                    }
                    x2p_area(rw[0], rw[1], b[l]);
                    if (!second) x2p_patch_two_pixel(b[l]);
                }
                std::vector<Slot>& cur = (j & 1) ? odd : even;
                for (int l = 0; l < 32; ++l) {
                    const int ch = 30 * st - 1 + l;
                    const int cc = std::min(std::max(ch, 0), nchunks - 1);
                    const int ll = l > 0 ? l - 1 : l, lr = l < 31 ? l + 1 : l;
                    x2p_build(b[l], b[ll][3], x2p_n10(b[ll]), b[lr][4], b[lr][0], cc == 0, cc == nchunks - 1, cur[l].A, cur[l].Bp);
                }
                const bool up = j > j0, down = j >= j0 && 2 * j < Y1;
                if (j & 1) {
                    if (up) store(2 * j - 1, even, odd);
                    if (down) store(2 * j, odd, even);
                } else {
                    if (up) store(2 * j - 1, odd, even);
                    if (down) store(2 * j, even, j == 0 ? even : odd);
                }
            }
            if (Y1 == h) store(h - 1, ((nh - 1) & 1) ? odd : even, ((nh - 1) & 1) ? odd : even);
        }
    return 0;
}

extern "C" int emu_lowres(const uint8_t* src, uint8_t* dst, int h, int w, long src_pitch, long dst_pitch,
                          double factor, int src_phase) {
    std::vector<uint32_t> blob;
    DevShape sh;
    if (!build_lowres_shape(h, w, factor, 8, blob, &sh)) return 2;
    if (blob.empty()) blob.push_back(0);
    choose_strip_rows(&sh, blob.data());
    if (sh.strip_rows > 0) return emu_lowres_x2(src, dst, h, w, src_pitch, dst_pitch, sh, blob.data(), src_phase);
    return emu_lowres_generic(src, dst, h, w, src_pitch, dst_pitch, factor, src_phase);
}

// The generic tiled kernel on any shape (exact-2x shapes included), for coverage of both kernels.
extern "C" int emu_lowres_tiled(const uint8_t* src, uint8_t* dst, int h, int w, long src_pitch, long dst_pitch,
                                double factor, int src_phase) {
    return emu_lowres_generic(src, dst, h, w, src_pitch, dst_pitch, factor, src_phase);
}

// Replays noise_kernel (compat / philox / field) over one image's flat element range.
extern "C" int emu_noise(const uint8_t* src, uint8_t* dst, const float* noise, float* field_out, long n_elems,
                         float sigma, uint64_t seed, uint64_t image_index, uint32_t offset) {
    const float K = sigma * ROD_NOISE_K_PER_SIGMA;
    for (long g = 0; g < (n_elems + 7) / 8; ++g) {
        float sf[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (noise == nullptr) {
            uint32_t r[4], t[4];
            const uint32_t ig_lo = (uint32_t)image_index, ig_hi = (uint32_t)(image_index >> 32);
            philox4x32_10((uint32_t)g, ig_lo, ig_hi, offset, (uint32_t)seed, (uint32_t)(seed >> 32), r);
            if (philox_needs_tail(r)) {
                philox4x32_10((uint32_t)g, ig_lo, ig_hi ^ ROD_PHILOX_TAIL_FLIP, offset, (uint32_t)seed,
                              (uint32_t)(seed >> 32), t);
                gauss8(r, t, sf);
            } else {
                gauss8(r, nullptr, sf);
            }
        }
        for (int j = 0; j < 8; ++j) {
            long e = 8 * g + j;
            if (e >= n_elems) break;
            if (noise) {
                if (field_out) field_out[e] = noise[e];
                if (dst) dst[e] = (uint8_t)noise_px((float)src[e], noise[e]);
            } else {
                if (field_out) field_out[e] = K * sf[j];
                if (dst) dst[e] = (uint8_t)noise_philox_px(src[e], sf[j], K);
            }
        }
    }
    return 0;
}

// Replays noise_table_kernel (the TABLE generator of Philox mode): per group of 16 elements the Philox block, per word
// four 8-bit table draws and their integer Hadamard mix through the same 32-bit two-form arithmetic the kernel uses
// (rod_core.h h4_word_k).  field_out receives k.  rounds: 10, or 7 (ROD_GAUSS_TABLE_PHILOX7).
extern "C" int emu_noise_table(const uint8_t* src, uint8_t* dst, float* field_out, long n_elems, float sigma,
                               uint64_t seed, uint64_t image_index, uint32_t offset, int rounds) {
    if (!(sigma >= ROD_GAUSS_TABLE_MIN_SIGMA && sigma <= ROD_GAUSS_TABLE_MAX_SIGMA)) return 1;
    int32_t X[256];
    build_gauss_table(sigma, X);
    const PhiloxKeys keys = philox_round_keys((uint32_t)seed, (uint32_t)(seed >> 32));
    const uint32_t ig_lo = (uint32_t)image_index, ig_hi = (uint32_t)(image_index >> 32);
    for (long g = 0; g < (n_elems + 15) / 16; ++g) {
        uint32_t r[4];
        if (rounds == 7) philox4x32_rk<7>((uint32_t)g, ig_lo, ig_hi, offset, keys, r);
        else philox4x32_rk<10>((uint32_t)g, ig_lo, ig_hi, offset, keys, r);
        for (int j = 0; j < 16; ++j) {
            const long e = 16 * g + j;
            if (e >= n_elems) break;
            const int k = h4_word_k(r[j >> 2], X, j & 3);
            if (field_out) field_out[e] = (float)k;
            if (dst) dst[e] = (uint8_t)noise_table_px(src[e], k);
        }
    }
    return 0;
}

// The table itself, for a direct comparison with the oracle's scipy-based one.
extern "C" int emu_gauss_table(float sigma, int32_t* out) {
    build_gauss_table(sigma, out);
    return 0;
}

// Replays filter2d_kernel: tiles of kF2dTH rows x kF2dTWB bytes staged with halo as floats (reflect-101 per pixel and
// per row), taps in row-major order, fused multiply-add below 4 * (3w / 4), separate multiply / add in the row tail.
extern "C" int emu_filter2d(const uint8_t* src, uint8_t* dst, int h, int w, long src_pitch, long dst_pitch,
                            const float* kernel, int k) {
    const int TH = 16, TWB = 768;
    struct Tap { int dy, dxb; float w; };
    std::vector<Tap> taps;
    for (int dy = 0; dy < k; ++dy)
        for (int dx = 0; dx < k; ++dx)
            if (kernel[dy * k + dx] != 0.0f) taps.push_back(Tap{dy, 3 * dx, kernel[dy * k + dx]});
    const int r = k >> 1, n = 3 * w, tp = TWB + 3 * (k - 1), n_vec = n & ~3;
    std::vector<float> tile((size_t)(TH + k - 1) * tp);
    for (int y0 = 0; y0 < h; y0 += TH)
        for (int b0 = 0; b0 < n; b0 += TWB) {
            const int th = std::min(TH, h - y0), twb = std::min(TWB, n - b0);
            std::fill(tile.begin(), tile.end(), -1e30f);
            for (int rr = 0; rr < th + 2 * r; ++rr) {
                const uint8_t* srow = src + (long)reflect101(y0 + rr - r, h) * src_pitch;
                for (int ci = 0; ci < twb + 6 * r; ++ci) {
                    const int i = b0 + ci, px = i / 3, c = i - 3 * px;
                    tile[(size_t)rr * tp + ci] = (float)srow[3 * reflect101(px - r, w) + c];
                }
            }
            for (int ry = 0; ry < th; ++ry)
                for (int x = 0; x < twb; ++x) {
                    float s = 0.f;
                    for (const Tap& t : taps) {
                        const float v = tile[(size_t)(ry + t.dy) * tp + x + t.dxb];
                        s = (b0 + x >= n_vec) ? fadd(s, fmul(v, t.w)) : fmaf(v, t.w, s);
                    }
                    float v = frint(s);
                    v = v < 0.f ? 0.f : (v > 255.f ? 255.f : v);
                    dst[(long)(y0 + ry) * dst_pitch + b0 + x] = (uint8_t)(int)v;
                }
        }
    return 0;
}

// Letterbox tables + per-pixel arithmetic (replays lb_pixel); out is uint8 HWC canvas (pre-normalise).
extern "C" int emu_letterbox_u8(const uint8_t* img, int h, int w, long pitch, uint8_t* canvas, int out_h, int out_w,
                                int pad) {
    int new_h, new_w, top, left;
    letterbox_geometry(h, w, out_h, out_w, &new_h, &new_w, &top, &left);
    const bool identity = (new_h == h && new_w == w);
    const bool area2 = (!identity && h == 2 * new_h && w == 2 * new_w);
    LinearAxis lx, ly;
    if (!identity && !area2) {
        lx = build_linear_axis(w, new_w, true);
        ly = build_linear_axis(h, new_h, false);
    }
    for (int Y = 0; Y < out_h; ++Y)
        for (int X = 0; X < out_w; ++X) {
            uint8_t* o = canvas + ((long)Y * out_w + X) * 3;
            const int cy = Y - top, cx = X - left;
            if (cy < 0 || cy >= new_h || cx < 0 || cx >= new_w) { o[0] = o[1] = o[2] = (uint8_t)pad; continue; }
            for (int c = 0; c < 3; ++c) {
                if (identity) { o[c] = img[(long)cy * pitch + cx * 3 + c]; continue; }
                if (area2) {
                    const uint8_t* s = img + (long)(2 * cy) * pitch + (2 * cx) * 3;
                    o[c] = (uint8_t)(((uint32_t)s[c] + s[3 + c] + s[pitch + c] + s[pitch + 3 + c] + 2u) >> 2);
                    continue;
                }
                const int s0 = lx.s0[cx], s1 = std::min(s0 + 1, w - 1);
                const uint8_t* r0 = img + (long)ly.s0[cy] * pitch;
                const uint8_t* r1 = img + (long)ly.s1[cy] * pitch;
                const uint32_t h0 = linear_h4(r0[3 * s0 + c], r0[3 * s1 + c], lx.coef[cx]);
                const uint32_t h1 = linear_h4(r1[3 * s0 + c], r1[3 * s1 + c], lx.coef[cx]);
                o[c] = (uint8_t)linear_v(h0, h1, ly.coef[cy]);
            }
        }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// JPEG: sequential replay of the device encoder (csrc/jpeg.cu) from the same rod_jpeg.h arithmetic, in the same
// passes: (1) quantised coefficients of every block incl. libjpeg's dummy blocks at the right / bottom edge,
// (2) bit length of every MCU, (3) prefix sum, (4) every MCU written at its bit offset, (5) final 1-padding, 0xFF
// stuffing, EOI.  `header` = the bytes SOI .. SOS that OpenCV writes for this size.  Returns the file length, or < 0.
#include "../../robust-object-detection_b200/csrc/rod_jpeg_host.h"

namespace {
struct WordWriter {   // MSB-first bits into a zero-initialised big-endian byte stream, at an absolute bit offset
    uint8_t* buf;
    uint64_t pos;
    void put(uint32_t code, int size) {
        for (int b = size - 1; b >= 0; --b, ++pos)
            if ((code >> b) & 1u) buf[pos >> 3] |= (uint8_t)(0x80u >> (pos & 7));
    }
};
}  // namespace

extern "C" long emu_jpeg_encode(const uint8_t* bgr, int h, int w, long pitch, const uint8_t* header, long header_len,
                                uint8_t* out, long out_cap) {
    using namespace rod::jpeg;
    HeaderInfo info;
    Tables tb;
    if (!parse_header(header, (size_t)header_len, &info, &tb)) return -1;
    if (info.height != h || info.width != w) return -2;
    const Geometry g = geometry(h, w);
    const int n_mcu = g.mcu_w * g.mcu_h;
    std::vector<int16_t> coef((size_t)n_mcu * 6 * 64);
    // pass 1: coefficients (zigzag order); dummy luma blocks: zero AC, DC of the preceding block of the MCU
    for (int my = 0; my < g.mcu_h; ++my)
        for (int mx = 0; mx < g.mcu_w; ++mx) {
            int16_t* mc = coef.data() + ((size_t)my * g.mcu_w + mx) * 6 * 64;
            for (int blk = 0; blk < 6; ++blk) {
                int16_t* zz = mc + 64 * blk;
                const bool real = blk >= 4 || (2 * mx + (blk & 1) < g.yblk_w && 2 * my + (blk >> 1) < g.yblk_h);
                if (!real) {
                    memset(zz, 0, 128);
                    zz[0] = (zz - 64)[0];
                    continue;
                }
                int d[64];
                block_samples(bgr, pitch, g, mx, my, blk, d);
                fdct_islow(d);
                const int t = blk < 4 ? 0 : 1;
                for (int z = 0; z < 64; ++z) {
                    const int nat = natural_order(z);
                    zz[z] = (int16_t)quantize(d[nat], tb.recip[t][nat], tb.corr[t][nat], tb.shift[t][nat]);
                }
            }
        }
    // pass 2 + 3: bit length per MCU, offsets.  DC predictor of a block = DC of the previous block of its component.
    auto encode_mcu = [&](int m, auto& sink) {
        const int16_t* mc = coef.data() + (size_t)m * 6 * 64;
        for (int blk = 0; blk < 6; ++blk) {
            int last;
            if (blk >= 1 && blk <= 3) last = mc[64 * (blk - 1)];
            else if (m == 0) last = 0;
            else last = (mc - 6 * 64)[64 * (blk == 0 ? 3 : blk)];
            const int hs = blk < 4 ? 0 : 1;
            encode_block(mc + 64 * blk, last, tb.ehufco[hs], tb.ehufsi[hs], tb.ehufco[2 + hs], tb.ehufsi[2 + hs], sink);
        }
    };
    std::vector<uint64_t> off((size_t)n_mcu + 1, 0);
    for (int m = 0; m < n_mcu; ++m) {
        BitCounter bc;
        encode_mcu(m, bc);
        off[m + 1] = off[m] + bc.bits;
    }
    const uint64_t total_bits = off[n_mcu];
    const size_t raw_bytes = (size_t)((total_bits + 7) >> 3);
    std::vector<uint8_t> raw(raw_bytes + 8, 0);
    for (int m = 0; m < n_mcu; ++m) {
        WordWriter ww{raw.data(), off[m]};
        encode_mcu(m, ww);
    }
    if (total_bits & 7) raw[raw_bytes - 1] |= (uint8_t)(0xFFu >> (total_bits & 7));   // flush_bits: pad with ones
    // pass 5: stuffing
    long o = 0;
    if (header_len > out_cap) return -3;
    memcpy(out, header, (size_t)header_len);
    o = header_len;
    for (size_t i = 0; i < raw_bytes; ++i) {
        if (o + 4 > out_cap) return -3;
        out[o++] = raw[i];
        if (raw[i] == 0xFF) out[o++] = 0x00;
    }
    out[o++] = 0xFF; out[o++] = 0xD9;
    return o;
}

// ---------------------------------------------------------------------------------------------------------------
// JPEG decoding: sequential replay of the device decoder (csrc/jpegdec.cu) from the same rod_jpegdec.h arithmetic, in the
// same passes: (1) markers, tables, unstuffed scan (host side of the product too), (2) Huffman decoding into coefficient
// blocks, (3) dequantisation + islow IDCT into the Y / Cb / Cr planes, (4) fancy chroma upsampling + colour conversion.
// Returns 0 and fills out (HWC BGR, pitch 3 * w, capacity out_cap bytes) and hw[2]; 1 / 2: ParseStatus; 3: the scan does
// not end in EOI; 4 / 5: corrupt / short stream; 6: out too small.
#include "../../robust-object-detection_b200/csrc/rod_jpegdec_host.h"

// mode 0: the sequential scan decoder (decode_scan); mode 1: the device's passes (self-synchronising subsequences, block
// scan, final pass, DC prefix sums) with the subsequences of a round visited in DESCENDING order, so that a round
// propagates a corrected state by one subsequence only (the slowest schedule the concurrent device threads can produce);
// hw[2] receives the number of rounds.
extern "C" int emu_jpeg_decode_mode(const uint8_t* file, long n, uint8_t* out, long out_cap, int* hw, int mode);
extern "C" int emu_jpeg_decode(const uint8_t* file, long n, uint8_t* out, long out_cap, int* hw) {
    int hw3[3];
    const int rc = emu_jpeg_decode_mode(file, n, out, out_cap, hw3, 0);
    hw[0] = hw3[0]; hw[1] = hw3[1];
    return rc;
}
extern "C" int emu_jpeg_decode_mode(const uint8_t* file, long n, uint8_t* out, long out_cap, int* hw, int mode) {
    using namespace rod::jpegdec;
    FileInfo info;
    std::vector<TableSet> tsv(1);
    TableSet& ts = tsv[0];
    const ParseStatus st = parse_file(file, (size_t)n, &info, &ts);
    if (st != PARSE_OK) return (int)st;
    hw[0] = info.height; hw[1] = info.width; hw[2] = 0;
    std::vector<uint8_t> stream((size_t)n - info.scan_begin + 64);
    std::vector<uint32_t> rst;
    const size_t sb = unstuff_scan(file, (size_t)n, info.scan_begin, stream.data(), &rst);
    if (sb == (size_t)-1) return 3;
    ImageRec im;
    memset(&im, 0, sizeof(im));
    im.h = info.height; im.w = info.width; im.stream_bytes = (uint32_t)sb;
    im.hs = (uint8_t)info.hs; im.vs = (uint8_t)info.vs; im.ncomp = (uint8_t)info.ncomp;
    const Layout L = layout_of(im);
    const int mcu_w = L.mcu_w, mcu_h = L.mcu_h;
    std::vector<SegRec> segs;
    if (!make_segments(im, 0, info.restart_interval, 0, sb, rst, &segs)) return 3;
    std::vector<int16_t> coef((size_t)L.mcus * L.nb * 64, 0);
    uint8_t nat[64];
    for (int z = 0; z < 64; ++z) nat[z] = (uint8_t)rod::jpeg::natural_order(z);
    hw[2] = 0;
    int rc = 0;
    if (mode == 0) {
        rc = decode_scan(im, segs.data(), (int)segs.size(), ts, nat, stream.data(), coef.data());
    } else {
        int err = 0;
        for (const SegRec& sg : segs) {   // every restart interval is a stream of its own
            const uint8_t* sbase = stream.data() + sg.stream_off;
            const uint32_t bit0 = 8u * sg.byte0, total_bits = 8u * sg.stream_bytes - bit0;
            const uint32_t n_sub = total_bits ? (total_bits + kSubBits - 1) / kSubBits : 1;
            std::vector<uint64_t> E(n_sub), U(n_sub);
            for (uint32_t q = 0; q < n_sub; ++q) {
                U[q] = span_state(bit0 + q * kSubBits, 0, 0, 0);
                E[q] = decode_span<false>(L, sg.stream_bytes, ts, nat, sbase, U[q], bit0 + (q + 1) * kSubBits, nullptr, 0, 0, nullptr);
            }
            int rounds = 0;
            for (bool changed = true; changed;) {
                changed = false;
                ++rounds;
                for (uint32_t q = n_sub - 1; q >= 1; --q) {
                    const uint64_t st = state_start(E[q - 1]);
                    if (st == U[q]) continue;
                    U[q] = st;
                    E[q] = decode_span<false>(L, sg.stream_bytes, ts, nat, sbase, st, bit0 + (q + 1) * kSubBits, nullptr, 0, 0, nullptr);
                    changed = true;
                }
            }
            if (rounds > hw[2]) hw[2] = rounds;
            std::vector<uint32_t> blk0(n_sub + 1, 0);
            for (uint32_t q = 0; q < n_sub; ++q) blk0[q + 1] = blk0[q] + state_nblk(E[q]);
            if (blk0[n_sub] < sg.n_blocks) { rc = 2; break; }
            for (uint32_t q = 0; q < n_sub; ++q)
                if (blk0[q] < sg.n_blocks)
                    decode_span<true>(L, sg.stream_bytes, ts, nat, sbase, q == 0 ? span_state(bit0, 0, 0, 0) : state_start(E[q - 1]),
                                      bit0 + (q + 1) * kSubBits, coef.data(), sg.first_block + blk0[q], sg.first_block + sg.n_blocks, &err);
            // DC prediction: prefix sums per component in coded order, from 0 in every interval
            int pred[3] = {0, 0, 0};
            for (uint32_t g = sg.first_block; g < sg.first_block + sg.n_blocks; ++g) {
                int16_t* blk = block_of(coef.data(), L, g);
                const int b = (int)(g % (uint32_t)L.nb), comp = b < L.nl ? 0 : b - L.nl + 1;
                pred[comp] += blk[0];
                blk[0] = (int16_t)pred[comp];
            }
        }
        if (err && rc == 0) rc = 1;
    }
    if (rc) return 3 + rc;
    const long ypitch = 8L * L.hs * mcu_w, cpitch = 8L * mcu_w;
    std::vector<uint8_t> yp((size_t)ypitch * 8 * L.vs * mcu_h), cbp((size_t)cpitch * 8 * mcu_h, 128), crp((size_t)cpitch * 8 * mcu_h, 128);
    for (int by = 0; by < L.vs * mcu_h; ++by)
        for (int bx = 0; bx < L.hs * mcu_w; ++bx)
            idct_islow(coef.data() + 64 * ((size_t)by * L.hs * mcu_w + bx), ts.quant[0], yp.data() + 8L * by * ypitch + 8 * bx, ypitch);
    if (im.ncomp == 3) {
        const int16_t* cbc = coef.data() + 64 * (size_t)L.nl * L.mcus;
        const int16_t* crc = cbc + 64 * (size_t)L.mcus;
        for (int by = 0; by < mcu_h; ++by)
            for (int bx = 0; bx < mcu_w; ++bx) {
                idct_islow(cbc + 64 * ((size_t)by * mcu_w + bx), ts.quant[1], cbp.data() + 8L * by * cpitch + 8 * bx, cpitch);
                idct_islow(crc + 64 * ((size_t)by * mcu_w + bx), ts.quant[2], crp.data() + 8L * by * cpitch + 8 * bx, cpitch);
            }
    }
    if ((long)im.h * im.w * 3 > out_cap) return 6;
    const int cw = (im.w + L.hs - 1) / L.hs, ch = (im.h + L.vs - 1) / L.vs;
    for (int y = 0; y < im.h; ++y)
        for (int x0 = 0; x0 < im.w; x0 += 4) {   // four pixels at a time, like a thread of jpegdec_color_kernel
            int cb[4] = {128, 128, 128, 128}, cr[4] = {128, 128, 128, 128};
            if (im.ncomp == 3) {
                chroma_quad(cbp.data(), cpitch, L.hs, L.vs, cw, ch, x0, y, cb);
                chroma_quad(crp.data(), cpitch, L.hs, L.vs, cw, ch, x0, y, cr);
            }
            for (int i = 0; i < 4 && x0 + i < im.w; ++i) {
                uint8_t* o = out + ((size_t)y * im.w + x0 + i) * 3;
                const int yy = yp[(size_t)y * ypitch + x0 + i];
                if (im.ncomp == 1) { o[0] = o[1] = o[2] = (uint8_t)yy; continue; }
                ycc_to_bgr(yy, cb[i], cr[i], o);
                // (the per-pixel form of the same arithmetic must agree)
                uint8_t q[3];
                ycc_to_bgr(yy, chroma_at(cbp.data(), cpitch, L.hs, L.vs, cw, ch, x0 + i, y), chroma_at(crp.data(), cpitch, L.hs, L.vs, cw, ch, x0 + i, y), q);
                if (q[0] != o[0] || q[1] != o[1] || q[2] != o[2]) return 9;
            }
        }
    return 0;
}
