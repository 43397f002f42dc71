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
static const int kLowresTH = 32, kLowresTW = 128;

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

// Replays lowres_kernel for one image (tables from rod_tables.h, phases B / C1 / C2 per tile).
extern "C" int emu_lowres(const uint8_t* src, uint8_t* dst, int h, int w, long src_pitch, long dst_pitch,
                          double factor, int dst_phase) {
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
    // worst-case footprint exactly as ensure_lowres_tables computes it
    int max_rows = 1, max_cols = 1;
    for (int y0 = 0; y0 < h; y0 += kLowresTH) {
        const int y1 = std::min(h, y0 + kLowresTH) - 1;
        max_rows = std::max(max_rows, (int)(ly_s[y1] >> 16) - (int)(ly_s[y0] & 0xFFFF) + 1);
    }
    for (int x0 = 0; x0 < w; x0 += kLowresTW) {
        const int x1 = std::min(w, x0 + kLowresTW) - 1;
        max_cols = std::max(max_cols, std::min(lx_s0[x1] + 1, sh.nw - 1) - lx_s0[x0] + 1);
    }
    const int half_pitch = (max_cols * 3 + 15) & ~15;
    const int hx_pitch = kLowresTW * 3 + 8;
    std::vector<uint8_t> half((size_t)max_rows * half_pitch);
    std::vector<uint16_t> hx((size_t)max_rows * hx_pitch);
    for (int y0 = 0; y0 < h; y0 += kLowresTH)
        for (int x0 = 0; x0 < w; x0 += kLowresTW) {
            const int th = std::min(kLowresTH, h - y0), tw = std::min(kLowresTW, w - x0);
            const int tw3 = tw * 3;
            const int j_lo = (int)(ly_s[y0] & 0xFFFFu), j_hi = (int)(ly_s[y0 + th - 1] >> 16);
            const int i_lo = lx_s0[x0], i_hi = std::min(lx_s0[x0 + tw - 1] + 1, sh.nw - 1);
            const int nj = j_hi - j_lo + 1, ni3 = (i_hi - i_lo + 1) * 3;
            if (nj > max_rows || ni3 > max_cols * 3) return 3;
            std::fill(half.begin(), half.end(), 0xEE);
            std::fill(hx.begin(), hx.end(), 0xEEEE);
            for (int idx = 0; idx < nj * ni3; ++idx) {
                const int jr = idx / ni3, o = idx - jr * ni3;
                const int ir = o / 3, c = o - 3 * ir;
                half[jr * half_pitch + o] = (uint8_t)area_value(src, src_pitch, sh, tab, j_lo + jr, i_lo + ir, c);
            }
            for (int idx = 0; idx < nj * tw3; ++idx) {
                const int jr = idx / tw3, o = idx - jr * tw3;
                const int xr = o / 3, c = o - 3 * xr;
                const int s0 = lx_s0[x0 + xr];
                const int s1 = std::min(s0 + 1, sh.nw - 1);
                const uint8_t* hr = half.data() + jr * half_pitch;
                hx[jr * hx_pitch + o] = (uint16_t)linear_h4(hr[(s0 - i_lo) * 3 + c], hr[(s1 - i_lo) * 3 + c], lx_a[x0 + xr]);
            }
            const long row0 = (long)y0 * dst_pitch + x0 * 3;
            const int chunks_max = (tw3 + 15 + 15) >> 4;
            for (int idx = 0; idx < th * chunks_max; ++idx) {
                const int r = idx / chunks_max, j = idx - r * chunks_max;
                uint8_t* drow = dst + row0 + (long)r * dst_pitch;
                const int shift = (int)((row0 + (long)r * dst_pitch + dst_phase) & 15);
                const int lo = 16 * j - shift;
                if (lo >= tw3) continue;
                const uint32_t ys = ly_s[y0 + r], yb = ly_b[y0 + r];
                const uint16_t* h0 = hx.data() + ((int)(ys & 0xFFFFu) - j_lo) * hx_pitch;
                const uint16_t* h1 = hx.data() + ((int)(ys >> 16) - j_lo) * hx_pitch;
                for (int b = 0; b < 16; ++b) {
                    const int i = lo + b;
                    if (i >= 0 && i < tw3) drow[i] = (uint8_t)linear_v(h0[i], h1[i], yb);
                }
            }
        }
    return 0;
}

// Replays noise_kernel (compat / philox / field) over one image's flat element range.
extern "C" int emu_noise(const uint8_t* src, uint8_t* dst, const float* noise, float* field_out, long n_elems,
                         float sigma, uint64_t seed, uint64_t image_index, uint32_t offset) {
    for (long g = 0; g < (n_elems + 3) / 4; ++g) {
        float nz[4] = {0, 0, 0, 0};
        if (noise == nullptr) {
            uint32_t r[4];
            float z[4];
            philox4x32_10((uint32_t)g, (uint32_t)image_index, (uint32_t)(image_index >> 32), offset, (uint32_t)seed,
                          (uint32_t)(seed >> 32), r);
            boxmuller4(r, z);
            for (int j = 0; j < 4; ++j) nz[j] = sigma * z[j];
        }
        for (int j = 0; j < 4; ++j) {
            long e = 4 * g + j;
            if (e >= n_elems) break;
            float nv = noise ? noise[e] : nz[j];
            if (field_out) field_out[e] = nv;
            if (dst) dst[e] = (uint8_t)noise_px((float)src[e], nv);
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
