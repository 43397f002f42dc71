// rod_core.h -- arithmetic of the corruption path, written once for device and host.
//
// Every function here is a pure per-element / per-chunk routine with no thread
// cooperation, so the very same code is (a) inlined into the sm_100a kernels and
// (b) compiled by g++ into tests/emu (a sequential harness that replays the kernels'
// tile loops on the CPU against the oracle, because the build container has no GPU).
// The harness is test infrastructure; the product never runs these on the host.
//
// Arithmetic specs: SURVEY.md section 8a (a1, a3, a4, a5), i.e. NumPy 2.3.5
// `astype/clip` semantics and OpenCV 4.13.0 filter2D / resize(INTER_AREA) /
// resize(INTER_LINEAR, 8U) as called from scripts/augmentations.py:30-45.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define ROD_HD __host__ __device__ __forceinline__
#else
#define ROD_HD inline
#endif

namespace rod {

// ---------------------------------------------------------------------------------
// Device-resident descriptors (built by plan.cu)
// ---------------------------------------------------------------------------------
struct DevImage {
    uint64_t src_off, dst_off;  // bytes from the src / dst base pointers
    int64_t src_pitch, dst_pitch;
    uint64_t elem_base;         // packed element index of this image's first byte (noise field offset)
    int32_t h, w;
    int32_t shape_id;           // index into the DevShape table (lowres / letterbox)
    int32_t contiguous;         // 1 when both pitches equal 3*w
};

// One work item of a kernel.  Meaning of a/b/c depends on the op:
//   noise : a = first element (flat index inside the image or row), b = element count, c = row (or -1: flat)
//   blur  : a = first row, b = row count
//   lowres: a = first output row, b = first output column (pixels)
struct Tile {
    int32_t img, a, b, c;
};

enum AreaMode { AREA_GENERAL = 0, AREA_FAST2 = 1, AREA_FASTN = 2, AREA_IDENTITY = 3 };

// Resize tables of one (src h, src w) -> (nh, nw) -> (h, w) lowres round trip, or of
// one letterbox geometry.  Offsets are in 32-bit words into the plan's table blob.
struct DevShape {
    int32_t h, w, nh, nw;       // full size, low-res size
    int32_t area_mode;
    int32_t ix, iy;             // integer scales (AREA_FASTN)
    int32_t xt, yt;             // taps per destination (AREA_GENERAL), row stride of the alpha tables
    uint32_t ax_first, ax_alpha;  // int32 first[nw]; float alpha[nw*xt] (unused taps are 0 with index clamped)
    uint32_t ay_first, ay_alpha;  // int32 first[nh]; float alpha[nh*yt]
    uint32_t ax_count, ay_count;  // int32 count[nw], count[nh]
    uint32_t lx_s0, lx_a;       // int32 s0[w]; uint32 (a0 | a1 << 16)[w]   (low-res -> full, x axis)
    uint32_t ly_s, ly_b;        // uint32 (s0 | s1 << 16)[h]; uint32 (b0 | b1 << 16)[h]
    float inv_area;             // float(1 / (ix*iy))
    int32_t lin_identity;       // 1 when (nh, nw) == (h, w): the INTER_LINEAR step is a copy
    int32_t x2;                 // 1 when w == 2 * nw: exact 2x on the x axis (closed-form coefficients)
    int32_t strip_rows;         // x2 kernel: output rows per strip (multiple of 8)
    int32_t strip_half_rows;    // x2 kernel: worst-case low-res rows one strip touches
    uint32_t ay_pack;           // x2 kernel, yt <= 3: uint4 per low-res row {first source row, beta0, beta1, beta2 bits}
    int32_t ay_packed;          // 1 when ay_pack is valid
    uint32_t ly_rc;             // float4 per output row: the vertical-stage constants (X2Row) of the exact-2x kernels
    int32_t x2w;                // 1: eligible for the warp-marching exact-2x kernel (lowres_x2w_kernel)
    int32_t x2p;                // 1: exact 2x in both axes, w % 4 == 0: the packed-integer kernel (lowres_x2p_kernel)
    uint32_t ly_rc2;            // float4 per output row {c0s, c1s, k0 + 2, bits of x2_vertical_cfix}: lowres_x2f_kernel
    uint32_t ly_rc3;            // float4 per output row: x2g_row_consts (general horizontal stage): lowres_x2g_kernel
    int32_t x2g;                // 1: odd width at factor 0.5 with the regular 3-tap / one-slip structure (lowres_x2g_kernel)
    uint32_t hy_pack;           // uint4 per low-res row {beta0, beta1, beta2, r0 | nA << 16 | nB << 24}: lowres_x2h_kernel
    int32_t x2h;                // 1: exact-2x width whose low-res row j has the three y taps 2j, 2j+1, 2j+2 (h = 2 nh + 1)
    int32_t x2i;                // odd width (x2g) with a regular y axis: 1 = taps 2j, 2j+1, 2j+2 (odd h), 2 = taps 2j, 2j+1 (even h)
};

// ---------------------------------------------------------------------------------
// float helpers: never contracted into FMA (OpenCV's resizeArea_ uses separate mul/add)
// ---------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
ROD_HD float fmul(float a, float b) { return __fmul_rn(a, b); }
ROD_HD float fadd(float a, float b) { return __fadd_rn(a, b); }
ROD_HD float frint(float a) { return rintf(a); }
ROD_HD float fadd_rz(float a, float b) { return __fadd_rz(a, b); }
ROD_HD uint32_t fbits(float a) { return __float_as_uint(a); }
ROD_HD float bitsf(uint32_t a) { return __uint_as_float(a); }
#else
ROD_HD float fmul(float a, float b) { volatile float r = a * b; return r; }
ROD_HD float fadd(float a, float b) { volatile float r = a + b; return r; }
ROD_HD float frint(float a) { return nearbyintf(a); }
// host stand-in for add.rz.f32, exact for the only use (0 <= a <= 255, b = 2^23)
ROD_HD float fadd_rz(float a, float b) { return (float)floor((double)a + (double)b); }
ROD_HD uint32_t fbits(float a) { uint32_t u; memcpy(&u, &a, 4); return u; }
ROD_HD float bitsf(uint32_t a) { float f; memcpy(&f, &a, 4); return f; }
#endif

// cv::borderInterpolate(BORDER_REFLECT_101), periodic form (valid for any reach, n >= 1).
ROD_HD int reflect101(int i, int n) {
    if (n == 1) return 0;
    int period = 2 * (n - 1);
    int m = i % period;
    if (m < 0) m += period;
    return m < n ? m : period - m;
}

// ---------------------------------------------------------------------------------
// a1: noise.  out = uint8(trunc(clamp(float(v) + noise, 0, 255)))  (augmentations.py:32-33)
// ---------------------------------------------------------------------------------
// vf is the pixel already converted to float (exact).
ROD_HD uint32_t noise_px(float vf, float nz) {
    float s = fadd(vf, nz);  // one RN add
    s = fmaxf(s, 0.0f);
    s = fminf(s, 255.0f);
    // truncation == numpy astype(uint8) on [0,255]: 2^23 + s rounded toward zero keeps floor(s) in the low byte
    return fbits(fadd_rz(s, 8388608.0f)) & 0xFFu;
}

// Philox4x32-10 (Salmon et al. 2011), counter c[4], key k[2].
ROD_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

ROD_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                          uint32_t r[4]) {
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        uint32_t hi0 = mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    r[0] = c0; r[1] = c1; r[2] = c2; r[3] = c3;
}

// Same function with the 10 round keys precomputed (rk[2i] = k0 + i * 0x9E3779B9, rk[2i+1] = k1 + i * 0xBB67AE85):
// in a kernel the keys are launch parameters, i.e. constant-bank operands of the XORs, and the per-round key
// additions disappear from the instruction stream.
struct PhiloxKeys {
    uint32_t rk[20];
};
inline PhiloxKeys philox_round_keys(uint32_t k0, uint32_t k1) {
    PhiloxKeys ks;
    for (int i = 0; i < 10; ++i) {
        ks.rk[2 * i] = k0 + (uint32_t)i * 0x9E3779B9u;
        ks.rk[2 * i + 1] = k1 + (uint32_t)i * 0xBB67AE85u;
    }
    return ks;
}
template <int ROUNDS>
ROD_HD void philox4x32_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKeys& ks, uint32_t r[4]) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < ROUNDS; ++i) {
        uint32_t hi0 = mulhi32(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = mulhi32(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ ks.rk[2 * i];
        uint32_t n2 = hi0 ^ c3 ^ ks.rk[2 * i + 1];
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    }
    r[0] = c0; r[1] = c1; r[2] = c2; r[3] = c3;
}
ROD_HD void philox4x32_10_rk(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKeys& ks, uint32_t r[4]) {
    philox4x32_rk<10>(c0, c1, c2, c3, ks, r);
}

// Philox-mode noise (rod_noise_u8 with noise == NULL).  One Philox block r[4] serves the GROUP of 8
// consecutive elements e = 8g .. 8g+7 of an image (counter = (g, image lo, image hi, offset), key = seed);
// word r[p] makes the Box-Muller pair (8g + 2p, 8g + 2p + 1):
//   radius uniform  u  = (hi16 + 0.5) * 2^-16, hi16 = r[p] >> 16      (stratified; reaches 4.86 sigma)
//     tail          if hi16 == 0 (probability 2^-16): u = (t[p] + 0.5) * 2^-48 with t = the Philox block
//                   at counter (g, image lo, image hi ^ 0x80000000, offset) -> the tail reaches 8.2 sigma
//   angle           th = 2 pi (lo16 + 0.5) / 65536, lo16 = r[p] & 0xffff
//   s[2p] = sqrt(-log2 u) cos th,  s[2p+1] = sqrt(-log2 u) sin th      (unit-free factors)
//   noise n = K s with K = sigma * sqrt(2 ln 2);   out = clamp(v + floor(n), 0, 255)
// (v is an integer, so v + floor(n) == floor(v + n): the reference's truncation of the clipped sum.)
// The fp32 constants below are part of the definition (oracle/corruption_oracle.py restates them).
#define ROD_PHILOX_TAIL_FLIP 0x80000000u
#define ROD_NOISE_K_PER_SIGMA 1.17741002251547466f  // sqrt(2 ln 2)
#define ROD_ANGLE_SCALE 9.58737992428525768e-05f    // float(2 pi / 65536)
#define ROD_ANGLE_BIAS (-804.247680664062500f)      // float((0.5 - 2^23) * 2 pi / 65536)
ROD_HD bool philox_needs_tail(const uint32_t r[4]) {
    uint32_t m = r[0] < r[1] ? r[0] : r[1];
    const uint32_t m2 = r[2] < r[3] ? r[2] : r[3];
    m = m < m2 ? m : m2;
    return m < 0x10000u;
}
// One Box-Muller pair from the word rw; tw is the tail word (used only when use_tail and rw < 2^16).
ROD_HD void gauss2(uint32_t rw, bool use_tail, uint32_t tw, float* s0, float* s1) {
#if defined(__CUDA_ARCH__)
    // 2^23 + hi16 and 2^23 + lo16 assembled by byte permutes (no I2F)
    const float xh = bitsf(__byte_perm(rw, 0x4B000000u, 0x7432));
    const float xl = bitsf(__byte_perm(rw, 0x4B000000u, 0x7410));
    float u = fmaf(xh, 1.52587890625e-05f, -127.99999237060546875f);  // (hi16 + 0.5) * 2^-16, exact
    if (use_tail && rw < 0x10000u) u = fmaf((float)tw, 3.5527136788005009e-15f, 1.7763568394002505e-15f);
    float l2, rad;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(u));
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad) : "f"(-l2));
    const float th = fmaf(xl, ROD_ANGLE_SCALE, ROD_ANGLE_BIAS);
    *s0 = rad * __cosf(th);
    *s1 = rad * __sinf(th);
#else
    double u = ((double)(rw >> 16) + 0.5) * 1.52587890625e-05;
    if (use_tail && rw < 0x10000u) u = ((double)tw + 0.5) * 3.5527136788005009e-15;
    const double rad = sqrt(-log2(u));
    const double th = (8388608.0 + (double)(rw & 0xFFFFu)) * (double)ROD_ANGLE_SCALE + (double)ROD_ANGLE_BIAS;
    *s0 = (float)(rad * cos(th));
    *s1 = (float)(rad * sin(th));
#endif
}
// t may be NULL when !philox_needs_tail(r).
ROD_HD void gauss8(const uint32_t r[4], const uint32_t* t, float s[8]) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int p = 0; p < 4; ++p) gauss2(r[p], t != nullptr, t != nullptr ? t[p] : 0u, &s[2 * p], &s[2 * p + 1]);
}
// int16 floor(K * s) in the low half of the result (two's complement), valid for |K s| < 32768:
// fma rounded toward -inf onto the integer grid of [2^23, 2^24) (1.5 * 2^23 keeps negatives in the binade).
ROD_HD uint32_t noise_floor16(float s, float K) {
#if defined(__CUDA_ARCH__)
    return fbits(__fmaf_rd(s, K, 12582912.0f));
#else
    return (uint32_t)(int32_t)floor((double)s * (double)K) & 0xFFFFu;
#endif
}
#if defined(__CUDACC__)
// four pixels (one word) + four factors -> four output bytes: clamp(v + floor(K s), 0, 255)
__device__ __forceinline__ uint32_t philox_word(uint32_t word, const float* s, float K) {
    const uint32_t f01 = __byte_perm(noise_floor16(s[0], K), noise_floor16(s[1], K), 0x5410);
    const uint32_t f23 = __byte_perm(noise_floor16(s[2], K), noise_floor16(s[3], K), 0x5410);
    const uint32_t v01 = __byte_perm(word, 0u, 0x4140), v23 = __byte_perm(word, 0u, 0x4342);
    const uint32_t q01 = __viaddmin_s16x2_relu(f01, v01, 0x00FF00FFu);
    const uint32_t q23 = __viaddmin_s16x2_relu(f23, v23, 0x00FF00FFu);
    return __byte_perm(q01, q23, 0x6420);
}
#endif
ROD_HD uint32_t noise_philox_px(uint32_t v, float s, float K) {
    int x = (int)v + (int)(int16_t)(noise_floor16(s, K) & 0xFFFFu);
    return (uint32_t)(x < 0 ? 0 : (x > 255 ? 255 : x));
}

// Philox-mode noise, TABLE generator (noise.cu noise_table_kernel; used when ROD_GAUSS_TABLE_MIN_SIGMA <= sigma <=
// ROD_GAUSS_TABLE_MAX_SIGMA).  Integer arithmetic only, ONE Philox block per GROUP OF 16 elements:
//   X[i], i = 0..255 : a 256-point equal-probability discretisation of N(0, (sigma/2)^2) in 1/256 pixel units
//                      (rod_tables.h build_gauss_table: cell means of the 256 quantile cells, tails stretched so that
//                      the 2nd, 4th and 6th moments are the Gaussian ones, rounded to integers)
//   group g = e >> 4, block r[4] = Philox4x32-10(counter (g, image lo, image hi, offset), key seed)
//   word r[q] -> four independent draws  xa = X[byte 0], xb = X[byte 1], xc = X[byte 2], xd = X[byte 3]  and their
//   4 x 4 Hadamard transform (an orthogonal mix: four uncorrelated sums, each the sum of four N(0, sigma^2/4) draws)
//     element 16g + 4q + 0 : k = floor((xa + xb + xc + xd) / 256)
//     element 16g + 4q + 1 : k = floor((xa - xb + xc - xd) / 256)
//     element 16g + 4q + 2 : k = floor((xa + xb - xc - xd) / 256)
//     element 16g + 4q + 3 : k = floor((xa - xb - xc + xd) / 256)
//   out = clamp(v + k, 0, 255)   (k is floor(noise), so this is the reference's truncation of the clipped sum).
// The sum of four 256-level draws has 2^32 equally likely outcomes on the 1/256 grid and tails to 6.2 sigma; its
// binned distribution is within 1e-3 relative of the Gaussian cell probabilities down to cells of 3e-6 (exact 4-fold
// convolution, tests/test_emulation.py).  On the device the table lives in shared memory as two 32-bit forms per entry,
// (x, x) and (x, -x) as 16-bit halves of one 32-bit integer, replicated per lane (bank = lane: conflict-free); the
// Hadamard transform of a word is then four 32-bit integer additions that produce two elements each.
#define ROD_GAUSS_TABLE_MAX_SIGMA 20.0f  // 4 * max |X| = 4 * 128 sigma * 3.0958 must stay below 32768
#define ROD_GAUSS_TABLE_MIN_SIGMA 3.0f   // below, the 256-level draws are too coarse against 1-pixel cells (Box-Muller)
#define ROD_GAUSS_H4_STRETCH_A (-1.4657745851475e-3)  // y = z (1 + A z^4 + B z^8): matches the 4th and 6th moments
#define ROD_GAUSS_H4_STRETCH_B (2.4950155916569e-5)
#ifndef ROD_GAUSS_AUTO
#define ROD_GAUSS_AUTO 0                 // table generator when MIN <= sigma <= MAX, else Box-Muller
#define ROD_GAUSS_BOXMULLER 1
#define ROD_GAUSS_TABLE_PHILOX7 2        // like AUTO, but the table generator runs on Philox4x32-7 (Random123's fastest
                                         // Crush-resistant round count) instead of Philox4x32-10
#endif
// The two shared-memory forms of table entry x (as 32-bit integers; the 16-bit halves are x and +-x modulo carries
// that cancel in the final sums because 32-bit addition is linear).
ROD_HD uint32_t h4_form_same(int32_t x) { return (uint32_t)x * 0x00010001u; }
ROD_HD uint32_t h4_form_diff(int32_t x) { return (uint32_t)x * 0xFFFF0001u; }
// A, C: (x, x) forms of draws a, c;  B, D: (x, -x) forms of draws b, d.  e01 / e23: elements (0, 1) / (2, 3) of the word as
// 16-bit halves 256 k + frac + 32768, i.e. byte 1 (byte 3) of each is k + 128.
ROD_HD void h4_combine(uint32_t A, uint32_t B, uint32_t C, uint32_t D, uint32_t* e01, uint32_t* e23) {
    const uint32_t u = A + B + 0x80008000u, v = C + D;
    *e01 = u + v;
    *e23 = u - v;
}
// element j (0..3) of Philox word w: k = floor(noise)
ROD_HD int h4_word_k(uint32_t w, const int32_t* X, int j) {
    uint32_t e01, e23;
    h4_combine(h4_form_same(X[w & 0xFFu]), h4_form_diff(X[(w >> 8) & 0xFFu]), h4_form_same(X[(w >> 16) & 0xFFu]),
               h4_form_diff(X[w >> 24]), &e01, &e23);
    const uint32_t e = (j & 2) ? e23 : e01;
    return (int)(((j & 1) ? (e >> 24) : (e >> 8)) & 0xFFu) - 128;
}
ROD_HD uint32_t noise_table_px(uint32_t v, int k) {
    const int x = (int)v + k;
    return (uint32_t)(x < 0 ? 0 : (x > 255 ? 255 : x));
}

// Phi^-1 in double (Acklam's rational start + Halley steps on erfc): host only (table builder, emu harness).
inline double ndtri_double(double p) {
    if (p > 0.5) return -ndtri_double(1.0 - p);
    static const double a[] = {-3.969683028665376e+01, 2.209460984245205e+02, -2.759285104469687e+02,
                               1.383577518672690e+02, -3.066479806614716e+01, 2.506628277459239e+00};
    static const double b[] = {-5.447609879822406e+01, 1.615858368580409e+02, -1.556989798598866e+02,
                               6.680131188771972e+01, -1.328068155288572e+01};
    static const double c[] = {-7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e+00,
                               -2.549732539343734e+00, 4.374664141464968e+00, 2.938163982698783e+00};
    static const double d[] = {7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e+00,
                               3.754408661907416e+00};
    double x;
    if (p < 0.02425) {
        const double q = sqrt(-2.0 * log(p));
        x = (((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) /
            ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1.0);
    } else {
        const double q = p - 0.5, r = q * q;
        x = (((((a[0] * r + a[1]) * r + a[2]) * r + a[3]) * r + a[4]) * r + a[5]) * q /
            (((((b[0] * r + b[1]) * r + b[2]) * r + b[3]) * r + b[4]) * r + 1.0);
    }
    for (int it = 0; it < 3; ++it) {
        const double e = 0.5 * erfc(-x * 0.70710678118654752440) - p;
        const double u = e * 2.50662827463100050242 * exp(0.5 * x * x);
        x -= u / (1.0 + 0.5 * x * u);
    }
    return x;
}

// ---------------------------------------------------------------------------------
// a3: horizontal k-tap box blur on an interleaved 3-channel row that has been staged
// (with its reflected halo) in a byte buffer.  `row` points at the byte of pixel 0.
// ---------------------------------------------------------------------------------
// Generic odd k (slow path): one output byte.
ROD_HD uint32_t blur_byte_generic(const uint8_t* row, int i, int k) {
    int r = k >> 1;
    int s = 0;
    for (int j = -r; j <= r; ++j) s += row[i + 3 * j];
    return (uint32_t)((2 * s + k) / (2 * k));
}

// Byte `p` (0..47) of a 12-word window.
ROD_HD uint32_t win_byte(const uint32_t* w, int p) { return (w[p >> 2] >> (8 * (p & 3))) & 0xFFu; }

// k = 9 fast path: 16 consecutive outputs from a 48-byte window w[0..11] that holds
// input bytes [-16, 32) relative to the first output byte.  Uses 3 interleaved running
// prefix sums (one per channel phase): C[p] = C[p-3] + B[p];  S[n] = C[n+12] - C[n-15].
// Rounded division without an integer multiply: with x = 2^23 + S (built by the same
// 3-input add that forms S) and c9 = 932068 * 2^-23 (1/9 * (1 + 4.8e-7), chosen so that
// 2^23 * c9 is an integer), fma(x, c9, 2^23 - 932068) = 2^23 + S*c9 exactly, rounded once
// to the integer grid: the low byte is round(S / 9) because S/9 is never within 1/18 of a
// tie and S * 4.8e-7 / 9 < 1.3e-4.  (== (2S + 9) / 18 == cv2.filter2D's saturate_cast.)
ROD_HD void blur9_chunk16(const uint32_t* w, uint32_t out[4]) {
    // window byte index q = p + 16, p in [-15, 27]
    uint32_t C[43];  // C[t] for p = t - 15
#pragma unroll
    for (int t = 0; t < 43; ++t) {
        const int q = t + 1;  // p + 16
        if (t < 3) { C[t] = 0u; continue; }  // chain heads p = -15,-14,-13 are outside every window
#if defined(__CUDA_ARCH__)
        C[t] = __dp4a(w[q >> 2], 1u << (8 * (q & 3)), C[t - 3]);
#else
        C[t] = C[t - 3] + win_byte(w, q);
#endif
    }
    const float c9 = 0.111111164093017578125f;  // 932068 / 2^23
    uint32_t qv[16];
#pragma unroll
    for (int n = 0; n < 16; ++n) {
        const uint32_t xs = C[n + 27] - C[n] + 0x4B000000u;  // float bits of 2^23 + S
        qv[n] = fbits(fmaf(bitsf(xs), c9, 7456540.0f));       // low byte = round(S / 9)
    }
#pragma unroll
    for (int g = 0; g < 4; ++g) {
#if defined(__CUDA_ARCH__)
        const uint32_t lo = __byte_perm(qv[4 * g], qv[4 * g + 1], 0x0040);
        const uint32_t hi = __byte_perm(qv[4 * g + 2], qv[4 * g + 3], 0x0040);
        out[g] = __byte_perm(lo, hi, 0x5410);
#else
        out[g] = (qv[4 * g] & 0xFFu) | ((qv[4 * g + 1] & 0xFFu) << 8) | ((qv[4 * g + 2] & 0xFFu) << 16) |
                 ((qv[4 * g + 3] & 0xFFu) << 24);
#endif
    }
}

// ---------------------------------------------------------------------------------
// a4: INTER_AREA, one low-res value (channel c of low-res pixel (dy, dx)).
// src points at pixel (0,0) of the full-res image, pitch in bytes.
// ---------------------------------------------------------------------------------
ROD_HD uint32_t area_value(const uint8_t* src, int64_t pitch, const DevShape& sh, const uint32_t* tab, int dy,
                           int dx, int c) {
    if (sh.area_mode == AREA_FAST2) {
        const uint8_t* p = src + (int64_t)(2 * dy) * pitch + (2 * dx) * 3 + c;
        return ((uint32_t)p[0] + p[3] + p[pitch] + p[pitch + 3] + 2u) >> 2;
    }
    if (sh.area_mode == AREA_IDENTITY) {
        return src[(int64_t)dy * pitch + dx * 3 + c];
    }
    if (sh.area_mode == AREA_FASTN) {
        uint32_t s = 0;
        for (int yy = 0; yy < sh.iy; ++yy) {
            const uint8_t* p = src + (int64_t)(dy * sh.iy + yy) * pitch + (dx * sh.ix) * 3 + c;
            for (int xx = 0; xx < sh.ix; ++xx) s += p[3 * xx];
        }
        float r = frint(fmul((float)s, sh.inv_area));
        r = r < 0.f ? 0.f : (r > 255.f ? 255.f : r);
        return (uint32_t)(int)r;
    }
    const int32_t* xfirst = (const int32_t*)(tab + sh.ax_first);
    const int32_t* xcount = (const int32_t*)(tab + sh.ax_count);
    const float* xalpha = (const float*)(tab + sh.ax_alpha) + dx * sh.xt;
    const int32_t* yfirst = (const int32_t*)(tab + sh.ay_first);
    const int32_t* ycount = (const int32_t*)(tab + sh.ay_count);
    const float* yalpha = (const float*)(tab + sh.ay_alpha) + dy * sh.yt;
    int sx0 = xfirst[dx], nx = xcount[dx];
    int sy0 = yfirst[dy], ny = ycount[dy];
    float sum = 0.f;
    for (int ty = 0; ty < ny; ++ty) {
        const uint8_t* p = src + (int64_t)(sy0 + ty) * pitch + sx0 * 3 + c;
        float buf = 0.f;
        for (int tx = 0; tx < nx; ++tx) buf = fadd(buf, fmul((float)p[3 * tx], xalpha[tx]));
        float prod = fmul(yalpha[ty], buf);
        sum = (ty == 0) ? prod : fadd(sum, prod);
    }
    float r = frint(sum);
    r = r < 0.f ? 0.f : (r > 255.f ? 255.f : r);
    return (uint32_t)(int)r;
}

// ---------------------------------------------------------------------------------
// a4 fast paths for exact-2x widths: one "unit" = 12 consecutive source bytes (4 BGR pixels)
// of one row -> the six horizontal pixel-pair sums (bytes k and k+3).  s[k] = acc[k] + pair_k.
// ---------------------------------------------------------------------------------
ROD_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, int sh) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_r(lo, hi, sh);
#else
    return (uint32_t)((((uint64_t)hi << 32) | lo) >> sh);
#endif
}
ROD_HD uint32_t sum_b0_b3(uint32_t w, uint32_t acc) {
#if defined(__CUDA_ARCH__)
    return __dp4a(w, 0x01000001u, acc);
#else
    return acc + (w & 0xFFu) + (w >> 24);
#endif
}
ROD_HD void pair_sums12(uint32_t w0, uint32_t w1, uint32_t w2, const uint32_t acc[6], uint32_t s[6]) {
    s[0] = sum_b0_b3(w0, acc[0]);                   // b0 + b3
    s[1] = sum_b0_b3(funnel_r(w0, w1, 8), acc[1]);  // b1 + b4
    s[2] = sum_b0_b3(funnel_r(w0, w1, 16), acc[2]); // b2 + b5
    s[3] = sum_b0_b3(funnel_r(w1, w2, 16), acc[3]); // b6 + b9
    s[4] = sum_b0_b3(funnel_r(w1, w2, 24), acc[4]); // b7 + b10
    s[5] = sum_b0_b3(w2, acc[5]);                   // b8 + b11
}
// resizeAreaFast_ 2x2: (a + b + c + d + 2) >> 2 for six low-res bytes from two source rows.
ROD_HD void area_fast2_unit(const uint32_t r0[3], const uint32_t r1[3], uint32_t out6[6]) {
    const uint32_t two[6] = {2u, 2u, 2u, 2u, 2u, 2u};
    uint32_t s[6];
    pair_sums12(r0[0], r0[1], r0[2], two, s);
    pair_sums12(r1[0], r1[1], r1[2], s, s);
#pragma unroll
    for (int q = 0; q < 6; ++q) out6[q] = s[q] >> 2;
}
// resizeArea_ with x scale exactly 2 (both x taps are 0.5) and `ny` float y taps:
//   buf = (S0 + S1) * 0.5 (exact);  sum = beta0*buf0 (+ beta_t*buf_t ...), each op rounded;  out = rint(sum).
// The factor 0.5 commutes with every rounding (power of two), so it is applied once inside the
// final fma, which also performs the round-half-even to the integer grid (2^23 * 1.5 magic).
ROD_HD void area_x2f_accumulate(const uint32_t r[3], float beta, bool first, float acc[6]) {
    const uint32_t magic[6] = {0x4B000000u, 0x4B000000u, 0x4B000000u, 0x4B000000u, 0x4B000000u, 0x4B000000u};
    uint32_t s[6];
    pair_sums12(r[0], r[1], r[2], magic, s);  // float bits of 2^23 + (S0 + S1)
    // beta * (x - 2^23) == fma(x, beta, -beta * 2^23): the fma is exact before its single rounding and
    // beta * 2^23 is exact, so this is fl(beta * S) -- the same value as the separate convert + multiply.
    const float nb = fmul(beta, -8388608.0f);
#pragma unroll
    for (int q = 0; q < 6; ++q) {
        const float prod = fmaf(bitsf(s[q]), beta, nb);
        acc[q] = first ? prod : fadd(acc[q], prod);
    }
}
ROD_HD void area_x2f_finish(const float acc[6], uint32_t out6[6]) {
#pragma unroll
    for (int q = 0; q < 6; ++q) out6[q] = fbits(fmaf(acc[q], 0.5f, 12582912.0f));  // result in the low byte
}

// ---------------------------------------------------------------------------------
// a5: INTER_LINEAR 8U.  Horizontal stage: (S[s0]*a0 + S[s1]*a1) >> 4 (fits 16 bits).
// ---------------------------------------------------------------------------------
ROD_HD uint32_t linear_h4(uint32_t s0v, uint32_t s1v, uint32_t a_packed) {
    return (s0v * (a_packed & 0xFFFFu) + s1v * (a_packed >> 16)) >> 4;
}
// Vertical stage + pack: (((b0*h0) >> 16) + ((b1*h1) >> 16) + 2) >> 2.
ROD_HD uint32_t linear_v(uint32_t h0, uint32_t h1, uint32_t b_packed) {
    return ((((b_packed & 0xFFFFu) * h0) >> 16) + (((b_packed >> 16) * h1) >> 16) + 2u) >> 2;
}

// ---------------------------------------------------------------------------------
// a5 fast path for exact-2x widths (w == 2 * nw).  OpenCV's x coefficients are then
// (2048,0) at x = 0 and x = w-1 and otherwise (1536,512) for odd x = 2i+1 (taps i, i+1) and
// (512,1536) for even x = 2i+2 (taps i, i+1), so the horizontal stage is hx = 32 * q with
// q = 3A + B or A + 3B, and ((b * hx) >> 16) == floor(b * q / 2048).
//
// One "chunk" = 8 output pixels x = 8m .. 8m+7 (24 bytes) needs the six low-res pixels
// p0..p5 = P[4m-1 .. 4m+4] (18 bytes; P[-1] := P[0], P[nw] := P[nw-1] are replicated by the
// producer).  win[] holds those 18 bytes starting at byte 0 of win[0].
// x[t] = float(2^22 + q[t]) for the 24 output bytes t = 3 * pixel + channel (the float's mantissa step is
// 0.5 there, so the dot products use doubled weights and the bit pattern 0x4A800000 + 2q).
// ---------------------------------------------------------------------------------
ROD_HD uint32_t dot2_lo(uint32_t w16x2, uint32_t bytes, uint32_t acc) {
#if defined(__CUDA_ARCH__)
    return __dp2a_lo(w16x2, bytes, acc);
#else
    return acc + (w16x2 & 0xFFFFu) * (bytes & 0xFFu) + (w16x2 >> 16) * ((bytes >> 8) & 0xFFu);
#endif
}
ROD_HD uint32_t dot2_hi(uint32_t w16x2, uint32_t bytes, uint32_t acc) {
#if defined(__CUDA_ARCH__)
    return __dp2a_hi(w16x2, bytes, acc);
#else
    return acc + (w16x2 & 0xFFFFu) * ((bytes >> 16) & 0xFFu) + (w16x2 >> 16) * (bytes >> 24);
#endif
}
// bytes (a, a+3, a+1, a+4) of the window, a = 3 * pair + c: [A_c, B_c, A_c+1, B_c+1]
ROD_HD uint32_t x2_gather(const uint32_t* win, int a) {
#if defined(__CUDA_ARCH__)
    const int k = a >> 2, o = a & 3;  // all four bytes lie in win[k], win[k+1]
    return __byte_perm(win[k], win[k + 1], (o) | ((o + 3) << 4) | ((o + 1) << 8) | ((o + 4) << 12));
#else
    const uint8_t* b = (const uint8_t*)win;
    return (uint32_t)b[a] | ((uint32_t)b[a + 3] << 8) | ((uint32_t)b[a + 1] << 16) | ((uint32_t)b[a + 4] << 24);
#endif
}
ROD_HD void x2_expand24(const uint32_t win[5], float x[24]) {
    const uint32_t W31 = 6u | (2u << 16), W13 = 2u | (6u << 16), M = 0x4A800000u;
    // pair m = (p_m, p_m+1); channels handled two at a time: (0,1) from one gather, (2, next pair's 0) from another
    // outputs: pair0 -> px0 (1,3); pair1 -> px1 (3,1), px2 (1,3); pair2 -> px3, px4; pair3 -> px5, px6; pair4 -> px7 (3,1)
#pragma unroll
    for (int m = 0; m < 5; ++m) {
        const uint32_t g01 = x2_gather(win, 3 * m);      // channels 0,1 of pair m
        const uint32_t g2x = x2_gather(win, 3 * m + 2);  // channel 2 of pair m (and channel 0 of pair m+1, unused)
        const int podd = 2 * m - 1, peven = 2 * m;       // output pixels fed by pair m
        if (podd >= 0) {
            x[3 * podd + 0] = bitsf(dot2_lo(W31, g01, M));
            x[3 * podd + 1] = bitsf(dot2_hi(W31, g01, M));
            x[3 * podd + 2] = bitsf(dot2_lo(W31, g2x, M));
        }
        if (peven < 8) {
            x[3 * peven + 0] = bitsf(dot2_lo(W13, g01, M));
            x[3 * peven + 1] = bitsf(dot2_hi(W13, g01, M));
            x[3 * peven + 2] = bitsf(dot2_lo(W13, g2x, M));
        }
    }
}

// Per-output-row constants of the float vertical stage (b0, b1 = 11-bit y coefficients, <= 2048).
struct X2Row {
    float c0s, c1s;   // b0 * 2^-11, b1 * 2^-11
    float k0;         // 2^23 - b0 * 2048
    float k2;         // 6291456.5 - b1 * 512
};
ROD_HD X2Row x2_row_consts(uint32_t b_packed) {
    X2Row r;
    const float b0 = (float)(b_packed & 0xFFFFu), b1 = (float)(b_packed >> 16);
    r.c0s = b0 * 4.8828125e-4f;
    r.c1s = b1 * 4.8828125e-4f;
    r.k0 = 8388608.0f - b0 * 2048.0f;
    r.k2 = 6291456.5f - b1 * 512.0f;
    return r;
}
// out = (floor(b0*q0/2048) + floor(b1*q1/2048) + 2) >> 2 with x = 2^22 + q, in three fp32 fused
// multiply-adds rounded toward zero; every intermediate is an exact integer in [2^23, 2^24):
//   y1 = x0*c0s + k0 = 2^23 + F0                      (x0*c0s = b0*2048 + q0*b0/2048)
//   y2 = x1*c1s + y1 = 2^23 + b1*2048 + F0 + F1       (< 2^24 because b1 <= 2048)
//   o  = y2/4 + k2   = 2^23 + ((F0 + F1 + 2) >> 2)    (k2 removes b1*512 and adds 0.5)
ROD_HD uint32_t x2_vertical(float x0, float x1, const X2Row& r) {
#if defined(__CUDA_ARCH__)
    const float y1 = __fmaf_rz(x0, r.c0s, r.k0);
    const float y2 = __fmaf_rz(x1, r.c1s, y1);
    return __float_as_uint(__fmaf_rz(y2, 0.25f, r.k2));  // low byte is the result
#else
    const double y1 = floor((double)x0 * r.c0s + r.k0);
    const double y2 = floor((double)x1 * r.c1s + y1);
    return fbits((float)floor(y2 * 0.25 + r.k2));
#endif
}

// ---------------------------------------------------------------------------------
// a4 + a5 for shapes that are exact 2x in BOTH axes (w = 2 nw, h = 2 nh, i.e. even w and h at factor 0.5): an
// all-integer pipeline on packed 16-bit halves, two values per 32-bit word (lowres.cu lowres_x2p_kernel).
//   INTER_AREA is resizeAreaFast_:  P = (s00 + s01 + s10 + s11 + 2) >> 2.
//   INTER_LINEAR x: with q = 3 N + F (N / F = the nearer / farther low-res pixel of the output pixel; at the image
//     border F := N, OpenCV's (2048, 0) coefficients) the horizontal stage is hx = 512 q and (hx >> 4) = 32 q.
//   INTER_LINEAR y: the two coefficients are exactly (1536, 512) with the heavier one on the nearer low-res row, so
//     out = (((1536 * 32 qn) >> 16) + ((512 * 32 qf) >> 16) + 2) >> 2 = (floor(3 qn / 4) + floor(qf / 4) + 2) >> 2.
//   Every low-res row therefore publishes, per output byte column, A = 64 floor(q / 4) and B' = 64 (floor(3 q / 4) + 2)
//   (each < 2^16), and an output row is  byte 1 of (B'[near row] + A[far row])  -- one packed add per two bytes.
// A lane owns a chunk of 8 output pixels = low-res pixels P0..P3 (+ the halo pixels P-1, P4 of its neighbours); packed
// word t of a row holds the values of output bytes (2t, 2t+1) of the chunk, t = 0..11.
// ---------------------------------------------------------------------------------
template <int I0, int I1>
ROD_HD uint32_t pk2(uint32_t w) {  // (byte I0 of w) | (byte I1 of w) << 16
#if defined(__CUDA_ARCH__)
    return __byte_perm(w, 0u, I0 | (4 << 4) | (I1 << 8) | (4 << 12));
#else
    return ((w >> (8 * I0)) & 0xFFu) | (((w >> (8 * I1)) & 0xFFu) << 16);
#endif
}
template <int SEL>
ROD_HD uint32_t perm(uint32_t a, uint32_t b) {  // prmt.b32 in its default mode, selector nibbles 0..7
#if defined(__CUDA_ARCH__)
    return __byte_perm(a, b, SEL);
#else
    const uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) r |= (uint32_t)((v >> (8 * ((SEL >> (4 * i)) & 7))) & 0xFFu) << (8 * i);
    return r;
#endif
}
// The lane's twelve low-res values of one low-res row from its 24 source bytes of the two source rows, as raw sums
// s00 + s01 + s10 + s11 + 2 in six packed words, then scaled to 16 P = ((sum) & 0x3FC) << 2.  Value (pixel p, channel c)
// sits in: (0,0) b4.lo  (0,1) b0.lo  (0,2) b0.hi  (1,0) b1.lo  (1,1) b1.hi  (1,2) b4.hi
//          (2,0) b5.lo  (2,1) b2.lo  (2,2) b2.hi  (3,0) b3.lo  (3,1) b3.hi  (3,2) b5.hi
// (pairs chosen so that the two left bytes, and the two right bytes, of a word's values share one source word).
ROD_HD void x2p_area(const uint32_t r0[6], const uint32_t r1[6], uint32_t b[6]) {
    const uint32_t two = 0x00020002u;
    b[0] = pk2<1, 2>(r0[0]) + pk2<0, 1>(r0[1]) + (pk2<1, 2>(r1[0]) + pk2<0, 1>(r1[1]) + two);
    b[1] = pk2<2, 3>(r0[1]) + pk2<1, 2>(r0[2]) + (pk2<2, 3>(r1[1]) + pk2<1, 2>(r1[2]) + two);
    b[2] = pk2<1, 2>(r0[3]) + pk2<0, 1>(r0[4]) + (pk2<1, 2>(r1[3]) + pk2<0, 1>(r1[4]) + two);
    b[3] = pk2<2, 3>(r0[4]) + pk2<1, 2>(r0[5]) + (pk2<2, 3>(r1[4]) + pk2<1, 2>(r1[5]) + two);
    b[4] = perm<0x5410>(sum_b0_b3(r1[0], sum_b0_b3(r0[0], 2u)), sum_b0_b3(r1[2], sum_b0_b3(r0[2], 2u)));
    b[5] = perm<0x5410>(sum_b0_b3(r1[3], sum_b0_b3(r0[3], 2u)), sum_b0_b3(r1[5], sum_b0_b3(r0[5], 2u)));
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 6; ++i) b[i] = (b[i] & 0x03FC03FCu) << 2;
}
// A chunk with two low-res pixels only (the last chunk of a row when w % 8 == 4): P2 := P1 (OpenCV's clamped right tap).
ROD_HD void x2p_patch_two_pixel(uint32_t b[6]) {
    b[2] = perm<0x7632>(b[1], b[4]);  // (P2.c1, P2.c2) := (P1.c1, P1.c2)
    b[5] = b[1];                      // P2.c0 := P1.c0 (low half)
}
ROD_HD uint32_t x2p_n10(const uint32_t b[6]) { return perm<0x5432>(b[5], b[3]); }  // (P3.c2, P3.c0)
// left_n9 / left_n10: the left neighbour lane's b[3] = (P3.c0, P3.c1) and x2p_n10(); right_b4 / right_b0: the right
// neighbour's b[4] (low half P0.c0) and b[0] = (P0.c1, P0.c2).  first / last: the chunk touches the image border.
ROD_HD void x2p_build(const uint32_t b[6], uint32_t left_n9, uint32_t left_n10, uint32_t right_b4, uint32_t right_b0,
                      bool first, bool last, uint32_t A[12], uint32_t Bp[12]) {
    uint32_t N[12], F[12];
    N[0] = perm<0x5410>(b[4], b[0]);   // (P0.c0, P0.c1)
    N[1] = perm<0x5432>(b[0], b[4]);   // (P0.c2, P0.c0)
    N[2] = b[0];                       // (P0.c1, P0.c2)
    N[3] = b[1];                       // (P1.c0, P1.c1)
    N[4] = perm<0x5432>(b[4], b[1]);   // (P1.c2, P1.c0)
    N[5] = perm<0x7632>(b[1], b[4]);   // (P1.c1, P1.c2)
    N[6] = perm<0x5410>(b[5], b[2]);   // (P2.c0, P2.c1)
    N[7] = perm<0x5432>(b[2], b[5]);   // (P2.c2, P2.c0)
    N[8] = b[2];                       // (P2.c1, P2.c2)
    N[9] = b[3];                       // (P3.c0, P3.c1)
    N[10] = x2p_n10(b);                // (P3.c2, P3.c0)
    N[11] = perm<0x7632>(b[3], b[5]);  // (P3.c1, P3.c2)
    const uint32_t lc01 = first ? N[0] : left_n9;    // (P-1.c0, P-1.c1)
    const uint32_t lc2 = first ? N[1] : left_n10;    // low half: P-1.c2
    const uint32_t rc0 = last ? b[3] : right_b4;     // low half: P4.c0
    const uint32_t rc12 = last ? N[11] : right_b0;   // (P4.c1, P4.c2)
    F[0] = lc01;
    F[1] = perm<0x5410>(lc2, b[1]);    // (P-1.c2, P1.c0)
    F[2] = N[5];
    F[3] = N[0];
    F[4] = perm<0x5432>(b[0], b[5]);   // (P0.c2, P2.c0)
    F[5] = N[8];
    F[6] = N[3];
    F[7] = perm<0x5432>(b[4], b[3]);   // (P1.c2, P3.c0)
    F[8] = N[11];
    F[9] = N[6];
    F[10] = perm<0x5432>(b[2], rc0);   // (P2.c2, P4.c0)
    F[11] = rc12;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int t = 0; t < 12; ++t) {
        const uint32_t q16 = N[t] * 3u + F[t];               // 16 q per half, q = 3 N + F <= 1020
        A[t] = q16 & 0xFFC0FFC0u;                            // 64 floor(q / 4)
        Bp[t] = (q16 * 3u + 0x00800080u) & 0xFFC0FFC0u;      // 64 (floor(3 q / 4) + 2)
    }
}
// One output row of the chunk: 24 bytes in six words.
ROD_HD void x2p_emit(const uint32_t near_bp[12], const uint32_t far_a[12], uint32_t w[6]) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int g = 0; g < 6; ++g)
        w[g] = perm<0x7531>(near_bp[2 * g] + far_a[2 * g], near_bp[2 * g + 1] + far_a[2 * g + 1]);
}

// Two outputs of the float vertical stage at once (lowres_x2f_kernel): the same two round-toward-zero multiply-adds
// per byte as x2_vertical, y2 = 2^23 + b1 * 2048 + F0 + F1 + 2 (k0p = X2Row.k0 + 2), but the final (.. + 2) >> 2 is done
// for a PAIR in integer arithmetic: the low 16 bits of y2's bit pattern are Dl + v with Dl = (b1 & 31) * 2048 and
// v = F0 + F1 + 2 <= 1022 (no wrap: Dl + v < 2^16), so (packed pair) * 64 + cfix, cfix = -64 * Dl * 0x10001 (mod 2^32),
// leaves 64 v in each half and byte 1 / byte 3 of the result are v >> 2 -- 2 FFMA + 1 PRMT/2 + 1 IMAD/2 per byte
// instead of 3 FFMA, and the bytes are already in place for the final byte permute.
ROD_HD uint32_t x2_vertical_cfix(uint32_t b_packed) { return (0u - (((b_packed >> 16) & 31u) << 17)) * 0x00010001u; }
ROD_HD uint32_t x2_vertical_pair(float x0a, float x1a, float x0b, float x1b, float c0s, float c1s, float k0p, uint32_t cfix) {
#if defined(__CUDA_ARCH__)
    const float ya = __fmaf_rz(x1a, c1s, __fmaf_rz(x0a, c0s, k0p));
    const float yb = __fmaf_rz(x1b, c1s, __fmaf_rz(x0b, c0s, k0p));
    return __byte_perm(__float_as_uint(ya), __float_as_uint(yb), 0x5410) * 64u + cfix;
#else
    const double ya = floor((double)x1a * c1s + floor((double)x0a * c0s + k0p));
    const double yb = floor((double)x1b * c1s + floor((double)x0b * c0s + k0p));
    return ((fbits((float)ya) & 0xFFFFu) | (fbits((float)yb) << 16)) * 64u + cfix;
#endif
}
// Pair sums of one source row for the float INTER_AREA y taps: s[k] = float bits of 2^23 + (byte k + byte k+3) for the
// lane's twelve low-res byte columns (two 12-byte units), and the tap product / accumulation on them
// (see area_x2f_accumulate: fma(x, beta, -beta * 2^23) == fl(beta * S) exactly).
ROD_HD void x2f_pairsums(const uint32_t rw[6], uint32_t s[12]) {
    const uint32_t magic[6] = {0x4B000000u, 0x4B000000u, 0x4B000000u, 0x4B000000u, 0x4B000000u, 0x4B000000u};
    pair_sums12(rw[0], rw[1], rw[2], magic, s);
    pair_sums12(rw[3], rw[4], rw[5], magic, s + 6);
}
ROD_HD void x2f_mac(const uint32_t s[12], float beta, bool first, float acc[12]) {
    const float nb = fmul(beta, -8388608.0f);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int q = 0; q < 12; ++q) {
        const float prod = fmaf(bitsf(s[q]), beta, nb);
        acc[q] = first ? prod : fadd(acc[q], prod);
    }
}

// ---------------------------------------------------------------------------------
// a4 + a5 for ODD widths at factor 0.5 (w = 2 nw + 1: lowres_x2g_kernel).  OpenCV's general INTER_AREA path then has,
// for EVERY low-res column dx, exactly three x taps on source pixels 2dx, 2dx+1, 2dx+2 (weights from the table, all
// different per column), and the INTER_LINEAR x axis has general 11-bit coefficients with ONE slip: output pixel
// x = 2i+2 always blends low-res pixels (i, i+1); x = 2i+1 blends (i, i+1) left of the slip column and (i-1, i) from
// it on.  A lane owns a chunk of 8 output pixels = low-res pixels P0..P3 = source pixels 0..8 of its 27-byte window.
// ---------------------------------------------------------------------------------
// One source row of the lane's window (7 aligned words = bytes 0..27) -> the horizontal INTER_AREA values of its twelve
// low-res byte columns, hb[3p + c] = fl(fl(S[6p+c] a0 + S[6p+3+c] a1) + S[6p+6+c] a2) in OpenCV's operation order
// (buf = 0; buf += S * alpha per tap: separate multiply and add).  al[3p + t] = weight of tap t of low-res pixel p.
ROD_HD float byte_magic(const uint32_t* w, int i) {  // float 2^23 + (byte i of the window)
#if defined(__CUDA_ARCH__)
    return __uint_as_float(__byte_perm(w[i >> 2], 0x4B000000u, 0x7440 | (i & 3)));
#else
    return bitsf(0x4B000000u | ((w[i >> 2] >> (8 * (i & 3))) & 0xFFu));
#endif
}
ROD_HD void x2g_hrow(const uint32_t w[7], const float al[12], float hb[12]) {
    float x[27];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 27; ++i) x[i] = byte_magic(w, i);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int p = 0; p < 4; ++p) {
        // fma(2^23 + S, a, -2^23 a) == fl(S a) exactly (2^23 a is exact, one rounding)
        const float n0 = fmul(al[3 * p], -8388608.0f), n1 = fmul(al[3 * p + 1], -8388608.0f), n2 = fmul(al[3 * p + 2], -8388608.0f);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int c = 0; c < 3; ++c) {
            const float t0 = fmaf(x[6 * p + c], al[3 * p], n0);
            const float t1 = fmaf(x[6 * p + 3 + c], al[3 * p + 1], n1);
            const float t2 = fmaf(x[6 * p + 6 + c], al[3 * p + 2], n2);
            hb[3 * p + c] = fadd(fadd(t0, t1), t2);
        }
    }
}
// The same with the products -2^23 * weight supplied (nal[i] = fmul(al[i], -8388608.0f): constant per lane and tile).
ROD_HD void x2g_hrow_pre(const uint32_t w[7], const float al[12], const float nal[12], float hb[12]) {
    float x[27];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 27; ++i) x[i] = byte_magic(w, i);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int p = 0; p < 4; ++p) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int c = 0; c < 3; ++c) {
            const float t0 = fmaf(x[6 * p + c], al[3 * p], nal[3 * p]);
            const float t1 = fmaf(x[6 * p + 3 + c], al[3 * p + 1], nal[3 * p + 1]);
            const float t2 = fmaf(x[6 * p + 6 + c], al[3 * p + 2], nal[3 * p + 2]);
            hb[3 * p + c] = fadd(fadd(t0, t1), t2);
        }
    }
}
// vertical INTER_AREA tap on those values: first tap sum = beta * buf, later taps sum = sum + beta * buf
ROD_HD void x2g_vmac(const float hb[12], float beta, bool first, float acc[12]) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int q = 0; q < 12; ++q) {
        const float prod = fmul(beta, hb[q]);
        acc[q] = first ? prod : fadd(acc[q], prod);
    }
}
// saturate_cast<uchar>(rint(sum)) of the twelve sums -> the lane's 12 low-res bytes in three words.  (The weights are
// positive and sum to 1 +- 1e-6, so the sum lies in [0, 255.001]: the magic-number rint needs no clamp.)
ROD_HD void x2g_round12(const float acc[12], uint32_t own[3]) {
    uint32_t b[12];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int q = 0; q < 12; ++q) b[q] = fbits(fadd(acc[q], 12582912.0f));  // result in the low byte
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int k = 0; k < 3; ++k)
        own[k] = perm<0x5410>(perm<0x0040>(b[4 * k], b[4 * k + 1]), perm<0x0040>(b[4 * k + 2], b[4 * k + 3]));
}
// A chunk at the right image border with fewer than four valid low-res pixels: the missing ones replicate the last
// valid pixel (OpenCV clamps the second x tap to nw - 1).  valid = 0..3; for valid == 0 all four become the left
// neighbour's last pixel, passed in bytes 1..3 of `left_px`.
ROD_HD void x2g_replicate(uint32_t own[3], int valid, uint32_t left_px) {
    if (valid >= 4) return;
    if (valid <= 1) {
        const uint32_t p = valid == 0 ? (left_px >> 8) : own[0];   // the pixel in bytes 0..2
        own[0] = perm<0x0210>(p, p);
        own[1] = perm<0x1021>(p, p);
        own[2] = perm<0x2102>(p, p);
    } else if (valid == 2) {                                      // pixel 1 = bytes 3..5
        const uint32_t o1 = own[1];
        own[1] = perm<0x4354>(own[0], o1);
        own[2] = perm<0x5435>(own[0], o1);
    } else {                                                      // pixel 2 = bytes 6..8
        own[2] = perm<0x4324>(own[1], own[2]);
    }
}
// Horizontal INTER_LINEAR stage of one low-res row for the chunk: win[] = the 18 bytes of P[-1..4] (as in x2_expand24),
// coef[x] = a0 | a1 << 16 of output pixel x, slip bit x>>1 of `slip` set = odd pixel x blends (P[k-1], P[k]) instead of
// (P[k], P[k+1]), k = x >> 1.  x[t] = float 2^23 + ((L a0 + R a1) >> 4) for output byte t = 3 * pixel + channel.
ROD_HD void x2g_expand24(const uint32_t win[5], const uint32_t coef[8], uint32_t slip, float x[24]) {
    const uint32_t M = 0x4B000000u;
    uint32_t g01[5], g2x[5];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int m = 0; m < 5; ++m) { g01[m] = x2_gather(win, 3 * m); g2x[m] = x2_gather(win, 3 * m + 2); }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int px = 0; px < 8; ++px) {
        const int k = px >> 1;
        uint32_t a = g01[k], b = g2x[k];               // even pixels, and slipped odd pixels: pair k = (P[k-1], P[k])
        if (px & 1) {
            const bool sl = (slip >> k) & 1u;
            a = sl ? g01[k] : g01[k + 1];
            b = sl ? g2x[k] : g2x[k + 1];
        }
        const uint32_t h0 = dot2_lo(coef[px], a, M), h1 = dot2_hi(coef[px], a, M), h2 = dot2_lo(coef[px], b, M);
        // 2^23 + H -> 2^23 + (H >> 4): one multiply-add rounded toward zero
#if defined(__CUDA_ARCH__)
        x[3 * px + 0] = __fmaf_rz(bitsf(h0), 0.0625f, 7864320.0f);
        x[3 * px + 1] = __fmaf_rz(bitsf(h1), 0.0625f, 7864320.0f);
        x[3 * px + 2] = __fmaf_rz(bitsf(h2), 0.0625f, 7864320.0f);
#else
        x[3 * px + 0] = (float)(8388608.0 + (double)((h0 - M) >> 4));
        x[3 * px + 1] = (float)(8388608.0 + (double)((h1 - M) >> 4));
        x[3 * px + 2] = (float)(8388608.0 + (double)((h2 - M) >> 4));
#endif
    }
}
// Vertical-stage constants for x = 2^23 + hx (general horizontal stage): x2_vertical(x0, x1, r) then gives
// (((b0 * hx0) >> 16) + ((b1 * hx1) >> 16) + 2) >> 2 in its low byte:
//   y1 = x0 * b0/65536 + (2^23 - 128 b0) = 2^23 + F0;  y2 = x1 * b1/65536 + y1 = 2^23 + 128 b1 + F0 + F1 (< 2^24);
//   o = y2 / 4 + (2^23 - 2^21 - 32 b1 + 0.5) = 2^23 + ((F0 + F1 + 2) >> 2), every step rounded toward zero.
ROD_HD X2Row x2g_row_consts(uint32_t b_packed) {
    X2Row r;
    const float b0 = (float)(b_packed & 0xFFFFu), b1 = (float)(b_packed >> 16);
    r.c0s = b0 * 1.52587890625e-05f;
    r.c1s = b1 * 1.52587890625e-05f;
    r.k0 = 8388608.0f - b0 * 128.0f;
    r.k2 = 6291456.5f - b1 * 32.0f;
    return r;
}

// Detector-input normalisation: half(float(u8) / 255.f) is done with __float2half_rn on
// device; the host harness does not cover it (tests compare against numpy float16).

}  // namespace rod
