// rod_jpeg.h -- arithmetic of a baseline JPEG encoder whose bytes equal cv2.imwrite's, written once for device and host.
//
// the files scripts/build_corrupted_testsets.py writes (`cv2.imwrite(str(dst / name), out)`, :124 / :164)
// are part of the test-set semantics, so a device encoder has to reproduce OpenCV 4.13.0's encoder bit for bit.  That
// encoder is libjpeg-turbo 3.1.2 with its defaults: baseline sequential DCT, YCbCr 4:2:0 (h2v2 chroma), quality 95,
// standard Huffman tables, no restart markers, JDCT_ISLOW.  The integer algorithms restated here (all published in the
// IJG / libjpeg-turbo sources, whose SIMD paths are bit-identical to their C paths):
//   jccolor.c   rgb_ycc_convert     16-bit fixed-point RGB -> YCbCr
//   jcsample.c  h2v2_downsample     2x2 box with the alternating bias 1, 2, 1, 2 ...; right / bottom edges replicated
//   jfdctint.c  jpeg_fdct_islow     13-bit fixed-point 8x8 forward DCT, output scaled by 8
//   jcdctmgr.c  quantize            division by 8 * q through a 16-bit reciprocal (compute_reciprocal), round half away
//   jchuff.c    encode_one_block    DC difference / AC run-length categories, MSB-first bit packing, 0xFF00 stuffing
// The header bytes (SOI .. SOS: JFIF APP0, DQT, SOF0, DHT) depend only on the image size and are taken from OpenCV itself
// (host side); the quantisation and Huffman tables used below are PARSED from that header.
// The same functions are compiled into tests/emu (CPU check against cv2.imencode) and into jpeg.cu.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define RJ_HD __host__ __device__ __forceinline__
#else
#define RJ_HD inline
#endif

namespace rod {
namespace jpeg {

// zigzag position -> natural (row-major) index
RJ_HD int natural_order(int z) {
    const uint8_t t[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                           41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                           30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
    return t[z];
}

// Tables of one encoder configuration (parsed from the header on the host, copied to the device).
struct Tables {
    // quantisation, per table (0: luma, 1: chroma), natural order: reciprocal, correction, shift of jcdctmgr.c compute_reciprocal
    uint16_t recip[2][64];
    uint16_t corr[2][64];
    uint8_t shift[2][64];   // shift + 16
    // Huffman: code and length per symbol; [0] DC luma, [1] DC chroma, [2] AC luma, [3] AC chroma
    uint16_t ehufco[4][256];
    uint8_t ehufsi[4][256];
};

// jccolor.c rgb_ycc_convert (SCALEBITS 16): one pixel.
RJ_HD void rgb_to_ycc(int r, int g, int b, int* y, int* cb, int* cr) {
    *y = (19595 * r + 38470 * g + 7471 * b + 32768) >> 16;
    *cb = (-11059 * r - 21709 * g + 32768 * b + 8421375) >> 16;
    *cr = (32768 * r - 27439 * g - 5329 * b + 8421375) >> 16;
}

// Geometry of one image in MCUs of 16 x 16 pixels (4:2:0).
struct Geometry {
    int h, w;
    int mcu_w, mcu_h;        // MCUs per row / column
    int yblk_w, yblk_h;      // real luma blocks per row / column: ceil(w / 8), ceil(h / 8) (the others are dummy blocks)
    int ch;                  // real chroma rows: ceil(h / 2)
};
RJ_HD Geometry geometry(int h, int w) {
    Geometry g;
    g.h = h; g.w = w;
    g.mcu_w = (w + 15) >> 4; g.mcu_h = (h + 15) >> 4;
    g.yblk_w = (w + 7) >> 3; g.yblk_h = (h + 7) >> 3;
    g.ch = (h + 1) >> 1;
    return g;
}

// The 64 level-shifted samples of block `blk` (0..3: luma blocks in raster order inside the MCU, 4: Cb, 5: Cr) of MCU
// (mx, my) from an HWC BGR image.  Luma: pixels beyond the right / bottom edge replicate the last column / row
// (jcsample.c expand_right_edge, jcprepct.c expand_bottom_edge).  Chroma: h2v2_downsample of the colour-converted image
// whose columns are replicated at full resolution and whose rows are replicated at the DOWNSAMPLED resolution (the last
// real chroma row is copied down), bias 1, 2, 1, 2, ... along a row.
RJ_HD void block_samples(const uint8_t* img, long pitch, const Geometry& g, int mx, int my, int blk, int* d) {
    if (blk < 4) {
        const int x0 = 16 * mx + 8 * (blk & 1), y0 = 16 * my + 8 * (blk >> 1);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = 0; j < 8; ++j) {
            const int y = (y0 + j < g.h) ? y0 + j : g.h - 1;
            const uint8_t* row = img + (long)y * pitch;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int i = 0; i < 8; ++i) {
                const int x = (x0 + i < g.w) ? x0 + i : g.w - 1;
                const uint8_t* px = row + 3 * x;
                int yy, cb, cr;
                rgb_to_ycc(px[2], px[1], px[0], &yy, &cb, &cr);
                d[8 * j + i] = yy - 128;
            }
        }
        return;
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 0; j < 8; ++j) {
        int cy = 8 * my + j;
        if (cy > g.ch - 1) cy = g.ch - 1;
        const int ya = (2 * cy < g.h) ? 2 * cy : g.h - 1, yb = (2 * cy + 1 < g.h) ? 2 * cy + 1 : g.h - 1;
        const uint8_t* ra = img + (long)ya * pitch;
        const uint8_t* rb = img + (long)yb * pitch;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int i = 0; i < 8; ++i) {
            const int cx = 8 * mx + i;
            const int xa = (2 * cx < g.w) ? 2 * cx : g.w - 1, xb = (2 * cx + 1 < g.w) ? 2 * cx + 1 : g.w - 1;
            int sum = 0;
            const uint8_t* p4[4] = {ra + 3 * xa, ra + 3 * xb, rb + 3 * xa, rb + 3 * xb};
            for (int q = 0; q < 4; ++q) {
                int yy, cb, cr;
                rgb_to_ycc(p4[q][2], p4[q][1], p4[q][0], &yy, &cb, &cr);
                sum += (blk == 4) ? cb : cr;
            }
            d[8 * j + i] = ((sum + 1 + (cx & 1)) >> 2) - 128;
        }
    }
}

// jfdctint.c jpeg_fdct_islow on 64 ints (level-shifted samples in, coefficients scaled by 8 out), in place.
RJ_HD void fdct_islow(int* d) {
    const int C298 = 2446, C390 = 3196, C541 = 4433, C765 = 6270, C899 = 7373, C1175 = 9633, C1501 = 12299, C1847 = 15137,
              C1961 = 16069, C2053 = 16819, C2562 = 20995, C3072 = 25172;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int pass = 0; pass < 2; ++pass) {
        // pass 0: rows (stride 1 inside a row), pass 1: columns (stride 8)
        const int es = pass == 0 ? 1 : 8, ls = pass == 0 ? 8 : 1;
        const int sh_even = pass == 0 ? 0 : 2;          // pass 1: DESCALE(x, PASS1_BITS); pass 0: x << PASS1_BITS
        const int sh_odd = pass == 0 ? 11 : 15;         // CONST_BITS -+ PASS1_BITS
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int l = 0; l < 8; ++l) {
            int* p = d + l * ls;
            const int d0 = p[0], d1 = p[es], d2 = p[2 * es], d3 = p[3 * es], d4 = p[4 * es], d5 = p[5 * es], d6 = p[6 * es],
                      d7 = p[7 * es];
            const int tmp0 = d0 + d7, tmp7 = d0 - d7, tmp1 = d1 + d6, tmp6 = d1 - d6, tmp2 = d2 + d5, tmp5 = d2 - d5,
                      tmp3 = d3 + d4, tmp4 = d3 - d4;
            const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
            if (pass == 0) {
                p[0] = (tmp10 + tmp11) << 2;
                p[4 * es] = (tmp10 - tmp11) << 2;
            } else {
                p[0] = (tmp10 + tmp11 + (1 << (sh_even - 1))) >> sh_even;
                p[4 * es] = (tmp10 - tmp11 + (1 << (sh_even - 1))) >> sh_even;
            }
            const int rnd = 1 << (sh_odd - 1);
            int z1 = (tmp12 + tmp13) * C541;
            p[2 * es] = (z1 + tmp13 * C765 + rnd) >> sh_odd;
            p[6 * es] = (z1 + tmp12 * (-C1847) + rnd) >> sh_odd;
            z1 = tmp4 + tmp7;
            int z2 = tmp5 + tmp6, z3 = tmp4 + tmp6, z4 = tmp5 + tmp7;
            const int z5 = (z3 + z4) * C1175;
            const int t4 = tmp4 * C298, t5 = tmp5 * C2053, t6 = tmp6 * C3072, t7 = tmp7 * C1501;
            z1 *= -C899; z2 *= -C2562; z3 *= -C1961; z4 *= -C390;
            z3 += z5; z4 += z5;
            p[7 * es] = (t4 + z1 + z3 + rnd) >> sh_odd;
            p[5 * es] = (t5 + z2 + z4 + rnd) >> sh_odd;
            p[3 * es] = (t6 + z2 + z3 + rnd) >> sh_odd;
            p[es] = (t7 + z1 + z4 + rnd) >> sh_odd;
        }
    }
}

// jcdctmgr.c quantize (the C path of libjpeg-turbo: 16-bit DCTELEM): coefficient -> quantised value.
RJ_HD int quantize(int coef, uint32_t recip, uint32_t corr, int shift16) {
    int t = (int)(int16_t)coef;
    if (t < 0) {
        const uint32_t product = ((uint32_t)(-t) + corr) * recip;
        return -(int)(int16_t)(product >> shift16);
    }
    const uint32_t product = ((uint32_t)t + corr) * recip;
    return (int)(int16_t)(product >> shift16);
}

// number of bits needed for |v| (JPEG category), v != 0 allowed up to 16 bits
RJ_HD int nbits_of(int v) {
    int a = v < 0 ? -v : v;
#if defined(__CUDA_ARCH__)
    return 32 - __clz(a);
#else
    int n = 0;
    while (a) { ++n; a >>= 1; }
    return n;
#endif
}

// Bit sink: `put(code, size)` appends the `size` low bits of `code`, MSB first.  Two implementations: a counter (pass 1)
// and a writer into a zero-initialised word buffer at an absolute bit offset (pass 2).
struct BitCounter {
    uint32_t bits = 0;
    RJ_HD void put(uint32_t, int size) { bits += (uint32_t)size; }
};

// jchuff.c encode_one_block for one block of quantised coefficients in ZIGZAG order (16-byte aligned); last_dc = the
// previous block's DC of the same component.  The coefficients are read eight at a time (one 16-byte load) and a chunk of
// eight zeros only extends the current run.
struct Coef8 {
    uint32_t w[4];
};
RJ_HD Coef8 load_coef8(const int16_t* zz, int chunk) {
    Coef8 c;
#if defined(__CUDA_ARCH__)
    const uint4 v = reinterpret_cast<const uint4*>(zz)[chunk];
    c.w[0] = v.x; c.w[1] = v.y; c.w[2] = v.z; c.w[3] = v.w;
#else
    for (int i = 0; i < 4; ++i)
        c.w[i] = (uint32_t)(uint16_t)zz[8 * chunk + 2 * i] | ((uint32_t)(uint16_t)zz[8 * chunk + 2 * i + 1] << 16);
#endif
    return c;
}
template <typename Sink>
RJ_HD void encode_block(const int16_t* zz, int last_dc, const uint16_t* dc_co, const uint8_t* dc_si, const uint16_t* ac_co,
                        const uint8_t* ac_si, Sink& sink) {
    int r = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int chunk = 0; chunk < 8; ++chunk) {
        const Coef8 c = load_coef8(zz, chunk);
        if (chunk > 0 && (c.w[0] | c.w[1] | c.w[2] | c.w[3]) == 0u) { r += 8; continue; }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int i = 0; i < 8; ++i) {
            int temp = (int)(int16_t)(c.w[i >> 1] >> (16 * (i & 1)));
            if (chunk == 0 && i == 0) {   // DC difference
                temp -= last_dc;
                int temp2 = temp;
                if (temp < 0) { temp = -temp; temp2--; }
                const int nbits = nbits_of(temp);
                sink.put(dc_co[nbits], dc_si[nbits]);
                if (nbits) sink.put((uint32_t)temp2 & ((1u << nbits) - 1u), nbits);
                continue;
            }
            if (temp == 0) { ++r; continue; }
            while (r > 15) { sink.put(ac_co[0xF0], ac_si[0xF0]); r -= 16; }
            int temp2 = temp;
            if (temp < 0) { temp = -temp; temp2--; }
            const int nbits = nbits_of(temp);
            const int sym = (r << 4) + nbits;
            sink.put(ac_co[sym], ac_si[sym]);
            sink.put((uint32_t)temp2 & ((1u << nbits) - 1u), nbits);
            r = 0;
        }
    }
    if (r > 0) sink.put(ac_co[0], ac_si[0]);
}

}  // namespace jpeg
}  // namespace rod
