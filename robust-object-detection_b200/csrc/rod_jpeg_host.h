// rod_jpeg_host.h -- host side of the JPEG encoder: parse the header OpenCV writes for an image size (SOI .. SOS) into the
// quantisation reciprocals (jcdctmgr.c compute_reciprocal) and the Huffman code tables (jchuff.c jpeg_make_c_derived_tbl).
// Used by jpeg.cu (uploaded to the device) and by tests/emu (CPU check against cv2.imencode).
#pragma once
#include <string.h>

#include <vector>

#include "rod_jpeg.h"

namespace rod {
namespace jpeg {

struct HeaderInfo {
    int height = 0, width = 0;
    int h_samp[3] = {0, 0, 0}, v_samp[3] = {0, 0, 0}, tq[3] = {0, 0, 0};
    int td[3] = {0, 0, 0}, ta[3] = {0, 0, 0};
    int restart_interval = 0;
    size_t header_len = 0;   // bytes up to and including the SOS segment
    bool baseline = false;
};

// compute_reciprocal for divisor = 8 * q (the islow DCT's output scale), 16-bit DCTELEM
inline void reciprocal_entry(unsigned divisor, uint16_t* recip, uint16_t* corr, uint8_t* shift16) {
    int b = 0;
    while ((divisor >> (b + 1)) != 0) ++b;   // flss(divisor) - 1
    int r = 16 + b;
    uint32_t fq = ((uint32_t)1 << r) / divisor, fr = ((uint32_t)1 << r) % divisor;
    uint32_t c = divisor / 2;
    if (fr == 0) { fq >>= 1; r--; }
    else if (fr <= divisor / 2U) c++;
    else fq++;
    *recip = (uint16_t)fq; *corr = (uint16_t)c; *shift16 = (uint8_t)r;
}

// Returns false when the header is not the baseline 4:2:0 three-component layout this encoder implements.
inline bool parse_header(const uint8_t* h, size_t n, HeaderInfo* info, Tables* tb) {
    memset(tb, 0, sizeof(*tb));
    uint8_t qt[4][64];
    bool have_q[4] = {false, false, false, false};
    bool have_h[4] = {false, false, false, false};   // [0] DC0 [1] DC1 [2] AC0 [3] AC1
    size_t i = 0;
    if (n < 4 || h[0] != 0xFF || h[1] != 0xD8) return false;
    i = 2;
    while (i + 4 <= n) {
        if (h[i] != 0xFF) return false;
        const int m = h[i + 1];
        const size_t L = ((size_t)h[i + 2] << 8) | h[i + 3];
        if (i + 2 + L > n) return false;
        const uint8_t* p = h + i + 4;
        const size_t pl = L - 2;
        if (m == 0xDB) {
            size_t o = 0;
            while (o < pl) {
                const int pq = p[o] >> 4, t = p[o] & 15;
                if (pq != 0 || t > 3 || o + 65 > pl) return false;
                memcpy(qt[t], p + o + 1, 64);   // zigzag order
                have_q[t] = true;
                o += 65;
            }
        } else if (m == 0xC0) {
            if (pl < 15 || p[0] != 8 || p[5] != 3) return false;
            info->height = (p[1] << 8) | p[2];
            info->width = (p[3] << 8) | p[4];
            for (int c = 0; c < 3; ++c) {
                info->h_samp[c] = p[7 + 3 * c] >> 4;
                info->v_samp[c] = p[7 + 3 * c] & 15;
                info->tq[c] = p[8 + 3 * c];
            }
            info->baseline = true;
        } else if (m == 0xC4) {
            size_t o = 0;
            while (o + 17 <= pl) {
                const int tc = p[o] >> 4, th = p[o] & 15;
                if (tc > 1 || th > 1) return false;
                const uint8_t* bits = p + o + 1;   // bits[0] = number of codes of length 1
                int total = 0;
                for (int l = 0; l < 16; ++l) total += bits[l];
                if (o + 17 + total > pl || total > 256) return false;
                const uint8_t* val = p + o + 17;
                const int slot = 2 * tc + th;
                // jpeg_make_c_derived_tbl: codes in order of increasing length
                unsigned code = 0;
                int k = 0;
                for (int l = 1; l <= 16; ++l) {
                    for (int q = 0; q < bits[l - 1]; ++q, ++k) {
                        tb->ehufco[slot][val[k]] = (uint16_t)code;
                        tb->ehufsi[slot][val[k]] = (uint8_t)l;
                        ++code;
                    }
                    code <<= 1;
                }
                have_h[slot] = true;
                o += 17 + total;
            }
        } else if (m == 0xDD) {
            info->restart_interval = (p[0] << 8) | p[1];
        } else if (m == 0xDA) {
            if (pl < 10 || p[0] != 3) return false;
            for (int c = 0; c < 3; ++c) {
                info->td[c] = p[2 + 2 * c] >> 4;
                info->ta[c] = p[2 + 2 * c] & 15;
            }
            info->header_len = i + 2 + L;
            break;
        } else if (m >= 0xC1 && m <= 0xCF && m != 0xC4 && m != 0xC8 && m != 0xCC) {
            return false;   // not baseline
        }
        i += 2 + L;
    }
    if (!info->baseline || info->header_len == 0 || info->restart_interval != 0) return false;
    if (info->h_samp[0] != 2 || info->v_samp[0] != 2 || info->h_samp[1] != 1 || info->v_samp[1] != 1 || info->h_samp[2] != 1 ||
        info->v_samp[2] != 1)
        return false;
    if (info->tq[0] != 0 || info->tq[1] != 1 || info->tq[2] != 1 || !have_q[0] || !have_q[1]) return false;
    if (info->td[0] != 0 || info->ta[0] != 0 || info->td[1] != 1 || info->ta[1] != 1 || info->td[2] != 1 || info->ta[2] != 1) return false;
    for (int s = 0; s < 4; ++s)
        if (!have_h[s]) return false;
    for (int t = 0; t < 2; ++t)
        for (int z = 0; z < 64; ++z) {
            const int nat = natural_order(z);
            reciprocal_entry(8u * qt[t][z], &tb->recip[t][nat], &tb->corr[t][nat], &tb->shift[t][nat]);
        }
    return true;
}

}  // namespace jpeg
}  // namespace rod
