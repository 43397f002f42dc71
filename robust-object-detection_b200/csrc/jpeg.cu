// jpeg.cu -- SURVEY 8f rank 1: baseline JPEG encoding of a device-resident batch whose bytes equal cv2.imwrite's
// (OpenCV 4.13.0 = libjpeg-turbo 3.1.2 defaults: YCbCr 4:2:0, quality 95, standard Huffman tables, islow DCT).
//
// Reference: scripts/build_corrupted_testsets.py:124 / :164 (`cv2.imwrite(str(dst_img_dir / img_path.name), out)`): the
// JPEG files are what the evaluation scripts read, so their bytes are part of the test-set semantics.  The arithmetic is
// rod_jpeg.h (shared with the CPU check against cv2.imencode in tests/emu); this file is its parallel schedule:
//   jpeg_coef_kernel    a CTA per run of 16 MCUs: pixel tile staged in shared memory, colour conversion (+ h2v2 chroma) per
//                       2x2 quad, then one thread per 8x8 block: islow DCT, reciprocal quantisation, zigzag -> int16
//                       coefficients [MCU][6][64], written back coalesced
//   jpeg_dummy_kernel   libjpeg's dummy blocks of partial MCUs at the right / bottom edge (DC of the preceding block)
//   jpeg_bits_kernel    one thread per MCU: bit length of its entropy-coded segment
//   jpeg_scan_kernel    one CTA per image: exclusive prefix sum of the MCU bit lengths
//   jpeg_write_kernel   one thread per MCU: its bits at its absolute bit offset of a zeroed big-endian stream (whole 32-bit
//                       words stored, the two boundary words OR-ed atomically)
//   jpeg_ffcount / ffscan / stuffwrite   0x00 inserted after every 0xFF: a warp per 2 KB chunk counts, a CTA per image scans
//                       (+ EOI, length), a warp per chunk copies its bytes to their place
// The header (SOI .. SOS) is OpenCV's own for that image size and is prepended on the host.
#include <algorithm>
#include <new>

#include "rod_internal.h"
#include "rod_jpeg_host.h"

namespace rod {

struct JpegImage {
    uint64_t img_off;        // byte offset of the image in the pixel buffer
    int64_t pitch;
    int32_t h, w;
    int32_t mcu_w, n_mcu;
    uint32_t mcu_first;      // index of the image's first MCU in the batch-wide MCU arrays
    uint32_t pad_;
    uint64_t raw_off, raw_cap;   // unstuffed stream (bytes, 4-byte aligned offset)
    uint64_t out_off, out_cap;   // stuffed stream + EOI
};

struct JpegParams {
    const JpegImage* images;
    int n_images;
    const uint8_t* pixels;
    const jpeg::Tables* tables;
    int16_t* coef;           // [total MCUs][6][64], zigzag order
    uint32_t* mcu_bits;      // [total MCUs]: bit length, then (after the scan) bit offset inside the image's stream
    uint32_t* total_bits;    // [n_images]
    uint8_t* raw;            // unstuffed streams
    uint8_t* out;            // stuffed streams
    uint32_t* out_len;       // [n_images] bytes written to out (incl. EOI); 0xFFFFFFFF: a buffer was too small
    const uint32_t* mcu_image;   // [total MCUs / 64 + 1]: image of every 64th MCU (search start)
    uint32_t total_mcu;
};

__device__ __forceinline__ int jpeg_image_of(const JpegParams& p, uint32_t m) {
    int i = (int)p.mcu_image[m >> 6];
    while (i + 1 < p.n_images && p.images[i + 1].mcu_first <= m) ++i;
    return i;
}

// A CTA takes a run of kCoefMcus MCUs of one MCU row: (1) the 16 x (16 * kCoefMcus) pixel tile is staged in shared memory
// with coalesced loads (edge pixels replicated), (2) every thread converts 2 x 2 pixel quads to 4 luma bytes + one Cb + one
// Cr byte (h2v2 with the alternating bias) into planar shared arrays, (3) one thread per 8 x 8 block runs the DCT and the
// quantisation from those planes, (4) the zigzag coefficients leave through shared memory as coalesced 16-byte stores.
constexpr int kCoefMcus = 16;                       // MCUs per CTA -> 96 block threads
constexpr int kCoefThreads = 6 * kCoefMcus;
constexpr int kCoefTileW = 16 * kCoefMcus;          // pixels per tile row
struct CoefTile {
    uint32_t mcu_first;   // first MCU (batch-wide index) of the run
    int32_t image;
};

__global__ void __launch_bounds__(kCoefThreads) jpeg_coef_kernel(JpegParams p, const CoefTile* tiles) {
    __shared__ __align__(16) uint8_t pix[16][kCoefTileW * 3];
    __shared__ __align__(16) uint8_t yplane[16][kCoefTileW];
    __shared__ __align__(16) uint8_t cplane[2][8][kCoefTileW / 2];
    __shared__ __align__(16) int16_t cst[kCoefThreads][64 + 8];   // (+8: rows 144 bytes apart spread over the banks)
    const CoefTile tl = tiles[blockIdx.x];
    const JpegImage im = p.images[tl.image];
    const jpeg::Geometry g = jpeg::geometry(im.h, im.w);
    const int lm = (int)(tl.mcu_first - im.mcu_first), my = lm / im.mcu_w, mx0 = lm - my * im.mcu_w;
    const int n_mcu_here = min(kCoefMcus, im.mcu_w - mx0);
    const int x0 = 16 * mx0, y0 = 16 * my;
    const int tile_px = 16 * n_mcu_here;
    const uint8_t* img = p.pixels + im.img_off;
    // (1) stage: row r of the tile = image row min(y0 + r, h - 1); columns beyond w - 1 replicate the last pixel
    {
        const int valid_px = min(tile_px, im.w - x0);           // real pixels per row
        const int nbytes = 3 * valid_px;
        for (int r = 0; r < 16; ++r) {
            const uint8_t* srow = img + (int64_t)min(y0 + r, im.h - 1) * im.pitch + 3 * x0;
            for (int b = threadIdx.x; b < nbytes; b += kCoefThreads) pix[r][b] = srow[b];
            for (int b = nbytes + threadIdx.x; b < 3 * tile_px; b += kCoefThreads) {
                const int c = b % 3;
                pix[r][b] = srow[3 * (valid_px - 1) + c];
            }
        }
    }
    __syncthreads();
    // (2) quads: luma of the four pixels; chroma quad (cx, cy) averages tile rows of the CLAMPED chroma row (the last real
    // chroma row is replicated downwards, jcprepct.c), columns as staged
    {
        const int quads_x = tile_px >> 1;
        for (int q = threadIdx.x; q < 8 * quads_x; q += kCoefThreads) {
            const int qy = q / quads_x, qx = q - qy * quads_x;
#pragma unroll
            for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                for (int dx = 0; dx < 2; ++dx) {
                    const uint8_t* px = &pix[2 * qy + dy][3 * (2 * qx + dx)];
                    int yy, cb, cr;
                    jpeg::rgb_to_ycc(px[2], px[1], px[0], &yy, &cb, &cr);
                    yplane[2 * qy + dy][2 * qx + dx] = (uint8_t)yy;
                }
            int cy = 8 * my + qy;
            if (cy > g.ch - 1) cy = g.ch - 1;
            const int ra = 2 * cy - y0, rb = min(2 * cy + 1, im.h - 1) - y0;
            int scb = 0, scr = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint8_t* px = &pix[(k & 2) ? rb : ra][3 * (2 * qx + (k & 1))];
                int yy, cb, cr;
                jpeg::rgb_to_ycc(px[2], px[1], px[0], &yy, &cb, &cr);
                scb += cb; scr += cr;
            }
            const int bias = 1 + ((8 * mx0 + qx) & 1);
            cplane[0][qy][qx] = (uint8_t)((scb + bias) >> 2);
            cplane[1][qy][qx] = (uint8_t)((scr + bias) >> 2);
        }
    }
    __syncthreads();
    // (3) one thread per block
    const int mloc = threadIdx.x / 6, blk = threadIdx.x - 6 * mloc;
    if (mloc < n_mcu_here) {
        const int mx = mx0 + mloc;
        const bool real = blk >= 4 || (2 * mx + (blk & 1) < g.yblk_w && 2 * my + (blk >> 1) < g.yblk_h);
        int16_t* zz = cst[threadIdx.x];
        if (!real) {
#pragma unroll
            for (int z = 0; z < 64; ++z) zz[z] = 0;   // DC set by jpeg_dummy_kernel
        } else {
            int d[64];
            if (blk < 4) {
                const int bx = 16 * mloc + 8 * (blk & 1), by = 8 * (blk >> 1);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint2 v = *reinterpret_cast<const uint2*>(&yplane[by + j][bx]);
#pragma unroll
                    for (int i = 0; i < 8; ++i) d[8 * j + i] = (int)(((i < 4 ? v.x : v.y) >> (8 * (i & 3))) & 0xFFu) - 128;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint2 v = *reinterpret_cast<const uint2*>(&cplane[blk - 4][j][8 * mloc]);
#pragma unroll
                    for (int i = 0; i < 8; ++i) d[8 * j + i] = (int)(((i < 4 ? v.x : v.y) >> (8 * (i & 3))) & 0xFFu) - 128;
                }
            }
            jpeg::fdct_islow(d);
            const int tq = blk < 4 ? 0 : 1;
            const jpeg::Tables& tb = *p.tables;
#pragma unroll
            for (int z = 0; z < 64; ++z) {
                const int nat = jpeg::natural_order(z);
                zz[z] = (int16_t)jpeg::quantize(d[nat], tb.recip[tq][nat], tb.corr[tq][nat], tb.shift[tq][nat]);
            }
        }
    }
    __syncthreads();
    // (4) 6 * n_mcu_here blocks of 128 bytes, contiguous in global memory
    {
        uint4* dst = reinterpret_cast<uint4*>(p.coef + (size_t)tl.mcu_first * 6 * 64);
        const int n16 = 6 * n_mcu_here * 8;
        for (int u = threadIdx.x; u < n16; u += kCoefThreads) dst[u] = *reinterpret_cast<const uint4*>(&cst[u >> 3][8 * (u & 7)]);
    }
}

// jccoefct.c compress_data: dummy blocks of a partial MCU carry the DC of the block before them (in MCU order), AC zero.
__global__ void __launch_bounds__(128) jpeg_dummy_kernel(JpegParams p) {
    const uint32_t m = blockIdx.x * 128u + threadIdx.x;
    if (m >= p.total_mcu) return;
    const JpegImage im = p.images[jpeg_image_of(p, m)];
    const int lm = (int)(m - im.mcu_first), my = lm / im.mcu_w, mx = lm - my * im.mcu_w;
    const jpeg::Geometry g = jpeg::geometry(im.h, im.w);
    if (2 * mx + 1 < g.yblk_w && 2 * my + 1 < g.yblk_h) return;   // all four luma blocks are real
    int16_t* mc = p.coef + (size_t)m * 6 * 64;
    for (int blk = 1; blk < 4; ++blk) {
        const bool real = 2 * mx + (blk & 1) < g.yblk_w && 2 * my + (blk >> 1) < g.yblk_h;
        if (!real) mc[64 * blk] = mc[64 * (blk - 1)];
    }
}

struct GlobalBitWriter {   // MSB-first bits into a zeroed big-endian stream; whole words stored, boundary words OR-ed
    uint32_t* words;       // the image's stream as 32-bit words (device memory is little-endian: words are byte-swapped)
    uint64_t acc;          // pending bits, left-aligned below bit (64 - fill) ... kept in the high `fill` bits
    int fill;              // pending bit count (< 32 after a flush)
    uint32_t widx;         // index of the word the pending bits start in
    bool first;            // the next flushed word may be shared with the previous MCU
    __device__ __forceinline__ void init(uint8_t* stream, uint32_t bitpos) {
        words = reinterpret_cast<uint32_t*>(stream);
        widx = bitpos >> 5;
        fill = (int)(bitpos & 31u);   // the leading bits of the first word belong to the previous MCU: zeros here
        acc = 0;
        first = true;
    }
    __device__ __forceinline__ void put(uint32_t code, int size) {
        acc |= (uint64_t)code << (64 - fill - size);
        fill += size;
        if (fill >= 32) {
            const uint32_t w = (uint32_t)(acc >> 32);
            const uint32_t be = __byte_perm(w, 0u, 0x0123);
            if (first) { atomicOr(words + widx, be); first = false; }
            else words[widx] = be;
            ++widx;
            acc <<= 32;
            fill -= 32;
        }
    }
    __device__ __forceinline__ void finish() {
        if (fill > 0) atomicOr(words + widx, __byte_perm((uint32_t)(acc >> 32), 0u, 0x0123));
    }
};

template <typename Sink>
__device__ __forceinline__ void jpeg_encode_mcu(const JpegParams& p, const JpegImage& im, uint32_t m, Sink& sink) {
    const int16_t* mc = p.coef + (size_t)m * 6 * 64;
    const jpeg::Tables& tb = *p.tables;
    const bool first_mcu = (m == im.mcu_first);
#pragma unroll 1
    for (int blk = 0; blk < 6; ++blk) {
        int last;
        if (blk >= 1 && blk <= 3) last = mc[64 * (blk - 1)];
        else if (first_mcu) last = 0;
        else last = (mc - 6 * 64)[64 * (blk == 0 ? 3 : blk)];
        const int hs = blk < 4 ? 0 : 1;
        jpeg::encode_block(mc + 64 * blk, last, tb.ehufco[hs], tb.ehufsi[hs], tb.ehufco[2 + hs], tb.ehufsi[2 + hs], sink);
    }
}

__global__ void __launch_bounds__(128) jpeg_bits_kernel(JpegParams p) {
    const uint32_t m = blockIdx.x * 128u + threadIdx.x;
    if (m >= p.total_mcu) return;
    const JpegImage im = p.images[jpeg_image_of(p, m)];
    jpeg::BitCounter bc;
    jpeg_encode_mcu(p, im, m, bc);
    p.mcu_bits[m] = bc.bits;
}

// one CTA per image: mcu_bits := exclusive prefix sum; total_bits[image] := sum
__global__ void __launch_bounds__(1024) jpeg_scan_kernel(JpegParams p) {
    __shared__ uint32_t warp_sum[32];
    __shared__ uint32_t carry_s;
    const JpegImage im = p.images[blockIdx.x];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < im.n_mcu; base += 1024) {
        const int i = base + (int)threadIdx.x;
        const uint32_t v = i < im.n_mcu ? p.mcu_bits[im.mcu_first + i] : 0u;
        uint32_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sum[warp] = x;
        __syncthreads();
        if (warp == 0) {
            uint32_t s = warp_sum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, s, o);
                if (lane >= o) s += y;
            }
            warp_sum[lane] = s;   // inclusive sums of the warps
        }
        __syncthreads();
        const uint32_t before = carry_s + (warp > 0 ? warp_sum[warp - 1] : 0u) + (x - v);
        if (i < im.n_mcu) p.mcu_bits[im.mcu_first + i] = before;
        __syncthreads();
        if (threadIdx.x == 0) carry_s += warp_sum[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const uint32_t bits = carry_s, nbytes = (bits + 7u) >> 3;
        p.total_bits[blockIdx.x] = bits;
        // flush_bits: the last byte is padded with ones.  OR-ed into the zeroed stream now: the MCU that ends in this word
        // finishes it with an atomicOr as well (a stream that ends on a word boundary has nothing to pad).
        if ((bits & 7u) != 0u && (uint64_t)nbytes + 8 <= im.raw_cap)
            atomicOr(reinterpret_cast<uint32_t*>(p.raw + im.raw_off) + ((nbytes - 1u) >> 2),
                     (0xFFu >> (bits & 7u)) << (8u * ((nbytes - 1u) & 3u)));
    }
}

__global__ void __launch_bounds__(128) jpeg_write_kernel(JpegParams p) {
    const uint32_t m = blockIdx.x * 128u + threadIdx.x;
    if (m >= p.total_mcu) return;
    const int ii = jpeg_image_of(p, m);
    const JpegImage im = p.images[ii];
    if (((uint64_t)p.total_bits[ii] + 7) / 8 + 8 > im.raw_cap) return;   // too small: reported by jpeg_ffscan_kernel
    GlobalBitWriter w;
    w.init(p.raw + im.raw_off, p.mcu_bits[m]);
    jpeg_encode_mcu(p, im, m, w);
    w.finish();
}

// ---- byte stuffing: out := raw with 0x00 inserted after every 0xFF, + EOI.  A warp per chunk of kStuffChunk stream bytes:
// (a) count the 0xFF bytes per chunk, (b) a CTA per image scans the counts (and writes EOI + the length), (c) every warp
// copies its chunk to its place, 128 bytes per step (a lane's four bytes go out as byte stores at the lane's scanned offset)
constexpr int kStuffChunk = 2048;

__device__ __forceinline__ uint32_t ff_bytes(uint32_t w) { return (uint32_t)__popc(__vcmpeq4(w, 0xFFFFFFFFu)) >> 3; }

__global__ void __launch_bounds__(128) jpeg_ffcount_kernel(JpegParams p, uint32_t* ff_count, const uint32_t* chunk_first) {
    const int img = blockIdx.y, lane = threadIdx.x & 31;
    const uint32_t c = blockIdx.x * 4u + (threadIdx.x >> 5);
    const JpegImage im = p.images[img];
    const uint32_t nbytes = (p.total_bits[img] + 7u) >> 3;
    if ((uint64_t)nbytes + 8 > im.raw_cap || (uint64_t)c * kStuffChunk >= nbytes) return;
    const uint8_t* raw = p.raw + im.raw_off + (size_t)c * kStuffChunk;
    const uint32_t n_here = min((uint32_t)kStuffChunk, nbytes - c * kStuffChunk);
    uint32_t cnt = 0;
#pragma unroll
    for (int step = 0; step < kStuffChunk / 512; ++step) {
        const uint32_t b0 = 512u * step + 16u * lane;
        if (b0 < n_here) {
            const uint4 v = *reinterpret_cast<const uint4*>(raw + b0);   // (bytes beyond the stream's end are zero)
            cnt += ff_bytes(v.x) + ff_bytes(v.y) + ff_bytes(v.z) + ff_bytes(v.w);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xFFFFFFFFu, cnt, o);
    if (lane == 0) ff_count[chunk_first[img] + c] = cnt;
}

__global__ void __launch_bounds__(1024) jpeg_ffscan_kernel(JpegParams p, uint32_t* ff_count, const uint32_t* chunk_first) {
    __shared__ uint32_t warp_sum[32];
    __shared__ uint32_t carry_s;
    const int img = blockIdx.x;
    const JpegImage im = p.images[img];
    const uint32_t nbytes = (p.total_bits[img] + 7u) >> 3;
    if ((uint64_t)nbytes + 8 > im.raw_cap) {
        if (threadIdx.x == 0) p.out_len[img] = 0xFFFFFFFFu;
        return;
    }
    const int n_chunks = (int)((nbytes + kStuffChunk - 1) / kStuffChunk);
    uint32_t* cnt = ff_count + chunk_first[img];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < n_chunks; base += 1024) {
        const int i = base + (int)threadIdx.x;
        const uint32_t v = i < n_chunks ? cnt[i] : 0u;
        uint32_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sum[warp] = x;
        __syncthreads();
        if (warp == 0) {
            uint32_t sm = warp_sum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, sm, o);
                if (lane >= o) sm += y;
            }
            warp_sum[lane] = sm;
        }
        __syncthreads();
        if (i < n_chunks) cnt[i] = carry_s + (warp > 0 ? warp_sum[warp - 1] : 0u) + (x - v);   // 0xFF bytes before the chunk
        __syncthreads();
        if (threadIdx.x == 0) carry_s += warp_sum[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const uint64_t end = (uint64_t)nbytes + carry_s;
        if (end + 2 > im.out_cap) p.out_len[img] = 0xFFFFFFFFu;
        else {
            uint8_t* out = p.out + im.out_off;
            out[end] = 0xFF; out[end + 1] = 0xD9;
            p.out_len[img] = (uint32_t)(end + 2);
        }
    }
}

__global__ void __launch_bounds__(128) jpeg_stuffwrite_kernel(JpegParams p, const uint32_t* ff_before, const uint32_t* chunk_first) {
    const int img = blockIdx.y, lane = threadIdx.x & 31;
    const uint32_t c = blockIdx.x * 4u + (threadIdx.x >> 5);
    const JpegImage im = p.images[img];
    const uint32_t nbytes = (p.total_bits[img] + 7u) >> 3;
    if ((uint64_t)nbytes + 8 > im.raw_cap || (uint64_t)c * kStuffChunk >= nbytes) return;
    if (p.out_len[img] == 0xFFFFFFFFu) return;
    const uint8_t* raw = p.raw + im.raw_off + (size_t)c * kStuffChunk;
    const uint32_t n_here = min((uint32_t)kStuffChunk, nbytes - c * kStuffChunk);
    uint8_t* out = p.out + im.out_off + (size_t)c * kStuffChunk + ff_before[chunk_first[img] + c];
    uint32_t run = 0;   // 0xFF bytes of the chunk's earlier steps
#pragma unroll 1
    for (uint32_t b0 = 0; b0 < n_here; b0 += 128u) {
        const uint32_t mine = b0 + 4u * lane;
        const uint32_t w = mine < n_here ? *reinterpret_cast<const uint32_t*>(raw + mine) : 0u;
        const uint32_t valid = mine < n_here ? min(4u, n_here - mine) : 0u;
        const uint32_t wm = valid == 4u ? w : (w & ((1u << (8u * valid)) - 1u));   // (valid == 0: w is 0 already)
        const uint32_t nff = ff_bytes(wm);
        uint32_t x = nff;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
            if (lane >= o) x += y;
        }
        uint8_t* o = out + mine + run + (x - nff);
        if (nff == 0u && valid == 4u) {
            o[0] = (uint8_t)w; o[1] = (uint8_t)(w >> 8); o[2] = (uint8_t)(w >> 16); o[3] = (uint8_t)(w >> 24);
        } else {
            for (uint32_t k = 0; k < valid; ++k) {
                const uint8_t b = (uint8_t)(w >> (8u * k));
                *o++ = b;
                if (b == 0xFF) *o++ = 0;
            }
        }
        run += __shfl_sync(0xFFFFFFFFu, x, 31);
    }
}

}  // namespace rod

using namespace rod;

// ---- host side -------------------------------------------------------------------------------------------------------
// Encoders are created per batch layout; their (large) device buffers come from a small cache of blocks keyed by rounded
// size, so that building one encoder per batch does not pay cudaMalloc / cudaFree (a device-wide synchronisation) each time.
#include <map>
#include <mutex>
#include <vector>
namespace {
std::mutex g_cache_mutex;
std::multimap<std::pair<int, size_t>, void*> g_cache;   // (device, rounded size) -> free block
size_t round_block(size_t n) {
    size_t r = 1 << 16;   // (plans keep ~25 small arrays each)
    while (r < n) r <<= 1;
    if (r > (64u << 20)) r = (n + (64u << 20) - 1) / (64u << 20) * (64u << 20);   // large blocks: multiples of 64 MiB
    return r;
}
size_t g_cache_bytes = 0;                                     // bytes parked in the cache
size_t cache_limit() {   // more than this is handed back to the driver (knob: ROD_BLOCK_CACHE_MB, default 16 GiB)
    static const size_t lim = [] {
        const char* e = getenv("ROD_BLOCK_CACHE_MB");
        return (size_t)(e && atol(e) >= 0 ? atol(e) : 16384) << 20;
    }();
    return lim;
}
cudaError_t cached_alloc(int dev, void** p, size_t n) {
    const size_t r = round_block(n);
    {
        std::lock_guard<std::mutex> lock(g_cache_mutex);
        auto it = g_cache.find({dev, r});
        if (it != g_cache.end()) { *p = it->second; g_cache.erase(it); g_cache_bytes -= r; return cudaSuccess; }
    }
    cudaError_t e = cudaMalloc(p, r);
    if (e == cudaErrorMemoryAllocation) {   // hand the parked blocks of this device back to the driver and try once more
        std::vector<void*> parked;
        {
            std::lock_guard<std::mutex> lock(g_cache_mutex);
            for (auto it = g_cache.begin(); it != g_cache.end();) {
                if (it->first.first == dev) { parked.push_back(it->second); g_cache_bytes -= it->first.second; it = g_cache.erase(it); }
                else ++it;
            }
        }
        (void)cudaGetLastError();
        if (!parked.empty()) {
            cudaDeviceSynchronize();   // blocks are parked after their owner synchronised; this covers the caller's own streams
            for (void* q : parked) cudaFree(q);
            e = cudaMalloc(p, r);
        }
    }
    return e;
}
void cached_free(int dev, void* p, size_t n) {   // dev: the device the block was allocated on
    if (p == nullptr) return;
    const size_t r = round_block(n);
    {
        std::lock_guard<std::mutex> lock(g_cache_mutex);
        if (g_cache_bytes + r <= cache_limit()) {
            g_cache.insert({{dev, r}, p});
            g_cache_bytes += r;
            return;
        }
    }
    cudaFree(p);
}
}  // namespace

namespace rod {   // the same cache serves the decoder (jpegdec.cu)
cudaError_t block_cache_alloc(int dev, void** p, size_t n) { return cached_alloc(dev, p, n); }
void block_cache_free(int dev, void* p, size_t n) { cached_free(dev, p, n); }
}  // namespace rod

// releases the cached device blocks of destroyed encoders
extern "C" void rod_jpeg_trim(void) {
    std::lock_guard<std::mutex> lock(g_cache_mutex);
    for (auto& kv : g_cache) cudaFree(kv.second);
    g_cache.clear();
    g_cache_bytes = 0;
}

struct rod_jpeg_encoder {
    int device = 0;
    int n_images = 0;
    std::vector<JpegImage> h_images;
    std::vector<uint32_t> h_mcu_image;
    uint32_t total_mcu = 0;
    uint64_t raw_bytes = 0, out_bytes = 0;
    JpegImage* d_images = nullptr;
    uint32_t* d_mcu_image = nullptr;
    CoefTile* d_coef_tiles = nullptr;
    int n_coef_tiles = 0;
    jpeg::Tables* d_tables = nullptr;
    int16_t* d_coef = nullptr;
    uint32_t* d_mcu_bits = nullptr;
    uint32_t* d_total_bits = nullptr;
    uint8_t* d_raw = nullptr;
    uint8_t* d_out = nullptr;
    uint32_t* d_out_len = nullptr;
    uint32_t* d_ff_count = nullptr;      // per stuffing chunk: 0xFF bytes in it, then (after the scan) before it
    uint32_t* d_chunk_first = nullptr;   // per image: index of its first chunk
    uint32_t total_chunks = 0, max_chunks = 0;
    std::vector<uint64_t> out_off;
};

extern "C" void rod_jpeg_destroy(rod_jpeg_encoder* e) {
    if (e == nullptr) return;
    // every buffer goes to the block cache and may be handed out again at once: kernels still reading them must be done
    cudaDeviceSynchronize();
    const size_t n = (size_t)e->n_images;
    cached_free(e->device, e->d_images, sizeof(JpegImage) * n);
    cached_free(e->device, e->d_mcu_image, sizeof(uint32_t) * e->h_mcu_image.size());
    cached_free(e->device, e->d_coef_tiles, sizeof(CoefTile) * (size_t)e->n_coef_tiles);
    cached_free(e->device, e->d_tables, sizeof(jpeg::Tables));
    cached_free(e->device, e->d_total_bits, sizeof(uint32_t) * n);
    cached_free(e->device, e->d_out_len, sizeof(uint32_t) * n);
    cached_free(e->device, e->d_ff_count, sizeof(uint32_t) * (size_t)e->total_chunks);
    cached_free(e->device, e->d_chunk_first, sizeof(uint32_t) * n);
    cached_free(e->device, e->d_coef, (size_t)e->total_mcu * 6 * 64 * sizeof(int16_t));
    cached_free(e->device, e->d_mcu_bits, sizeof(uint32_t) * (size_t)e->total_mcu);
    cached_free(e->device, e->d_raw, e->raw_bytes + 64);
    cached_free(e->device, e->d_out, e->out_bytes + 64);
    delete e;
}

// `images`: n descriptors (src_offset / src_pitch / height / width describe the pixels to encode; the dst_* fields are
// ignored).  `header`: the bytes SOI .. SOS OpenCV writes with the wanted parameters (any image size): its DQT / DHT
// segments define the tables.  Output streams are laid out back to back; rod_jpeg_stream_offset tells where.
extern "C" int rod_jpeg_create(const rod_image_desc* images, int n_images, const uint8_t* header, uint64_t header_len,
                               rod_jpeg_encoder** out_enc) {
    if (out_enc == nullptr) return ROD_ERR_INVALID_ARG;
    *out_enc = nullptr;
    if (images == nullptr || n_images < 1 || header == nullptr) return ROD_ERR_INVALID_ARG;
    jpeg::HeaderInfo info;
    jpeg::Tables tb;
    if (!jpeg::parse_header(header, (size_t)header_len, &info, &tb)) return ROD_ERR_UNSUPPORTED;
    rod_jpeg_encoder* e = new (std::nothrow) rod_jpeg_encoder();
    if (e == nullptr) return ROD_ERR_OOM;
    if (cudaGetDevice(&e->device) != cudaSuccess) { delete e; cudaGetLastError(); return ROD_ERR_NO_DEVICE; }
    e->n_images = n_images;
    e->h_images.resize(n_images);
    e->out_off.resize(n_images + 1);
    uint64_t mcu = 0, raw = 0, out = 0;
    for (int i = 0; i < n_images; ++i) {
        const rod_image_desc& d = images[i];
        if (d.height < 1 || d.width < 1 || d.height > 65535 || d.width > 65535 || d.src_pitch < 3LL * d.width) {
            delete e;
            return ROD_ERR_INVALID_ARG;
        }
        JpegImage& im = e->h_images[i];
        im.img_off = d.src_offset; im.pitch = d.src_pitch; im.h = d.height; im.w = d.width;
        im.mcu_w = (d.width + 15) / 16;
        im.n_mcu = im.mcu_w * ((d.height + 15) / 16);
        im.mcu_first = (uint32_t)mcu;
        im.pad_ = 0;
        // an MCU of uniform noise codes to ~0.6 of its 768 raw bytes at quality 95; 1.25x raw + 4 KB leaves ample room and is
        // checked on the device (out_len = 0xFFFFFFFF: the caller falls back to its own encoder for that image)
        const uint64_t cap = ((uint64_t)im.n_mcu * 960 + 4096 + 15) & ~(uint64_t)15;
        im.raw_off = raw; im.raw_cap = cap;
        im.out_off = out; im.out_cap = cap + cap / 8;
        e->out_off[i] = out;
        raw += cap;
        out += (im.out_cap + 15) & ~(uint64_t)15;
        mcu += (uint64_t)im.n_mcu;
        if (mcu > 0x7FFFFFFFull) { delete e; return ROD_ERR_UNSUPPORTED; }
    }
    e->out_off[n_images] = out;
    e->total_mcu = (uint32_t)mcu;
    e->raw_bytes = raw; e->out_bytes = out;
    e->h_mcu_image.resize(e->total_mcu / 64 + 2);
    {
        int img = 0;
        for (uint32_t b = 0; b < e->h_mcu_image.size(); ++b) {
            const uint64_t m = (uint64_t)b * 64;
            while (img + 1 < n_images && e->h_images[img + 1].mcu_first <= m) ++img;
            e->h_mcu_image[b] = (uint32_t)img;
        }
    }
    std::vector<CoefTile> ctiles;
    for (int i = 0; i < n_images; ++i) {
        const JpegImage& im = e->h_images[i];
        const int mcu_h = im.n_mcu / im.mcu_w;
        for (int my = 0; my < mcu_h; ++my)
            for (int mx = 0; mx < im.mcu_w; mx += kCoefMcus) ctiles.push_back(CoefTile{im.mcu_first + (uint32_t)(my * im.mcu_w + mx), i});
    }
    e->n_coef_tiles = (int)ctiles.size();
    std::vector<uint32_t> chunk_first(n_images);
    for (int i = 0; i < n_images; ++i) {
        const uint32_t nc = (uint32_t)((e->h_images[i].raw_cap + kStuffChunk - 1) / kStuffChunk);
        chunk_first[i] = e->total_chunks;
        e->total_chunks += nc;
        e->max_chunks = std::max(e->max_chunks, nc);
    }
    cudaError_t err = cudaSuccess;
    auto alloc = [&](void** p, size_t n) { if (err == cudaSuccess) err = cached_alloc(e->device, p, n); };   // (sizes repeated in rod_jpeg_destroy)
    alloc((void**)&e->d_coef_tiles, sizeof(CoefTile) * ctiles.size());
    if (err == cudaSuccess) err = cudaMemcpy(e->d_coef_tiles, ctiles.data(), sizeof(CoefTile) * ctiles.size(), cudaMemcpyHostToDevice);
    alloc((void**)&e->d_images, sizeof(JpegImage) * n_images);
    alloc((void**)&e->d_mcu_image, sizeof(uint32_t) * e->h_mcu_image.size());
    alloc((void**)&e->d_tables, sizeof(jpeg::Tables));
    auto calloc_ = [&](void** p, size_t n) { if (err == cudaSuccess) err = cached_alloc(e->device, p, n); };
    calloc_((void**)&e->d_coef, (size_t)e->total_mcu * 6 * 64 * sizeof(int16_t));
    calloc_((void**)&e->d_mcu_bits, sizeof(uint32_t) * (size_t)e->total_mcu);
    alloc((void**)&e->d_total_bits, sizeof(uint32_t) * n_images);
    calloc_((void**)&e->d_raw, e->raw_bytes + 64);
    calloc_((void**)&e->d_out, e->out_bytes + 64);
    alloc((void**)&e->d_out_len, sizeof(uint32_t) * n_images);
    alloc((void**)&e->d_ff_count, sizeof(uint32_t) * (size_t)e->total_chunks);
    alloc((void**)&e->d_chunk_first, sizeof(uint32_t) * n_images);
    if (err == cudaSuccess) err = cudaMemcpy(e->d_chunk_first, chunk_first.data(), sizeof(uint32_t) * n_images, cudaMemcpyHostToDevice);
    if (err == cudaSuccess) err = cudaMemcpy(e->d_images, e->h_images.data(), sizeof(JpegImage) * n_images, cudaMemcpyHostToDevice);
    if (err == cudaSuccess) err = cudaMemcpy(e->d_mcu_image, e->h_mcu_image.data(), sizeof(uint32_t) * e->h_mcu_image.size(), cudaMemcpyHostToDevice);
    if (err == cudaSuccess) err = cudaMemcpy(e->d_tables, &tb, sizeof(tb), cudaMemcpyHostToDevice);
    if (err != cudaSuccess) {
        const int rc = cuda_fail(err);
        rod_jpeg_destroy(e);
        return rc;
    }
    *out_enc = e;
    return ROD_OK;
}

extern "C" uint64_t rod_jpeg_stream_offset(const rod_jpeg_encoder* e, int i) {
    return (e != nullptr && i >= 0 && i <= e->n_images) ? e->out_off[i] : 0;
}
extern "C" const uint8_t* rod_jpeg_stream_base(const rod_jpeg_encoder* e) { return e ? e->d_out : nullptr; }
extern "C" const uint32_t* rod_jpeg_stream_lengths(const rod_jpeg_encoder* e) { return e ? e->d_out_len : nullptr; }

// Encode every image of `pixels` (device pointer, laid out by the descriptors given at creation).  Afterwards the device
// arrays rod_jpeg_stream_base() + rod_jpeg_stream_offset(i) hold the entropy-coded segment + EOI of image i and
// rod_jpeg_stream_lengths()[i] its length (0xFFFFFFFF: did not fit); the file is OpenCV's header for that size + those bytes.
extern "C" int rod_jpeg_encode(rod_jpeg_encoder* e, const uint8_t* pixels, void* stream_) {
    if (e == nullptr || pixels == nullptr) return ROD_ERR_INVALID_ARG;
    cudaStream_t stream = (cudaStream_t)stream_;
    JpegParams p;
    p.images = e->d_images; p.n_images = e->n_images; p.pixels = pixels; p.tables = e->d_tables; p.coef = e->d_coef;
    p.mcu_bits = e->d_mcu_bits; p.total_bits = e->d_total_bits; p.raw = e->d_raw; p.out = e->d_out; p.out_len = e->d_out_len;
    p.mcu_image = e->d_mcu_image; p.total_mcu = e->total_mcu;
    ROD_CUDA(cudaMemsetAsync(e->d_raw, 0, e->raw_bytes + 64, stream));
    const unsigned mcu_blocks = (e->total_mcu + 127u) / 128u;
    jpeg_coef_kernel<<<e->n_coef_tiles, kCoefThreads, 0, stream>>>(p, e->d_coef_tiles);
    jpeg_dummy_kernel<<<mcu_blocks, 128, 0, stream>>>(p);
    jpeg_bits_kernel<<<mcu_blocks, 128, 0, stream>>>(p);
    jpeg_scan_kernel<<<e->n_images, 1024, 0, stream>>>(p);
    jpeg_write_kernel<<<mcu_blocks, 128, 0, stream>>>(p);
    const dim3 chunk_grid((e->max_chunks + 3u) / 4u, (unsigned)e->n_images);
    jpeg_ffcount_kernel<<<chunk_grid, 128, 0, stream>>>(p, e->d_ff_count, e->d_chunk_first);
    jpeg_ffscan_kernel<<<e->n_images, 1024, 0, stream>>>(p, e->d_ff_count, e->d_chunk_first);
    jpeg_stuffwrite_kernel<<<chunk_grid, 128, 0, stream>>>(p, e->d_ff_count, e->d_chunk_first);
    ROD_CUDA(cudaGetLastError());
    return ROD_OK;
}

// Copy the lengths and the encoded streams to the host: host_len[n_images]; image i's bytes land at
// host_out + rod_jpeg_stream_offset(i) (host_out: rod_jpeg_stream_offset(n_images) bytes, ideally page-locked).
// Returns after the copies are complete.  Images whose length is 0xFFFFFFFF are skipped.
extern "C" int rod_jpeg_download(rod_jpeg_encoder* e, uint8_t* host_out, uint32_t* host_len, void* stream_) {
    if (e == nullptr || host_out == nullptr || host_len == nullptr) return ROD_ERR_INVALID_ARG;
    cudaStream_t stream = (cudaStream_t)stream_;
    ROD_CUDA(cudaMemcpyAsync(host_len, e->d_out_len, sizeof(uint32_t) * e->n_images, cudaMemcpyDeviceToHost, stream));
    ROD_CUDA(cudaStreamSynchronize(stream));
    for (int i = 0; i < e->n_images; ++i) {
        if (host_len[i] == 0xFFFFFFFFu || host_len[i] == 0) continue;
        ROD_CUDA(cudaMemcpyAsync(host_out + e->out_off[i], e->d_out + e->out_off[i], host_len[i], cudaMemcpyDeviceToHost, stream));
    }
    ROD_CUDA(cudaStreamSynchronize(stream));
    return ROD_OK;
}
