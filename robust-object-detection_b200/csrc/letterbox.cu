// letterbox.cu -- training-path output stage (BASELINE config 5): uint8 HWC BGR -> fp16 NCHW RGB
// with Ultralytics' LetterBox (INTER_LINEAR resize-to-fit + constant pad) and /255 normalise.
//
// Not part of /root/reference (it lives in Ultralytics 8.3.x, which the reference's
// train_*_augmented.py scripts call through model.train(...)); the resize arithmetic is the same
// 8-bit fixed-point INTER_LINEAR of SURVEY 8a/a5, the rest is pad(114) + channel swap +
// half(float(u8) / 255.f).  One thread produces two horizontally adjacent output pixels
// (3 channels each) and writes one half2 per colour plane.
#include <cuda_fp16.h>

#include "rod_internal.h"

namespace rod {

struct LbParams {
    const DevImage* images;
    const DevLetterbox* lb;
    const uint32_t* tab;
    const Tile* tiles;
    int n_tiles;
    const uint8_t* img;   // corrupted images, laid out by the plan's dst descriptors
    const uint8_t* src;   // original images (src descriptors): read directly when opcodes[i] == ROD_OP_NONE
    const uint8_t* opcodes;
    __half* out;
    int out_h, out_w;
    int pad;
};

constexpr int kLbRowsPerThread = kLbTH / 4;  // 256 threads = 64 columns x 4 row groups

// Six consecutive bytes (two BGR pixels) starting at an arbitrary address, as byte0..3 / byte4..5 of (lo, hi).
// Aligned 32-bit loads; a word is read only if at least one of its bytes lies below `safe_end` (the end of the
// image), so the over-read is at most 3 bytes inside an aligned word that is partly valid (never a new page).
__device__ __forceinline__ void load6(const uint8_t* p, const uint8_t* safe_end, uint32_t& lo, uint32_t& hi) {
    const uint8_t* a = reinterpret_cast<const uint8_t*>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)3);
    const int sh = (int)(reinterpret_cast<uintptr_t>(p) & 3) * 8;
    const uint32_t w0 = __ldg(reinterpret_cast<const uint32_t*>(a));
    const uint32_t w1 = (a + 4 < safe_end) ? __ldg(reinterpret_cast<const uint32_t*>(a + 4)) : 0u;
    const uint32_t w2 = (sh != 0 && a + 8 < safe_end) ? __ldg(reinterpret_cast<const uint32_t*>(a + 8)) : 0u;
    lo = __funnelshift_r(w0, w1, sh);
    hi = __funnelshift_r(w1, w2, sh);
}

__global__ void __launch_bounds__(256) letterbox_kernel(LbParams p) {
    __shared__ __half lut[256];  // half(float(v) / 255.f): the normalisation of preprocess_batch, one division per value
    lut[threadIdx.x] = __float2half_rn(__fdiv_rn((float)threadIdx.x, 255.0f));
    __syncthreads();
    const size_t plane = (size_t)p.out_h * p.out_w;
    const int cx_l = threadIdx.x & (kLbTW - 1), rgrp = threadIdx.x / kLbTW;
    const __half padh = lut[p.pad];
    for (int ti = blockIdx.x; ti < p.n_tiles; ti += gridDim.x) {
        const Tile t = p.tiles[ti];
        const DevImage im = p.images[t.img];
        const DevLetterbox g = p.lb[im.shape_id];
        const bool from_src = (p.opcodes != nullptr && p.opcodes[t.img] == ROD_OP_NONE);
        const uint8_t* base = from_src ? p.src + im.src_off : p.img + im.dst_off;
        const int64_t pitch = from_src ? im.src_pitch : im.dst_pitch;
        const uint8_t* safe_end = base + (int64_t)(g.h - 1) * pitch + 3 * g.w;
        __half* o = p.out + (size_t)t.img * 3 * plane;
        const int X = t.b + cx_l;
        if (X >= p.out_w) continue;
        const int cx = X - g.left;
        const bool col_in = cx >= 0 && cx < g.new_w;
        int s0 = 0;
        uint32_t a = 0;
        bool one_tap = false;
        if (col_in && !g.identity && !g.area2) {
            s0 = reinterpret_cast<const int32_t*>(p.tab + g.lx_s0)[cx];
            a = p.tab[g.lx_a + cx];
            one_tap = (s0 + 1 > g.w - 1);  // right border: both taps are the last pixel
        }
        const int Ya = t.a + rgrp * kLbRowsPerThread;
#pragma unroll 2
        for (int rr = 0; rr < kLbRowsPerThread; ++rr) {
            const int Y = Ya + rr;
            if (Y >= p.out_h) break;
            const int cy = Y - g.top;
            uint32_t v[3];
            bool content = col_in && cy >= 0 && cy < g.new_h;
            if (content) {
                if (g.identity) {
                    const uint8_t* sp = base + (int64_t)cy * pitch + cx * 3;
                    v[0] = sp[0]; v[1] = sp[1]; v[2] = sp[2];
                } else if (g.area2) {
                    const uint8_t* sp = base + (int64_t)(2 * cy) * pitch + (2 * cx) * 3;
#pragma unroll
                    for (int c = 0; c < 3; ++c) v[c] = ((uint32_t)sp[c] + sp[3 + c] + sp[pitch + c] + sp[pitch + 3 + c] + 2u) >> 2;
                } else {
                    const uint32_t ys = p.tab[g.ly_s + cy], yb = p.tab[g.ly_b + cy];
                    const uint8_t* r0 = base + (int64_t)(ys & 0xFFFFu) * pitch + 3 * s0;
                    const uint8_t* r1 = base + (int64_t)(ys >> 16) * pitch + 3 * s0;
                    uint32_t lo0, hi0, lo1, hi1;
                    load6(r0, safe_end, lo0, hi0);
                    load6(r1, safe_end, lo1, hi1);
                    if (one_tap) {  // pixel 1 := pixel 0
                        hi0 = __byte_perm(lo0, 0u, 0x4421); lo0 = __byte_perm(lo0, 0u, 0x0210);
                        hi1 = __byte_perm(lo1, 0u, 0x4421); lo1 = __byte_perm(lo1, 0u, 0x0210);
                    }
                    // channel c: taps are bytes c and c + 3 of the six
                    const uint32_t g0[3] = {__byte_perm(lo0, hi0, 0x4430), __byte_perm(lo0, hi0, 0x4441), __byte_perm(lo0, hi0, 0x4452)};
                    const uint32_t g1[3] = {__byte_perm(lo1, hi1, 0x4430), __byte_perm(lo1, hi1, 0x4441), __byte_perm(lo1, hi1, 0x4452)};
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        const uint32_t h0 = dot2_lo(a, g0[c], 0u) >> 4;
                        const uint32_t h1 = dot2_lo(a, g1[c], 0u) >> 4;
                        v[c] = linear_v(h0, h1, yb);
                    }
                }
            }
            __half* dstp = o + (size_t)Y * p.out_w + X;
            // BGR -> RGB: source channel c lands in plane 2 - c
            dstp[2 * plane] = content ? lut[v[0]] : padh;
            dstp[plane] = content ? lut[v[1]] : padh;
            dstp[0] = content ? lut[v[2]] : padh;
        }
    }
}

int launch_letterbox(const rod_plan* plan, const uint8_t* img, const uint8_t* src, const uint8_t* opcodes, void* out_f16,
                     int pad_value, cudaStream_t stream) {
    if (plan->n_lb_tiles == 0) return ROD_OK;
    LbParams p;
    p.images = plan->d_images;
    p.lb = plan->d_lb;
    p.tab = plan->d_lb_tab;
    p.tiles = plan->d_lb_tiles;
    p.n_tiles = plan->n_lb_tiles;
    p.img = img;
    p.src = src;
    p.opcodes = opcodes;
    p.out = reinterpret_cast<__half*>(out_f16);
    p.out_h = plan->lb_out_h;
    p.out_w = plan->lb_out_w;
    p.pad = pad_value;
    const int grid = grid_for(plan, plan->n_lb_tiles, 8);
    letterbox_kernel<<<grid, 256, 0, stream>>>(p);
    ROD_CUDA(cudaGetLastError());
    return ROD_OK;
}

}  // namespace rod
