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
    const uint8_t* img;  // corrupted images, laid out by the plan's dst descriptors
    __half* out;
    int out_h, out_w;
    int pad;
};

__device__ __forceinline__ void lb_pixel(const LbParams& p, const DevImage& im, const DevLetterbox& g,
                                         const uint8_t* base, int Y, int X, uint32_t v[3]) {
    const int cy = Y - g.top, cx = X - g.left;
    if (cy < 0 || cy >= g.new_h || cx < 0 || cx >= g.new_w) {
        v[0] = v[1] = v[2] = (uint32_t)p.pad;
        return;
    }
    const int64_t pitch = im.dst_pitch;
    if (g.identity) {
        const uint8_t* s = base + (int64_t)cy * pitch + cx * 3;
        v[0] = s[0]; v[1] = s[1]; v[2] = s[2];
        return;
    }
    if (g.area2) {
        const uint8_t* s = base + (int64_t)(2 * cy) * pitch + (2 * cx) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = ((uint32_t)s[c] + s[3 + c] + s[pitch + c] + s[pitch + 3 + c] + 2u) >> 2;
        return;
    }
    const int32_t* lx_s0 = reinterpret_cast<const int32_t*>(p.tab + g.lx_s0);
    const int s0 = lx_s0[cx];
    const int s1 = min(s0 + 1, g.w - 1);
    const uint32_t a = p.tab[g.lx_a + cx];
    const uint32_t ys = p.tab[g.ly_s + cy], yb = p.tab[g.ly_b + cy];
    const uint8_t* r0 = base + (int64_t)(ys & 0xFFFFu) * pitch;
    const uint8_t* r1 = base + (int64_t)(ys >> 16) * pitch;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const uint32_t h0 = linear_h4(r0[3 * s0 + c], r0[3 * s1 + c], a);
        const uint32_t h1 = linear_h4(r1[3 * s0 + c], r1[3 * s1 + c], a);
        v[c] = linear_v(h0, h1, yb);
    }
}

__global__ void __launch_bounds__(256) letterbox_kernel(LbParams p) {
    const int tiles_x = (p.out_w + kLbTW - 1) / kLbTW;
    const size_t plane = (size_t)p.out_h * p.out_w;
    for (int ti = blockIdx.x; ti < p.n_tiles; ti += gridDim.x) {
        const Tile t = p.tiles[ti];
        const DevImage im = p.images[t.img];
        const DevLetterbox g = p.lb[im.shape_id];
        const uint8_t* base = p.img + im.dst_off;
        __half* o = p.out + (size_t)t.img * 3 * plane;
        (void)tiles_x;
        // 16 rows x 64 cols per tile, 2 pixels per thread -> 512 thread-items, 2 per thread
        for (int idx = threadIdx.x; idx < kLbTH * (kLbTW / 2); idx += blockDim.x) {
            const int r = idx / (kLbTW / 2), q = idx - r * (kLbTW / 2);
            const int Y = t.a + r, X = t.b + 2 * q;
            if (Y >= p.out_h || X >= p.out_w) continue;
            uint32_t v0[3], v1[3] = {0, 0, 0};
            lb_pixel(p, im, g, base, Y, X, v0);
            const bool two = (X + 1 < p.out_w);
            if (two) lb_pixel(p, im, g, base, Y, X + 1, v1);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                // BGR -> RGB: source channel c lands in plane 2 - c
                __half* dst = o + (size_t)(2 - c) * plane + (size_t)Y * p.out_w + X;
                const __half a = __float2half_rn(__fdiv_rn((float)v0[c], 255.0f));
                if (two && ((((uintptr_t)dst) & 3) == 0)) {
                    const __half b = __float2half_rn(__fdiv_rn((float)v1[c], 255.0f));
                    *reinterpret_cast<__half2*>(dst) = __halves2half2(a, b);
                } else {
                    dst[0] = a;
                    if (two) dst[1] = __float2half_rn(__fdiv_rn((float)v1[c], 255.0f));
                }
            }
        }
    }
}

int launch_letterbox(const rod_plan* plan, const uint8_t* img, void* out_f16, int pad_value, cudaStream_t stream) {
    if (plan->n_lb_tiles == 0) return ROD_OK;
    LbParams p;
    p.images = plan->d_images;
    p.lb = plan->d_lb;
    p.tab = plan->d_lb_tab;
    p.tiles = plan->d_lb_tiles;
    p.n_tiles = plan->n_lb_tiles;
    p.img = img;
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
