// restoration.cu -- SURVEY 8f rank 4: the (corrupted, clean) training pairs of the restoration U-Net, on device.
//
// Reference: scripts/train_restoration.py:104-129 (RestorationDataset.__getitem__): crop a patch (a strided view,
// :84,:93), optional cv2.flip(patch, 1), clean = patch.copy(), corrupted = one of the three corruptions of the
// patch (:95-102), then for both: BGR -> RGB, float32 / 255.0, HWC -> CHW.  Here:
//   gather_patches_kernel  crop + optional horizontal flip -> contiguous uint8 patches (the "clean" bytes)
//   (the corruption kernels of this library run on those patches, op-code per patch)
//   format_pairs_kernel    both uint8 patch sets -> float32 [N,3,P,P] RGB planes, value / 255 (one IEEE division)
#include "rod_internal.h"

namespace rod {

struct GatherParams {
    const DevImage* images;  // outer plan: src_off / src_pitch describe the crop, (h, w) the patch size
    const DevImage* inner;   // inner plan: contiguous patch layout (src == dst offsets)
    int n_images;
    const uint8_t* src;
    uint8_t* clean;
    const uint8_t* flips;    // 1: horizontal flip (cv2.flip(patch, 1)), may be NULL
};

__global__ void __launch_bounds__(256) gather_patches_kernel(GatherParams p) {
    // one CTA row-group per (patch, 8 rows): blockIdx.y = patch, blockIdx.x = row group
    const int i = blockIdx.y;
    const DevImage im = p.images[i];
    const DevImage in = p.inner[i];
    const bool flip = p.flips != nullptr && p.flips[i] != 0;
    const int n = 3 * im.w;
    for (int y = blockIdx.x * 8 + (threadIdx.x >> 5); y < im.h; y += gridDim.x * 8) {
        const uint8_t* srow = p.src + im.src_off + (int64_t)y * im.src_pitch;
        uint8_t* drow = p.clean + in.src_off + (int64_t)y * in.src_pitch;
        for (int b = threadIdx.x & 31; b < n; b += 32) {
            const int px = b / 3, c = b - 3 * px;
            drow[b] = flip ? srow[3 * (im.w - 1 - px) + c] : srow[b];
        }
    }
}

struct FormatParams {
    const DevImage* inner;
    int n_images;
    const uint8_t* clean;
    const uint8_t* corrupted;
    float* clean_out;      // [N,3,h,w]
    float* corrupted_out;  // [N,3,h,w]
    uint64_t plane_stride; // floats per image (3 * h * w of the first image; all patches share one size)
};

__global__ void __launch_bounds__(256) format_pairs_kernel(FormatParams p) {
    __shared__ float lut[256];  // float32(v) / 255.0f
    lut[threadIdx.x] = __fdiv_rn((float)threadIdx.x, 255.0f);
    __syncthreads();
    const int i = blockIdx.y;
    const DevImage in = p.inner[i];
    const size_t plane = (size_t)in.h * in.w;
    for (int y = blockIdx.x * 8 + (threadIdx.x >> 5); y < in.h; y += gridDim.x * 8) {
        const uint8_t* a = p.clean + in.src_off + (int64_t)y * in.src_pitch;
        const uint8_t* b = p.corrupted + in.src_off + (int64_t)y * in.src_pitch;
        float* oa = p.clean_out + (size_t)i * p.plane_stride + (size_t)y * in.w;
        float* ob = p.corrupted_out + (size_t)i * p.plane_stride + (size_t)y * in.w;
        for (int x = threadIdx.x & 31; x < in.w; x += 32) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {  // BGR -> RGB: source channel c lands in plane 2 - c
                oa[(size_t)(2 - c) * plane + x] = lut[a[3 * x + c]];
                ob[(size_t)(2 - c) * plane + x] = lut[b[3 * x + c]];
            }
        }
    }
}

// ---- resize-first branch (train_restoration.py:79-81, 88-90): cv2.resize(img, (nw, nh)) with the default INTER_LINEAR
// on 8-bit data = fixed-point coefficients (11 bits) per axis, horizontal stage (S0 a0 + S1 a1) >> 4, vertical stage
// (((b0 h0) >> 16) + ((b1 h1) >> 16) + 2) >> 2 (rod_core.h linear_h4 / linear_v).  The per-axis coefficients are
// computed in the kernel exactly as rod_tables.h build_linear_axis does on the host (double scale, float fraction,
// round-half-even to 11 bits; x: index and fraction clamped at both ends, y: fraction kept, indices clipped).
struct ResizeParams {
    const uint8_t* src;
    uint8_t* dst;
    int h, w, nh, nw;
    int64_t src_pitch, dst_pitch;
    double scale_x, scale_y;  // 1.0 / ((double)dsize / (double)ssize)
};

__device__ __forceinline__ void linear_axis_entry(int d, double scale, int ssize, bool clamp_x, int* s0, int* s1, uint32_t* coef) {
    float f = (float)__dadd_rn(__dmul_rn((double)d + 0.5, scale), -0.5);
    int s = (int)floorf(f);
    f = __fsub_rn(f, (float)s);
    if (clamp_x) {
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= ssize - 1) { s = ssize - 1; f = 0.f; }
    }
    const float c0 = __fsub_rn(1.f, f);
    const int a0 = __float2int_rn(__fmul_rn(c0, 2048.f)), a1 = __float2int_rn(__fmul_rn(f, 2048.f));
    *s0 = min(max(s, 0), ssize - 1);
    *s1 = min(max(s + 1, 0), ssize - 1);
    *coef = (uint32_t)(a0 & 0xFFFF) | ((uint32_t)(a1 & 0xFFFF) << 16);
}

__global__ void __launch_bounds__(256) resize_linear_kernel(ResizeParams p) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= p.nw || y >= p.nh) return;
    int xs0, xs1, ys0, ys1;
    uint32_t xa, yb;
    linear_axis_entry(x, p.scale_x, p.w, true, &xs0, &xs1, &xa);
    linear_axis_entry(y, p.scale_y, p.h, false, &ys0, &ys1, &yb);
    const uint8_t* r0 = p.src + (int64_t)ys0 * p.src_pitch;
    const uint8_t* r1 = p.src + (int64_t)ys1 * p.src_pitch;
    uint8_t* o = p.dst + (int64_t)y * p.dst_pitch + 3 * x;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const uint32_t h0 = linear_h4(r0[3 * xs0 + c], r0[3 * xs1 + c], xa);
        const uint32_t h1 = linear_h4(r1[3 * xs0 + c], r1[3 * xs1 + c], xa);
        o[c] = (uint8_t)linear_v(h0, h1, yb);
    }
}

int launch_resize_linear(const uint8_t* src, int h, int w, int64_t src_pitch, uint8_t* dst, int nh, int nw, int64_t dst_pitch,
                         cudaStream_t stream) {
    ResizeParams p;
    p.src = src; p.dst = dst; p.h = h; p.w = w; p.nh = nh; p.nw = nw; p.src_pitch = src_pitch; p.dst_pitch = dst_pitch;
    p.scale_x = 1.0 / ((double)nw / (double)w);
    p.scale_y = 1.0 / ((double)nh / (double)h);
    const dim3 grid((unsigned)((nw + 31) / 32), (unsigned)((nh + 7) / 8));
    resize_linear_kernel<<<grid, 256, 0, stream>>>(p);
    ROD_CUDA(cudaGetLastError());
    return ROD_OK;
}

int launch_gather_patches(const rod_plan* plan, const rod_plan* inner, const uint8_t* src, uint8_t* clean,
                          const uint8_t* flips, cudaStream_t stream) {
    GatherParams p;
    p.images = plan->d_images; p.inner = inner->d_images; p.n_images = plan->n_images;
    p.src = src; p.clean = clean; p.flips = flips;
    dim3 grid(8, plan->n_images);
    gather_patches_kernel<<<grid, 256, 0, stream>>>(p);
    ROD_CUDA(cudaGetLastError());
    return ROD_OK;
}

int launch_format_pairs(const rod_plan* inner, const uint8_t* clean, const uint8_t* corrupted, float* clean_out,
                        float* corrupted_out, cudaStream_t stream) {
    FormatParams p;
    p.inner = inner->d_images; p.n_images = inner->n_images;
    p.clean = clean; p.corrupted = corrupted; p.clean_out = clean_out; p.corrupted_out = corrupted_out;
    p.plane_stride = 3ull * inner->h_images[0].h * inner->h_images[0].w;
    dim3 grid(8, inner->n_images);
    format_pairs_kernel<<<grid, 256, 0, stream>>>(p);
    ROD_CUDA(cudaGetLastError());
    return ROD_OK;
}

}  // namespace rod
