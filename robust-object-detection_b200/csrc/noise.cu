// noise.cu -- a1: fused uint8 -> fp32 -> saturate -> uint8 Gaussian-noise kernel (sm_100a).
//
// Reference: scripts/augmentations.py:30-33 (apply_noise).  Two sources for the field:
//   NOISE_COMPAT  a supplied float32 tensor (what np.random.normal(...).astype(f32) drew) -> bit-exact
//   NOISE_PHILOX  Philox4x32-10 + Box-Muller generated in registers, keyed by
//                 (seed, global image index, element/4, offset): no HBM traffic for the field
// plus NOISE_COPY (ROD_OP_NONE images of a mixed batch) and NOISE_FIELD (dump the Philox field).
//
// HBM-bound elementwise op: one work item is a span of <= 16384 consecutive bytes; each
// thread moves 16 bytes per step with 128-bit loads/stores that bypass L1.  Spans whose
// addresses are not 16-byte aligned (pitched rows of strided views) take a byte path.
#include <stdlib.h>

#include "rod_internal.h"

namespace rod {

struct NoiseParams {
    const DevImage* images;
    const Tile* tiles;
    int n_tiles;
    const uint8_t* src;
    uint8_t* dst;
    const float* noise;
    float* field_out;
    float sigma;
    uint32_t key0, key1;
    uint64_t first_image;
    uint32_t offset;
    const uint8_t* opcodes;
    int my_op;
    unsigned int* counter;  // zeroed before the launch
};

__device__ __forceinline__ uint4 ldg_stream16(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ float4 ldg_stream16f(const void* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint32_t ldg_stream4(const void* p) {
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg4(void* p, uint32_t v) {
    asm volatile("st.global.L1::no_allocate.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void stg16(void* p, uint4 v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w)
                 : "memory");
}

// byte b of word -> exact float, without I2F: 0x4B0000bb is 2^23 + bb
__device__ __forceinline__ float byte_to_float(uint32_t word, int b) {
    return __uint_as_float(__byte_perm(word, 0x4B000000u, 0x7440 | b)) - 8388608.0f;
}

__device__ __forceinline__ uint32_t noise_word(uint32_t word, float n0, float n1, float n2, float n3) {
    uint32_t o0 = noise_px(byte_to_float(word, 0), n0);
    uint32_t o1 = noise_px(byte_to_float(word, 1), n1);
    uint32_t o2 = noise_px(byte_to_float(word, 2), n2);
    uint32_t o3 = noise_px(byte_to_float(word, 3), n3);
    return o0 | (o1 << 8) | (o2 << 16) | (o3 << 24);
}

// One Philox group: the eight factors s[0..7] of elements 8g .. 8g+7 (rod_core.h gauss8).
__device__ __forceinline__ void group_gauss8(const NoiseParams& p, uint32_t ig_lo, uint32_t ig_hi, uint32_t g,
                                             float s[8]) {
    uint32_t r[4];
    philox4x32_10(g, ig_lo, ig_hi, p.offset, p.key0, p.key1, r);
    if (philox_needs_tail(r)) {  // 2^-14 of the groups: refine the radius of the words whose high half is 0
        uint32_t t[4];
        philox4x32_10(g, ig_lo, ig_hi ^ ROD_PHILOX_TAIL_FLIP, p.offset, p.key0, p.key1, t);
        gauss8(r, t, s);
    } else {
        gauss8(r, nullptr, s);
    }
}

template <int MODE>
__global__ void __launch_bounds__(256, 4) noise_kernel(NoiseParams p) {
    const float K = p.sigma * ROD_NOISE_K_PER_SIGMA;
    // a warp takes a quarter span (<= 4096 bytes) at a time from a shared counter: no tail, no block-level coupling
    const uint32_t lane = threadIdx.x & 31u;
    for (;;) {
        uint32_t id = 0;
        if (lane == 0) id = atomicAdd(p.counter, 1u);
        id = __shfl_sync(0xFFFFFFFFu, id, 0);
        if ((int)(id >> 2) >= p.n_tiles) break;
        Tile t = p.tiles[id >> 2];
        if (p.opcodes != nullptr && p.opcodes[t.img] != p.my_op) continue;
        const int piece0 = (int)(id & 3u) * (kNoiseSpan / 4);
        if (piece0 >= t.b) continue;
        t.a += piece0;
        t.b = min(kNoiseSpan / 4, t.b - piece0);
        const DevImage im = p.images[t.img];
        const uint64_t img_global = p.first_image + (uint64_t)t.img;
        const uint32_t ig_lo = (uint32_t)img_global, ig_hi = (uint32_t)(img_global >> 32);
        const uint8_t* s = nullptr;
        uint8_t* d = nullptr;
        uint32_t e0;  // element index (inside the image) of the span's first byte
        if (t.c < 0) {
            e0 = (uint32_t)t.a;
            if (MODE != NOISE_FIELD) { s = p.src + im.src_off + e0; d = p.dst + im.dst_off + e0; }
        } else {
            e0 = (uint32_t)t.c * 3u * (uint32_t)im.w + (uint32_t)t.a;
            if (MODE != NOISE_FIELD) {
                s = p.src + im.src_off + (int64_t)t.c * im.src_pitch + t.a;
                d = p.dst + im.dst_off + (int64_t)t.c * im.dst_pitch + t.a;
            }
        }
        const uint32_t n = (uint32_t)t.b;
        const float* nzp = (MODE == NOISE_COMPAT) ? p.noise + im.elem_base + e0 : nullptr;
        float* fout = (MODE == NOISE_FIELD) ? p.field_out + im.elem_base + e0 : nullptr;

        uint32_t done = 0;  // elements of the span finished by the vector paths
        if (MODE == NOISE_COMPAT) {
            // 4 pixels (one word) + one float4 of the field per thread step; every warp instruction is fully
            // coalesced (128 B of pixels, 512 B of field); four independent steps in flight per thread
            const bool vec4 = (e0 & 3u) == 0 && ((((uintptr_t)s) | ((uintptr_t)d)) & 3) == 0 && (((uintptr_t)nzp) & 15) == 0;
            const uint32_t nw = vec4 ? (n >> 2) : 0;
            for (uint32_t base = lane; base < nw; base += 4 * 32u) {
                uint32_t px[4];
                float4 f[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t idx = base + u * 32u;
                    if (idx < nw) {
                        px[u] = ldg_stream4(s + 4 * idx);
                        f[u] = ldg_stream16f(nzp + 4 * idx);
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t idx = base + u * 32u;
                    if (idx < nw) stg4(d + 4 * idx, noise_word(px[u], f[u].x, f[u].y, f[u].z, f[u].w));
                }
            }
            done = nw << 2;
        } else {
            // 16 bytes per thread step = two Philox groups: needs the span to start on a group boundary
            bool vec = (e0 & 7u) == 0;
            if (MODE != NOISE_FIELD) vec = vec && ((((uintptr_t)s) | ((uintptr_t)d)) & 15) == 0;
            if (MODE == NOISE_FIELD) vec = vec && (((uintptr_t)fout) & 15) == 0;
            const uint32_t nvec = vec ? (n >> 4) : 0;
            for (uint32_t i = lane; i < nvec; i += 32u) {
                const uint32_t e = 16u * i;
                uint4 v = make_uint4(0, 0, 0, 0);
                if (MODE != NOISE_FIELD) v = ldg_stream16(s + e);
                if (MODE == NOISE_COPY) {
                    stg16(d + e, v);
                    continue;
                }
                const uint32_t in[4] = {v.x, v.y, v.z, v.w};
                uint32_t out[4];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    float sf[8];
                    group_gauss8(p, ig_lo, ig_hi, ((e0 + e) >> 3) + h, sf);
                    if (MODE == NOISE_FIELD) {
                        float4* fo = reinterpret_cast<float4*>(fout + e + 8 * h);
                        fo[0] = make_float4(K * sf[0], K * sf[1], K * sf[2], K * sf[3]);
                        fo[1] = make_float4(K * sf[4], K * sf[5], K * sf[6], K * sf[7]);
                    } else {
                        out[2 * h] = philox_word(in[2 * h], sf, K);
                        out[2 * h + 1] = philox_word(in[2 * h + 1], sf + 4, K);
                    }
                }
                if (MODE != NOISE_FIELD) stg16(d + e, make_uint4(out[0], out[1], out[2], out[3]));
            }
            done = nvec << 4;
        }

        // remainder (and the whole span when unaligned): one Philox group (<= 8 elements) per thread step
        const uint32_t r0 = done;                // first element not yet done, relative to the span
        if (r0 < n) {
            const uint32_t ea = e0 + r0, eb = e0 + n;      // absolute element range [ea, eb)
            const uint32_t g_first = ea >> 3, g_last = (eb - 1) >> 3;
            for (uint32_t g = g_first + lane; g <= g_last; g += 32u) {
                float sf[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                if (MODE == NOISE_PHILOX || MODE == NOISE_FIELD) group_gauss8(p, ig_lo, ig_hi, g, sf);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t e = 8u * g + j;
                    if (e < ea || e >= eb) continue;
                    const uint32_t rel = e - e0;
                    if (MODE == NOISE_FIELD) { fout[rel] = K * sf[j]; continue; }
                    const uint32_t v = s[rel];
                    if (MODE == NOISE_COPY) { d[rel] = (uint8_t)v; continue; }
                    if (MODE == NOISE_COMPAT) d[rel] = (uint8_t)noise_px(__uint_as_float(0x4B000000u | v) - 8388608.0f, nzp[rel]);
                    else d[rel] = (uint8_t)noise_philox_px(v, sf[j], K);
                }
            }
        }
    }
}

int launch_noise(const rod_plan* plan, int mode, const uint8_t* src, uint8_t* dst, const float* noise,
                 float* field_out, float sigma, uint64_t seed, uint64_t first_image, uint32_t offset,
                 const uint8_t* opcodes, int my_op, cudaStream_t stream, int img_lo, int img_hi) {
    if (plan->n_noise_tiles == 0) return ROD_OK;
    NoiseParams p;
    p.images = plan->d_images;
    const int t_lo = plan->noise_tile_start[img_lo], t_hi = plan->noise_tile_start[img_hi];
    if (t_hi <= t_lo) return ROD_OK;
    p.tiles = plan->d_noise_tiles + t_lo;
    p.n_tiles = t_hi - t_lo;
    p.src = src; p.dst = dst; p.noise = noise; p.field_out = field_out;
    p.sigma = sigma;
    p.key0 = (uint32_t)seed; p.key1 = (uint32_t)(seed >> 32);
    p.first_image = first_image;
    p.offset = offset;
    p.opcodes = opcodes; p.my_op = my_op;
    p.counter = plan->d_counters + (plan->launch_seq++ & 255u);
    ROD_CUDA(cudaMemsetAsync(p.counter, 0, sizeof(unsigned int), stream));
    int per_sm = 4;  // CTAs per SM the grid is sized for: 4 are resident, a longer queue evens out the tail (knob: ROD_NOISE_CTAS)
    const char* e_ctas = getenv("ROD_NOISE_CTAS");
    if (e_ctas && atoi(e_ctas) >= 1 && atoi(e_ctas) <= 64) per_sm = atoi(e_ctas);
    const int grid = grid_for(plan, (p.n_tiles + 1) / 2, per_sm);  // a CTA's 8 warps cover two spans at a time
    switch (mode) {
        case NOISE_COMPAT: noise_kernel<NOISE_COMPAT><<<grid, 256, 0, stream>>>(p); break;
        case NOISE_PHILOX: noise_kernel<NOISE_PHILOX><<<grid, 256, 0, stream>>>(p); break;
        case NOISE_COPY: noise_kernel<NOISE_COPY><<<grid, 256, 0, stream>>>(p); break;
        case NOISE_FIELD: noise_kernel<NOISE_FIELD><<<grid, 256, 0, stream>>>(p); break;
        default: return ROD_ERR_INVALID_ARG;
    }
    ROD_CUDA(cudaGetLastError());
    return ROD_OK;
}

}  // namespace rod
