// lowres.cu -- a4+a5: fused INTER_AREA downscale + 8-bit INTER_LINEAR upscale (sm_100a).
//
// Reference: scripts/augmentations.py:41-45 (apply_lowres): cv2.resize(img,(nw,nh),INTER_AREA)
// followed by cv2.resize(small,(w,h),INTER_LINEAR).  The low-resolution intermediate of one
// output tile (plus its one-pixel apron) is produced in shared memory and consumed from
// there: it never exists in HBM.  Per tile of kLowresTH x kLowresTW output pixels:
//   phase B : low-res tile (u8)  <- source pixels (area_value: exact OpenCV arithmetic)
//   phase C1: horizontal fixed-point pass, (S0*a0 + S1*a1) >> 4 as u16 per low-res row
//   phase C2: vertical pass + pack, 16 output bytes per thread, 128-bit stores at the
//             destination's 16-byte phase (edges bytewise)
#include "rod_internal.h"

namespace rod {

struct LowresParams {
    const DevImage* images;
    const Tile* tiles;
    int n_tiles;
    const DevShape* shapes;
    const uint32_t* tab;
    const uint8_t* src;
    uint8_t* dst;
    const uint8_t* opcodes;
    int half_rows, half_cols;  // allocation (worst case over the plan) of the low-res tile
};

__device__ __forceinline__ void stg16_lr(void* p, uint4 v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w)
                 : "memory");
}

__global__ void __launch_bounds__(256) lowres_kernel(LowresParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int half_pitch = (p.half_cols * 3 + 15) & ~15;
    const int hx_pitch = kLowresTW * 3 + 8;  // u16 elements; +8 keeps rows 16-byte aligned
    uint8_t* half = smem;
    uint16_t* hx = reinterpret_cast<uint16_t*>(smem + (size_t)p.half_rows * half_pitch);

    for (int ti = blockIdx.x; ti < p.n_tiles; ti += gridDim.x) {
        const Tile t = p.tiles[ti];
        if (p.opcodes != nullptr && p.opcodes[t.img] != ROD_OP_LOWRES) continue;
        const DevImage im = p.images[t.img];
        const DevShape sh = p.shapes[im.shape_id];
        const uint8_t* simg = p.src + im.src_off;
        uint8_t* dimg = p.dst + im.dst_off;
        const int y0 = t.a, x0 = t.b;
        const int th = min(kLowresTH, im.h - y0), tw = min(kLowresTW, im.w - x0);
        const int tw3 = tw * 3;

        if (sh.lin_identity) {  // factor maps (h, w) onto itself: both resizes are copies
            for (int idx = threadIdx.x; idx < th * tw3; idx += blockDim.x) {
                const int r = idx / tw3, o = idx - r * tw3;
                dimg[(int64_t)(y0 + r) * im.dst_pitch + x0 * 3 + o] = simg[(int64_t)(y0 + r) * im.src_pitch + x0 * 3 + o];
            }
            continue;
        }
        const int32_t* lx_s0 = reinterpret_cast<const int32_t*>(p.tab + sh.lx_s0);
        const uint32_t* lx_a = p.tab + sh.lx_a;
        const uint32_t* ly_s = p.tab + sh.ly_s;
        const uint32_t* ly_b = p.tab + sh.ly_b;
        // low-res rows / cols this tile reads (tables are monotonic)
        const int j_lo = (int)(ly_s[y0] & 0xFFFFu), j_hi = (int)(ly_s[y0 + th - 1] >> 16);
        const int i_lo = lx_s0[x0], i_hi = min(lx_s0[x0 + tw - 1] + 1, sh.nw - 1);
        const int nj = j_hi - j_lo + 1, ni3 = (i_hi - i_lo + 1) * 3;

        // ---- phase B: low-res tile
        for (int idx = threadIdx.x; idx < nj * ni3; idx += blockDim.x) {
            const int jr = idx / ni3, o = idx - jr * ni3;
            const int ir = o / 3, c = o - 3 * ir;
            half[jr * half_pitch + o] = (uint8_t)area_value(simg, im.src_pitch, sh, p.tab, j_lo + jr, i_lo + ir, c);
        }
        __syncthreads();
        // ---- phase C1: horizontal pass
        for (int idx = threadIdx.x; idx < nj * tw3; idx += blockDim.x) {
            const int jr = idx / tw3, o = idx - jr * tw3;
            const int xr = o / 3, c = o - 3 * xr;
            const int s0 = lx_s0[x0 + xr];
            const int s1 = min(s0 + 1, sh.nw - 1);
            const uint8_t* hr = half + jr * half_pitch;
            hx[jr * hx_pitch + o] = (uint16_t)linear_h4(hr[(s0 - i_lo) * 3 + c], hr[(s1 - i_lo) * 3 + c], lx_a[x0 + xr]);
        }
        __syncthreads();
        // ---- phase C2: vertical pass, 16 bytes per thread at the destination's 16-byte phase
        {
            const int64_t row0 = (int64_t)y0 * im.dst_pitch + x0 * 3;
            const int chunks_max = (tw3 + 15 + 15) >> 4;  // chunks per row for any phase
            for (int idx = threadIdx.x; idx < th * chunks_max; idx += blockDim.x) {
                const int r = idx / chunks_max, j = idx - r * chunks_max;
                uint8_t* drow = dimg + row0 + (int64_t)r * im.dst_pitch;
                const int shift = (int)((uintptr_t)drow & 15);
                const int lo = 16 * j - shift;
                if (lo >= tw3) continue;
                const uint32_t ys = ly_s[y0 + r], yb = ly_b[y0 + r];
                const uint16_t* h0 = hx + ((int)(ys & 0xFFFFu) - j_lo) * hx_pitch;
                const uint16_t* h1 = hx + ((int)(ys >> 16) - j_lo) * hx_pitch;
                uint32_t out[4];
                if (lo >= 0 && lo + 16 <= tw3) {
                    if ((lo & 7) == 0) {
                        const uint4 a0 = *reinterpret_cast<const uint4*>(h0 + lo);
                        const uint4 a1 = *reinterpret_cast<const uint4*>(h0 + lo + 8);
                        const uint4 b0 = *reinterpret_cast<const uint4*>(h1 + lo);
                        const uint4 b1 = *reinterpret_cast<const uint4*>(h1 + lo + 8);
                        const uint32_t u0[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                        const uint32_t u1[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            const uint32_t v0 = linear_v(u0[2 * g] & 0xFFFFu, u1[2 * g] & 0xFFFFu, yb);
                            const uint32_t v1 = linear_v(u0[2 * g] >> 16, u1[2 * g] >> 16, yb);
                            const uint32_t v2 = linear_v(u0[2 * g + 1] & 0xFFFFu, u1[2 * g + 1] & 0xFFFFu, yb);
                            const uint32_t v3 = linear_v(u0[2 * g + 1] >> 16, u1[2 * g + 1] >> 16, yb);
                            out[g] = v0 | (v1 << 8) | (v2 << 16) | (v3 << 24);
                        }
                    } else {
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            uint32_t o = 0;
#pragma unroll
                            for (int b = 0; b < 4; ++b) {
                                const int i = lo + 4 * g + b;
                                o |= linear_v(h0[i], h1[i], yb) << (8 * b);
                            }
                            out[g] = o;
                        }
                    }
                    stg16_lr(drow + lo, make_uint4(out[0], out[1], out[2], out[3]));
                } else {
                    for (int b = 0; b < 16; ++b) {
                        const int i = lo + b;
                        if (i >= 0 && i < tw3) drow[i] = (uint8_t)linear_v(h0[i], h1[i], yb);
                    }
                }
            }
        }
        __syncthreads();  // smem is rewritten by the next tile
    }
}

int launch_lowres(const rod_plan* plan, const uint8_t* src, uint8_t* dst, const uint8_t* opcodes,
                  cudaStream_t stream, int img_lo, int img_hi) {
    if (plan->n_lowres_tiles == 0) return ROD_OK;
    LowresParams p;
    p.images = plan->d_images;
    const int t_lo = plan->lowres_tile_start[img_lo], t_hi = plan->lowres_tile_start[img_hi];
    if (t_hi <= t_lo) return ROD_OK;
    p.tiles = plan->d_lowres_tiles + t_lo;
    p.n_tiles = t_hi - t_lo;
    p.shapes = plan->d_shapes;
    p.tab = plan->d_tab;
    p.src = src; p.dst = dst; p.opcodes = opcodes;
    p.half_rows = plan->lowres_half_rows;
    p.half_cols = plan->lowres_half_cols;
    const int half_pitch = (p.half_cols * 3 + 15) & ~15;
    const size_t smem = (size_t)p.half_rows * half_pitch + (size_t)p.half_rows * (kLowresTW * 3 + 8) * 2;
    if (smem > 227 * 1024) return ROD_ERR_UNSUPPORTED;
    ROD_CUDA(cudaFuncSetAttribute(lowres_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int ctas_per_sm = (int)((220 * 1024) / (smem + 1024));
    ctas_per_sm = ctas_per_sm < 1 ? 1 : (ctas_per_sm > 8 ? 8 : ctas_per_sm);
    const int grid = grid_for(plan, p.n_tiles, ctas_per_sm);
    lowres_kernel<<<grid, 256, smem, stream>>>(p);
    ROD_CUDA(cudaGetLastError());
    return ROD_OK;
}

}  // namespace rod
