// plan.cu -- batch plans: descriptor upload, tile lists, resize tables (host side of librod_b200.so).
#include <stdlib.h>

#include <map>
#include <new>

#include "rod_internal.h"
#include "rod_tables.h"

namespace rod {

thread_local int g_last_cuda_error = 0;

int grid_for(const rod_plan* plan, int n_tiles, int ctas_per_sm) {
    long long cap = (long long)plan->sm_count * ctas_per_sm;
    return (int)(n_tiles < cap ? n_tiles : cap);
}

// The device arrays of a plan come from the block cache (jpeg.cu) and go back to it: a batch driver builds and drops a plan
// per batch, and ~30 cudaMalloc / cudaFree calls per plan (each a device-wide synchronisation) cost more than its kernels.
static cudaError_t plan_alloc(const rod_plan* plan, void** p, size_t n) {
    const cudaError_t e = block_cache_alloc(plan->device, p, n);
    if (e == cudaSuccess) {
        std::lock_guard<std::mutex> lock(plan->cached_mutex);
        plan->cached_blocks.push_back({*p, n});
    }
    return e;
}
static void plan_free(const rod_plan* plan, void* p) {
    if (p == nullptr) return;
    {
        std::lock_guard<std::mutex> lock(plan->cached_mutex);
        for (size_t i = 0; i < plan->cached_blocks.size(); ++i)
            if (plan->cached_blocks[i].first == p) {
                block_cache_free(plan->device, p, plan->cached_blocks[i].second);
                plan->cached_blocks.erase(plan->cached_blocks.begin() + (long)i);
                return;
            }
    }
    cudaFree(p);   // allocated elsewhere (staging buffers of the host entry points)
}

template <typename T>
static int upload(const rod_plan* plan, const std::vector<T>& v, T** dptr) {
    *dptr = nullptr;
    if (v.empty()) return ROD_OK;
    ROD_CUDA(plan_alloc(plan, (void**)dptr, v.size() * sizeof(T)));
    ROD_CUDA(cudaMemcpy(*dptr, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return ROD_OK;
}

static void tile_starts(const std::vector<Tile>& tl, int n_images, std::vector<int>& st) {
    st.assign(n_images + 1, (int)tl.size());
    for (int t = (int)tl.size() - 1; t >= 0; --t) st[tl[t].img] = t;
    for (int i = n_images - 1; i >= 0; --i) st[i] = std::min(st[i], st[i + 1]);
}

int ensure_lowres_tables(rod_plan* plan, double factor) {
    if (plan->d_shapes != nullptr && plan->lowres_factor == factor) return ROD_OK;
    if (!(factor > 0.0) || factor > 1.0) return ROD_ERR_UNSUPPORTED;
    std::vector<uint32_t> blob;
    std::vector<DevShape> shapes(plan->shapes.size());
    bool all_identity = true;
    int max_rows = 1, max_cols = 1, max_src_rows = 1;
    size_t x2_smem = 0;
    // tuning knobs (benchmarks only): CTA size, forced strip height, shared-memory budget per CTA
    const char* e_thr = getenv("ROD_X2_THREADS");
    const char* e_rows = getenv("ROD_X2_STRIP");
    const int x2_threads = (e_thr && atoi(e_thr) == 256) ? 256 : 128;
    const int x2_force_rows = e_rows ? atoi(e_rows) : 0;
    const size_t x2_max_smem = (x2_threads == 128) ? 54 * 1024 : 100 * 1024;
    plan->lowres_x2_threads = x2_threads;
    for (size_t s = 0; s < plan->shapes.size(); ++s) {
        const int h = plan->shapes[s].h, w = plan->shapes[s].w;
        if (!build_lowres_shape(h, w, factor, kMaxAreaTaps, blob, &shapes[s])) return ROD_ERR_UNSUPPORTED;
        DevShape& sh = shapes[s];
        if (sh.lin_identity) continue;
        all_identity = false;
        x2_smem = std::max(x2_smem, choose_strip_rows(&sh, blob.data(), x2_max_smem, x2_threads, x2_force_rows));
        if (sh.strip_rows > 0) continue;
        int rows, cols, srows;
        lowres_tile_footprint(sh, blob.data(), kLowresTH, kLowresTWB, &rows, &cols, &srows);
        max_rows = std::max(max_rows, rows);
        max_cols = std::max(max_cols, cols);
        max_src_rows = std::max(max_src_rows, srows);
    }
    if (blob.empty()) blob.push_back(0);
    // warp-marching kernel: one tile = (band of rows, strip of 30 chunks); band height from the amount of work
    // (aim at >= 4 tiles per resident warp; 16 warps per SM)
    const char* e_march = getenv("ROD_X2_MARCH");
    const bool march = !(e_march && atoi(e_march) == 0);
    auto march_image = [&](int i) {
        const DevImage& im = plan->h_images[i];
        return shapes[im.shape_id].x2w != 0 && (im.src_pitch & 3) == 0 && (im.src_off & 3) == 0;
    };
    auto al8_image = [&](int i) {
        const DevImage& im = plan->h_images[i];
        return (im.src_pitch & 7) == 0 && (im.src_off & 7) == 0 && (im.w & 7) == 0;
    };
    auto n_strips = [&](int w) { return ((w + 7) / 8 + 29) / 30; };
    int band_rows = 0;
    if (march) {
        long long strip_rows = 0;
        for (int i = 0; i < plan->n_images; ++i)
            if (march_image(i)) strip_rows += (long long)plan->h_images[i].h * n_strips(plan->h_images[i].w);
        const char* e_band = getenv("ROD_X2_BAND");
        band_rows = e_band ? atoi(e_band) : (int)(strip_rows / (80LL * plan->sm_count));  // ~5 tiles per resident warp (measured best on B200)
        band_rows = std::max(24, std::min(256, band_rows & ~7));  // <= kX2wMaxBandRows (per-warp shared tables)
    }
    // packed-integer kernel (exact 2x in both axes): additionally needs 4-byte aligned destination rows (32 / 64-bit stores)
    const char* e_packed = getenv("ROD_X2_PACKED");
    const bool packed_on = !(e_packed && atoi(e_packed) == 0);
    auto packed_image = [&](int i) {
        const DevImage& im = plan->h_images[i];
        return packed_on && shapes[im.shape_id].x2p != 0 && (im.dst_pitch & 3) == 0 && (im.dst_off & 3) == 0;
    };
    auto packed_class = [&](int i) {  // 0: 16-byte copy units, 1: 8-byte, 2: 4-byte
        const DevImage& im = plan->h_images[i];
        const uint64_t a = (uint64_t)im.src_pitch | (uint64_t)im.dst_pitch | im.src_off | im.dst_off;
        if ((im.w & 15) == 0 && (a & 15) == 0) return 0;
        if ((im.w & 7) == 0 && (a & 7) == 0) return 1;
        return 2;
    };
    // float-tap staged kernel (exact-2x width, general y taps): 4-byte aligned destination rows, and a y scale small
    // enough for its ring of 8 source rows
    const char* e_x2f = getenv("ROD_X2_FLOAT_STAGED");
    const bool x2f_on = !(e_x2f && atoi(e_x2f) == 0);
    auto float_image = [&](int i) {
        const DevImage& im = plan->h_images[i];
        const DevShape& sh = shapes[im.shape_id];
        return x2f_on && sh.x2w != 0 && sh.area_mode == AREA_GENERAL && sh.ay_packed && (im.dst_pitch & 3) == 0 &&
               (im.dst_off & 3) == 0 && (sh.h <= 8 || 4LL * sh.h <= 9LL * sh.nh);
    };
    const char* e_x2h = getenv("ROD_X2_REGULAR");
    const bool x2h_on = !(e_x2h && atoi(e_x2h) == 0);
    auto regular_image = [&](int i) { return x2h_on && float_image(i) && shapes[plan->h_images[i].shape_id].x2h != 0; };
    // odd-width staged kernel: regular 3-tap / one-slip shapes whose y scale fits the ring of 8 source rows
    const char* e_x2g = getenv("ROD_X2_ODD_STAGED");
    const bool x2g_on = !(e_x2g && atoi(e_x2g) == 0);
    auto odd_image = [&](int i) {
        const DevShape& sh = shapes[plan->h_images[i].shape_id];
        return x2g_on && sh.x2g != 0 && (sh.h <= 8 || 4LL * sh.h <= 9LL * sh.nh);
    };
    const char* e_x2i = getenv("ROD_X2_ODD_REGULAR");
    const bool x2i_on = !(e_x2i && atoi(e_x2i) == 0);
    std::vector<Tile> gen_tiles, x2_tiles, x2w_tiles, x2w4_tiles, x2_rest_tiles, x2p_tiles[3], x2f_tiles[3], x2h_tiles[3], x2g_tiles, x2i_tiles[2], gen_all;
    build_strip_tiles(plan->h_images, shapes, false, kLowresTH, kLowresTWB, gen_all);
    for (const Tile& t : gen_all)
        if (!odd_image(t.img)) gen_tiles.push_back(t);
    build_strip_tiles(plan->h_images, shapes, true, kLowresTH, kLowresTWB, x2_tiles);
    int band_rows_odd = 0;
    {
        long long strip_rows = 0;
        for (int i = 0; i < plan->n_images; ++i)
            if (odd_image(i)) strip_rows += (long long)plan->h_images[i].h * n_strips(plan->h_images[i].w);
        const char* e_band = getenv("ROD_X2G_BAND_DIV");  // tiles per SM the band height aims at (benchmark knob)
        const long long div = e_band && atoi(e_band) >= 4 ? atoi(e_band) : 60;
        band_rows_odd = std::max(24, std::min(256, (int)(strip_rows / (div * plan->sm_count)) & ~7));
    }
    for (int i = 0; i < plan->n_images; ++i) {
        const DevShape& sh = shapes[plan->h_images[i].shape_id];
        if (odd_image(i)) {
            std::vector<Tile>& list = (x2i_on && sh.x2i != 0) ? x2i_tiles[sh.x2i - 1] : x2g_tiles;
            for (int y = 0; y < plan->h_images[i].h; y += band_rows_odd)
                for (int st = 0; st < n_strips(plan->h_images[i].w); ++st)
                    list.push_back(Tile{i, y, std::min(plan->h_images[i].h, y + band_rows_odd), st});
            continue;
        }
        if (sh.strip_rows <= 0) continue;
        if (march && march_image(i)) {
            for (int y = 0; y < plan->h_images[i].h; y += band_rows)
                for (int st = 0; st < n_strips(plan->h_images[i].w); ++st)
                    (packed_image(i) ? x2p_tiles[packed_class(i)] : regular_image(i) ? x2h_tiles[packed_class(i)] : float_image(i) ? x2f_tiles[packed_class(i)] : al8_image(i) ? x2w_tiles : x2w4_tiles).push_back(Tile{i, y, std::min(plan->h_images[i].h, y + band_rows), st});
        } else {
            for (int y = 0; y < plan->h_images[i].h; y += sh.strip_rows) x2_rest_tiles.push_back(Tile{i, y, 0, 0});
        }
    }
    void* old[] = {plan->d_shapes, plan->d_tab, plan->d_lowres_tiles, plan->d_lowres_x2_tiles, plan->d_lowres_x2w_tiles, plan->d_lowres_x2w4_tiles,
                   plan->d_lowres_x2_rest_tiles, plan->d_lowres_x2p_tiles[0], plan->d_lowres_x2p_tiles[1], plan->d_lowres_x2p_tiles[2],
                   plan->d_lowres_x2f_tiles[0], plan->d_lowres_x2f_tiles[1], plan->d_lowres_x2f_tiles[2], plan->d_lowres_x2g_tiles,
                   plan->d_lowres_x2h_tiles[0], plan->d_lowres_x2h_tiles[1], plan->d_lowres_x2h_tiles[2],
                   plan->d_lowres_x2i_tiles[0], plan->d_lowres_x2i_tiles[1]};
    for (void* q : old) plan_free(plan, q);
    plan->d_shapes = nullptr; plan->d_tab = nullptr; plan->d_lowres_tiles = nullptr; plan->d_lowres_x2_tiles = nullptr;
    plan->d_lowres_x2w_tiles = nullptr; plan->d_lowres_x2w4_tiles = nullptr; plan->d_lowres_x2_rest_tiles = nullptr;
    int rc = upload(plan, shapes, &plan->d_shapes);
    plan->d_lowres_x2g_tiles = nullptr;
    if (rc == ROD_OK) rc = upload(plan, x2g_tiles, &plan->d_lowres_x2g_tiles);
    for (int u = 0; u < 2; ++u) {
        plan->d_lowres_x2i_tiles[u] = nullptr;
        if (rc == ROD_OK) rc = upload(plan, x2i_tiles[u], &plan->d_lowres_x2i_tiles[u]);
    }
    for (int u = 0; u < 3; ++u) {
        plan->d_lowres_x2p_tiles[u] = nullptr;
        if (rc == ROD_OK) rc = upload(plan, x2p_tiles[u], &plan->d_lowres_x2p_tiles[u]);
        plan->d_lowres_x2f_tiles[u] = nullptr;
        if (rc == ROD_OK) rc = upload(plan, x2f_tiles[u], &plan->d_lowres_x2f_tiles[u]);
        plan->d_lowres_x2h_tiles[u] = nullptr;
        if (rc == ROD_OK) rc = upload(plan, x2h_tiles[u], &plan->d_lowres_x2h_tiles[u]);
    }
    if (rc == ROD_OK) rc = upload(plan, blob, &plan->d_tab);
    if (rc == ROD_OK) rc = upload(plan, gen_tiles, &plan->d_lowres_tiles);
    if (rc == ROD_OK) rc = upload(plan, x2_tiles, &plan->d_lowres_x2_tiles);
    if (rc == ROD_OK) rc = upload(plan, x2w_tiles, &plan->d_lowres_x2w_tiles);
    if (rc == ROD_OK) rc = upload(plan, x2w4_tiles, &plan->d_lowres_x2w4_tiles);
    if (rc == ROD_OK) rc = upload(plan, x2_rest_tiles, &plan->d_lowres_x2_rest_tiles);
    if (rc != ROD_OK) return rc;
    plan->n_lowres_tiles = (int)gen_tiles.size();
    plan->n_lowres_x2_tiles = (int)x2_tiles.size();
    plan->n_lowres_x2w_tiles = (int)x2w_tiles.size();
    plan->n_lowres_x2w4_tiles = (int)x2w4_tiles.size();
    plan->n_lowres_x2g_tiles = (int)x2g_tiles.size();
    tile_starts(x2g_tiles, plan->n_images, plan->lowres_x2g_tile_start);
    for (int u = 0; u < 2; ++u) {
        plan->n_lowres_x2i_tiles[u] = (int)x2i_tiles[u].size();
        tile_starts(x2i_tiles[u], plan->n_images, plan->lowres_x2i_tile_start[u]);
    }
    for (int u = 0; u < 3; ++u) {
        plan->n_lowres_x2p_tiles[u] = (int)x2p_tiles[u].size();
        tile_starts(x2p_tiles[u], plan->n_images, plan->lowres_x2p_tile_start[u]);
        plan->n_lowres_x2f_tiles[u] = (int)x2f_tiles[u].size();
        tile_starts(x2f_tiles[u], plan->n_images, plan->lowres_x2f_tile_start[u]);
        plan->n_lowres_x2h_tiles[u] = (int)x2h_tiles[u].size();
        tile_starts(x2h_tiles[u], plan->n_images, plan->lowres_x2h_tile_start[u]);
    }
    tile_starts(x2w4_tiles, plan->n_images, plan->lowres_x2w4_tile_start);
    plan->n_lowres_x2_rest_tiles = (int)x2_rest_tiles.size();
    tile_starts(gen_tiles, plan->n_images, plan->lowres_tile_start);
    tile_starts(x2_tiles, plan->n_images, plan->lowres_x2_tile_start);
    tile_starts(x2w_tiles, plan->n_images, plan->lowres_x2w_tile_start);
    tile_starts(x2_rest_tiles, plan->n_images, plan->lowres_x2_rest_tile_start);
    plan->lowres_x2w_band_rows = band_rows;
    plan->lowres_x2_smem = x2_smem;
    plan->lowres_factor = factor;
    plan->lowres_all_identity = all_identity;
    plan->lowres_all_x2w = true;
    for (const DevShape& sh : shapes)
        if (!sh.lin_identity && !sh.x2w) plan->lowres_all_x2w = false;
    for (int i = 0; i < plan->n_images; ++i)  // 32-bit source loads in the fused kernel
        if ((plan->h_images[i].src_pitch & 3) != 0 || (plan->h_images[i].src_off & 3) != 0) plan->lowres_all_x2w = false;
    plan->lowres_half_rows = max_rows;
    plan->lowres_half_cols = max_cols;
    plan->lowres_src_rows = max_src_rows;
    return ROD_OK;
}

int ensure_letterbox_tables(rod_plan* plan, int out_h, int out_w) {
    if (plan->d_lb != nullptr && plan->lb_out_h == out_h && plan->lb_out_w == out_w) return ROD_OK;
    if (out_h < 1 || out_w < 1) return ROD_ERR_INVALID_ARG;
    std::vector<uint32_t> blob;
    std::vector<DevLetterbox> lbs(plan->shapes.size());
    for (size_t s = 0; s < plan->shapes.size(); ++s) {
        DevLetterbox g;
        memset(&g, 0, sizeof(g));
        g.h = plan->shapes[s].h; g.w = plan->shapes[s].w;
        letterbox_geometry(g.h, g.w, out_h, out_w, &g.new_h, &g.new_w, &g.top, &g.left);
        if (g.new_h < 1 || g.new_w < 1 || g.h > 65535 || g.w > 65535) return ROD_ERR_UNSUPPORTED;
        g.identity = (g.new_h == g.h && g.new_w == g.w);
        g.area2 = (!g.identity && g.h == 2 * g.new_h && g.w == 2 * g.new_w);
        if (!g.identity && !g.area2) {
            LinearAxis lx = build_linear_axis(g.w, g.new_w, true);
            LinearAxis ly = build_linear_axis(g.h, g.new_h, false);
            std::vector<uint32_t> ys(g.new_h);
            for (int y = 0; y < g.new_h; ++y) ys[y] = (uint32_t)ly.s0[y] | ((uint32_t)ly.s1[y] << 16);
            g.lx_s0 = blob_push(blob, lx.s0);
            g.lx_a = blob_push(blob, lx.coef);
            g.ly_s = blob_push(blob, ys);
            g.ly_b = blob_push(blob, ly.coef);
            std::vector<uint32_t> pack((size_t)2 * out_w);
            for (int X = 0; X < out_w; ++X) {
                const int cx = X - g.left;
                const bool in = cx >= 0 && cx < g.new_w;
                pack[2 * X] = in ? (uint32_t)(3 * lx.s0[cx]) : 0xFFFFFFFFu;
                pack[2 * X + 1] = in ? lx.coef[cx] : 0u;
            }
            g.lx_pack = blob_push(blob, pack);
            // The fused kernel's lane l resamples output columns 2l, 2l+1 (+64 per step) from six-byte windows that lie
            // 2 * 3 * (w / new_w) bytes apart.  If p lanes further the window is a multiple of 128 bytes away (within a word),
            // lanes l, l+p, l+2p, ... hit the same bank: the kernel then shifts every such group by one step.
            const double stride_words = 2.0 * 3.0 * ((double)g.w / (double)g.new_w) / 4.0;
            g.lane_group = 0;
            for (int pp = 2; pp <= 16; ++pp) {
                const double x = pp * stride_words / 32.0;
                if (fabs(x - nearbyint(x)) * 32.0 < 1.0) { g.lane_group = pp; break; }
            }
        }
        lbs[s] = g;
    }
    if (blob.empty()) blob.push_back(0);
    std::vector<Tile> tiles;
    for (int i = 0; i < plan->n_images; ++i)
        for (int y = 0; y < out_h; y += kLbTH)
            for (int x = 0; x < out_w; x += kLbTW) tiles.push_back(Tile{i, y, x, 0});
    if (plan->d_lb) { plan_free(plan, plan->d_lb); plan->d_lb = nullptr; }
    if (plan->d_lb_tab) { plan_free(plan, plan->d_lb_tab); plan->d_lb_tab = nullptr; }
    if (plan->d_lb_tiles) { plan_free(plan, plan->d_lb_tiles); plan->d_lb_tiles = nullptr; }
    int rc = upload(plan, lbs, &plan->d_lb);
    if (rc != ROD_OK) return rc;
    rc = upload(plan, blob, &plan->d_lb_tab);
    if (rc != ROD_OK) return rc;
    rc = upload(plan, tiles, &plan->d_lb_tiles);
    if (rc != ROD_OK) return rc;
    plan->n_lb_tiles = (int)tiles.size();
    plan->lb_all_linear = true;
    for (const DevLetterbox& g : lbs)
        if (g.identity || g.area2) plan->lb_all_linear = false;
    plan->lb_out_h = out_h;
    plan->lb_out_w = out_w;
    return ROD_OK;
}

}  // namespace rod

using namespace rod;

extern "C" int rod_plan_create(const rod_image_desc* images, int n_images, rod_plan** out_plan) {
    if (out_plan == nullptr) return ROD_ERR_INVALID_ARG;
    *out_plan = nullptr;
    if (images == nullptr || n_images < 1) return ROD_ERR_INVALID_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return ROD_ERR_NO_DEVICE;
    }
    rod_plan* plan = new (std::nothrow) rod_plan();
    if (plan == nullptr) return ROD_ERR_OOM;
    cudaError_t e = cudaGetDevice(&plan->device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&plan->sm_count, cudaDevAttrMultiProcessorCount, plan->device);
    if (e != cudaSuccess) { delete plan; return cuda_fail(e); }
    plan->n_images = n_images;
    plan->descs.assign(images, images + n_images);
    plan->h_images.resize(n_images);
    std::map<std::pair<int, int>, int> shape_ids;
    uint64_t elem = 0;
    for (int i = 0; i < n_images; ++i) {
        const rod_image_desc& d = images[i];
        if (d.height < 1 || d.width < 1 || d.width > 65535 || d.height > 65535 || d.src_pitch < 3LL * d.width ||
            d.dst_pitch < 3LL * d.width || (uint64_t)d.height * d.width * 3 > 0x7FFFFFFFull) {
            delete plan;
            return ROD_ERR_INVALID_ARG;
        }
        DevImage& im = plan->h_images[i];
        im.src_off = d.src_offset; im.dst_off = d.dst_offset;
        im.src_pitch = d.src_pitch; im.dst_pitch = d.dst_pitch;
        im.h = d.height; im.w = d.width;
        im.elem_base = elem;
        im.contiguous = (d.src_pitch == 3LL * d.width && d.dst_pitch == 3LL * d.width) ? 1 : 0;
        elem += (uint64_t)d.height * d.width * 3;
        auto key = std::make_pair(d.height, d.width);
        auto it = shape_ids.find(key);
        if (it == shape_ids.end()) {
            it = shape_ids.emplace(key, (int)plan->shapes.size()).first;
            plan->shapes.push_back(HostShape{d.height, d.width});
        }
        im.shape_id = it->second;
        plan->max_w = std::max(plan->max_w, d.width);
        plan->all_contiguous = plan->all_contiguous && im.contiguous;
        plan->src_min_offset = (i == 0) ? d.src_offset : std::min<uint64_t>(plan->src_min_offset, d.src_offset);
        plan->src_extent = std::max<uint64_t>(plan->src_extent, d.src_offset + (uint64_t)(d.height - 1) * d.src_pitch + 3ull * d.width);
        plan->dst_extent = std::max<uint64_t>(plan->dst_extent, d.dst_offset + (uint64_t)(d.height - 1) * d.dst_pitch + 3ull * d.width);
    }
    plan->payload_bytes = elem;
    std::vector<Tile> nt, bt;
    build_noise_tiles(plan->h_images, kNoiseSpan, nt);
    build_blur_tiles(plan->h_images, kBlurRowsPerTile, bt);
    auto starts = [&](const std::vector<Tile>& tl, std::vector<int>& st) {
        st.assign(n_images + 1, (int)tl.size());
        for (int t = (int)tl.size() - 1; t >= 0; --t) st[tl[t].img] = t;
        for (int i = n_images - 1; i >= 0; --i) st[i] = std::min(st[i], st[i + 1]);
    };
    starts(nt, plan->noise_tile_start);
    starts(bt, plan->blur_tile_start);
    for (int i = 1; i < n_images; ++i) {
        const rod_image_desc& a = images[i - 1];
        const rod_image_desc& b = images[i];
        if (b.src_offset < a.src_offset + (uint64_t)a.height * a.src_pitch - (a.src_pitch - 3ull * a.width) ||
            b.dst_offset < a.dst_offset + (uint64_t)a.height * a.dst_pitch - (a.dst_pitch - 3ull * a.width))
            plan->monotonic = false;
    }
    plan->n_noise_tiles = (int)nt.size();
    plan->n_blur_tiles = (int)bt.size();
    if (plan_alloc(plan, (void**)&plan->d_counters, kCounterRing * sizeof(unsigned int)) != cudaSuccess) { rod_plan_destroy(plan); return ROD_ERR_OOM; }
    int rc = upload(plan, plan->h_images, &plan->d_images);
    if (rc == ROD_OK) rc = upload(plan, nt, &plan->d_noise_tiles);
    if (rc == ROD_OK) rc = upload(plan, bt, &plan->d_blur_tiles);
    if (rc != ROD_OK) { rod_plan_destroy(plan); return rc; }
    *out_plan = plan;
    return ROD_OK;
}

// Installs (kernel != NULL) or removes (NULL) a general k x k float32 blur kernel: while installed, ROD_OP_BLUR runs
// the 2-D filter (filter2d.cu) instead of the horizontal box.
extern "C" int rod_set_blur_kernel(rod_plan* plan, const float* kernel, int k) {
    if (plan == nullptr) return ROD_ERR_INVALID_ARG;
    if (kernel == nullptr) {
        plan->f2d_ntaps = 0;
        plan->f2d_k = 0;
        return ROD_OK;
    }
    if (k < 1 || (k & 1) == 0) return ROD_ERR_INVALID_ARG;
    if (k * k >= kF2dMaxElems) return ROD_ERR_UNSUPPORTED;  // OpenCV's DFT path: not reproducible bit-exactly
    std::vector<float4> taps;
    for (int dy = 0; dy < k; ++dy)
        for (int dx = 0; dx < k; ++dx) {
            const float w = kernel[dy * k + dx];
            if (w == 0.0f) continue;
            float4 t;
            const int dxb = 3 * dx;
            memcpy(&t.x, &dy, 4);
            memcpy(&t.y, &dxb, 4);
            t.z = w;
            t.w = 0.f;
            taps.push_back(t);
        }
    if (taps.empty()) return ROD_ERR_INVALID_ARG;
    if (plan->d_f2d_taps) { plan_free(plan, plan->d_f2d_taps); plan->d_f2d_taps = nullptr; }
    int rc = upload(plan, taps, &plan->d_f2d_taps);
    if (rc != ROD_OK) return rc;
    if (plan->d_f2d_tiles == nullptr) {
        std::vector<Tile> tiles;
        build_grid_tiles(plan->h_images, kF2dTH, kF2dTWB, tiles);
        rc = upload(plan, tiles, &plan->d_f2d_tiles);
        if (rc != ROD_OK) return rc;
        plan->n_f2d_tiles = (int)tiles.size();
        tile_starts(tiles, plan->n_images, plan->f2d_tile_start);
    }
    plan->f2d_ntaps = (int)taps.size();
    plan->f2d_k = k;
    return ROD_OK;
}

extern "C" void rod_plan_destroy(rod_plan* plan) {
    if (plan == nullptr) return;
    if (plan->inner) rod_plan_destroy(plan->inner);
    // the arrays go back to a cache and may be handed to the next plan at once: kernels of this plan that are still running
    // must be done first (what the implicit synchronisation of cudaFree used to guarantee, once instead of thirty times)
    cudaDeviceSynchronize();
    if (plan->d_patch_clean) cudaFree(plan->d_patch_clean);
    if (plan->d_patch_corrupted) cudaFree(plan->d_patch_corrupted);
    void* ptrs[] = {plan->d_images, plan->d_noise_tiles, plan->d_blur_tiles, plan->d_lowres_tiles, plan->d_lowres_x2_tiles,
                    plan->d_lowres_x2w_tiles, plan->d_lowres_x2w4_tiles, plan->d_lowres_x2_rest_tiles, plan->d_shapes,
                    plan->d_lowres_x2p_tiles[0], plan->d_lowres_x2p_tiles[1], plan->d_lowres_x2p_tiles[2],
                    plan->d_lowres_x2f_tiles[0], plan->d_lowres_x2f_tiles[1], plan->d_lowres_x2f_tiles[2], plan->d_lowres_x2g_tiles,
                    plan->d_lowres_x2h_tiles[0], plan->d_lowres_x2h_tiles[1], plan->d_lowres_x2h_tiles[2],
                    plan->d_lowres_x2i_tiles[0], plan->d_lowres_x2i_tiles[1],
                    plan->d_tab, plan->d_lb, plan->d_lb_tab, plan->d_lb_tiles, plan->d_scratch, plan->d_stage_src,
                    plan->d_f2d_taps, plan->d_f2d_tiles, plan->d_counters,
                    plan->d_stage_dst, plan->d_stage_noise, plan->d_stage_ops};
    for (void* p : ptrs) plan_free(plan, p);
    for (cudaStream_t s : plan->streams)
        if (s) cudaStreamDestroy(s);
    for (cudaStream_t s : plan->aux_streams)
        if (s) cudaStreamDestroy(s);
    for (cudaStream_t s : plan->lr_streams)
        if (s) cudaStreamDestroy(s);
    if (plan->lr_ev_fork) cudaEventDestroy(plan->lr_ev_fork);
    for (cudaEvent_t e : plan->lr_ev_join)
        if (e) cudaEventDestroy(e);
    if (plan->ev_fork) cudaEventDestroy(plan->ev_fork);
    for (cudaEvent_t e : plan->ev_join)
        if (e) cudaEventDestroy(e);
    delete plan;
}

extern "C" int rod_plan_num_images(const rod_plan* plan) { return plan ? plan->n_images : 0; }
extern "C" uint64_t rod_plan_payload_bytes(const rod_plan* plan) { return plan ? plan->payload_bytes : 0; }
