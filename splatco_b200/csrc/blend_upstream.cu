// Same-GPU COMPARATOR, not the product path: a plain restatement of the upstream rasterizer's blend structure
// (SURVEY.md Appendix A.4 / A.5 -- Inria diff-gaussian-rasterization as forked by Scaffold-GS; the reference's own
// source is in the absent submodules.zip, so this is builder-authored from the published algorithm): one CTA per 16x16
// tile, one thread per pixel in row-major order, batches of 256 instances staged through shared memory, every pixel
// evaluates every staged instance, `__syncthreads_count(done) == 256` ends a tile, and the backward adds its nine
// per-(pixel, instance) gradients with one global fp32 atomicAdd each.  bench.py times it next to the product kernels
// on the same workspaces (`gpu_baseline.upstream_structure`); tests/test_raster_gpu.py checks it against the oracle.
#include "common.cuh"

namespace splatco {

constexpr int UP_THREADS = TILE * TILE;

__global__ void __launch_bounds__(UP_THREADS)
blend_fwd_upstream_kernel(const int2 *__restrict__ ranges, const uint32_t *__restrict__ point_list, const float4 *__restrict__ rec,
                          int W, int H, int gx, const float *__restrict__ bg, float *__restrict__ out_color,
                          float *__restrict__ final_T, int32_t *__restrict__ n_contrib) {
    __shared__ float2 s_xy[UP_THREADS];
    __shared__ float4 s_co[UP_THREADS];
    __shared__ float s_rgb[UP_THREADS][3];
    const int tile = blockIdx.y * gx + blockIdx.x;
    const int px = blockIdx.x * TILE + (threadIdx.x & 15), py = blockIdx.y * TILE + (threadIdx.x >> 4);
    const bool inside = px < W && py < H;
    const float fx = (float)px, fy = (float)py;
    const int2 range = ranges[tile];
    int todo = range.y - range.x;
    bool done = !inside;
    float T = 1.f, C[3] = {0.f, 0.f, 0.f};
    int contributor = 0, last = 0;
    for (int base = range.x; todo > 0; base += UP_THREADS, todo -= UP_THREADS) {
        if (__syncthreads_count(done) == UP_THREADS) break;
        if ((int)threadIdx.x < todo) {
            const uint32_t id = point_list[base + threadIdx.x];
            const float4 r0 = rec[3 * (size_t)id], r1 = rec[3 * (size_t)id + 1], r2 = rec[3 * (size_t)id + 2];
            s_xy[threadIdx.x] = make_float2(r0.x, r0.y);
            s_co[threadIdx.x] = make_float4(r0.z, r0.w, r1.x, r1.y);
            s_rgb[threadIdx.x][0] = r1.z; s_rgb[threadIdx.x][1] = r1.w; s_rgb[threadIdx.x][2] = r2.x;
        }
        __syncthreads();
        for (int j = 0; !done && j < min(UP_THREADS, todo); ++j) {
            ++contributor;
            const float dx = s_xy[j].x - fx, dy = s_xy[j].y - fy;
            const float4 co = s_co[j];
            const float power = -0.5f * (co.x * dx * dx + co.z * dy * dy) - co.y * dx * dy;
            if (power > 0.f) continue;
            const float alpha = fminf(0.99f, co.w * __expf(power));
            if (alpha < 1.f / 255.f) continue;
            const float test_T = T * (1.f - alpha);
            if (test_T < 0.0001f) { done = true; continue; }
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) C[ch] += s_rgb[j][ch] * alpha * T;
            T = test_T;
            last = contributor;
        }
    }
    if (inside) {
        const int pid = py * W + px;
        final_T[pid] = T;
        n_contrib[pid] = last;
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) out_color[(size_t)ch * H * W + pid] = C[ch] + T * bg[ch];
    }
}

__global__ void __launch_bounds__(UP_THREADS)
blend_bwd_upstream_kernel(const int2 *__restrict__ ranges, const uint32_t *__restrict__ point_list, const float4 *__restrict__ rec,
                          int W, int H, int gx, const float *__restrict__ bg, const float *__restrict__ final_T,
                          const int32_t *__restrict__ n_contrib, const float *__restrict__ dL_dpix, float *__restrict__ dL_dmean2D,
                          float *__restrict__ dL_dconic, float *__restrict__ dL_dopacity, float *__restrict__ dL_dcolor) {
    __shared__ uint32_t s_id[UP_THREADS];
    __shared__ float2 s_xy[UP_THREADS];
    __shared__ float4 s_co[UP_THREADS];
    __shared__ float s_rgb[UP_THREADS][3];
    const int tile = blockIdx.y * gx + blockIdx.x;
    const int px = blockIdx.x * TILE + (threadIdx.x & 15), py = blockIdx.y * TILE + (threadIdx.x >> 4);
    const bool inside = px < W && py < H;
    const int pid = py * W + px;
    const float fx = (float)px, fy = (float)py;
    const int2 range = ranges[tile];
    int todo = range.y - range.x;
    const int rounds = (todo + UP_THREADS - 1) / UP_THREADS;
    bool done = !inside;
    const float T_final = inside ? final_T[pid] : 0.f;
    float T = T_final;
    int contributor = todo;
    const int last = inside ? n_contrib[pid] : 0;
    float accum[3] = {0.f, 0.f, 0.f}, dpix[3] = {0.f, 0.f, 0.f}, last_color[3] = {0.f, 0.f, 0.f};
    float last_alpha = 0.f;
    if (inside)
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) dpix[ch] = dL_dpix[(size_t)ch * H * W + pid];
    const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;
    for (int i = 0; i < rounds; ++i, todo -= UP_THREADS) {
        __syncthreads();
        const int progress = i * UP_THREADS + threadIdx.x;
        if (range.x + progress < range.y) {                       // back to front
            const uint32_t id = point_list[range.y - progress - 1];
            const float4 r0 = rec[3 * (size_t)id], r1 = rec[3 * (size_t)id + 1], r2 = rec[3 * (size_t)id + 2];
            s_id[threadIdx.x] = id;
            s_xy[threadIdx.x] = make_float2(r0.x, r0.y);
            s_co[threadIdx.x] = make_float4(r0.z, r0.w, r1.x, r1.y);
            s_rgb[threadIdx.x][0] = r1.z; s_rgb[threadIdx.x][1] = r1.w; s_rgb[threadIdx.x][2] = r2.x;
        }
        __syncthreads();
        for (int j = 0; !done && j < min(UP_THREADS, todo); ++j) {
            --contributor;
            if (contributor >= last) continue;
            const float dx = s_xy[j].x - fx, dy = s_xy[j].y - fy;
            const float4 co = s_co[j];
            const float power = -0.5f * (co.x * dx * dx + co.z * dy * dy) - co.y * dx * dy;
            if (power > 0.f) continue;
            const float G = __expf(power);
            const float alpha = fminf(0.99f, co.w * G);
            if (alpha < 1.f / 255.f) continue;
            T = T / (1.f - alpha);
            const float dchannel_dcolor = alpha * T;
            float dL_dalpha = 0.f;
            const uint32_t id = s_id[j];
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                const float c = s_rgb[j][ch];
                accum[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum[ch];
                last_color[ch] = c;
                dL_dalpha += (c - accum[ch]) * dpix[ch];
                atomicAdd(&dL_dcolor[3 * (size_t)id + ch], dchannel_dcolor * dpix[ch]);
            }
            dL_dalpha *= T;
            last_alpha = alpha;
            float bg_dot = 0.f;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) bg_dot += bg[ch] * dpix[ch];
            dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot;
            const float dL_dG = co.w * dL_dalpha;
            const float gdx = G * dx, gdy = G * dy;
            const float dG_ddelx = -gdx * co.x - gdy * co.y, dG_ddely = -gdy * co.z - gdx * co.y;
            atomicAdd(&dL_dmean2D[3 * (size_t)id], dL_dG * dG_ddelx * ddelx_dx);
            atomicAdd(&dL_dmean2D[3 * (size_t)id + 1], dL_dG * dG_ddely * ddely_dy);
            atomicAdd(&dL_dconic[3 * (size_t)id], -0.5f * gdx * dx * dL_dG);
            atomicAdd(&dL_dconic[3 * (size_t)id + 1], -0.5f * gdx * dy * dL_dG);
            atomicAdd(&dL_dconic[3 * (size_t)id + 2], -0.5f * gdy * dy * dL_dG);
            atomicAdd(&dL_dopacity[id], G * dL_dalpha);
        }
    }
}

// Work census of a view's blend (for the issue-slot roofline bench.py reports): per pixel, how many instances the
// front-to-back walk visits before it terminates, and how many of those are LIVE (power <= 0 and alpha >= 1/255, i.e.
// they contribute colour and receive gradients).  out[0] = visited (pixel, instance) pairs, out[1] = live pairs.
__global__ void __launch_bounds__(UP_THREADS)
blend_census_kernel(const int2 *__restrict__ ranges, const uint32_t *__restrict__ point_list, const float4 *__restrict__ rec,
                    int W, int H, int gx, unsigned long long *__restrict__ out) {
    __shared__ float2 s_xy[UP_THREADS];
    __shared__ float4 s_co[UP_THREADS];
    const int tile = blockIdx.y * gx + blockIdx.x;
    const int px = blockIdx.x * TILE + (threadIdx.x & 15), py = blockIdx.y * TILE + (threadIdx.x >> 4);
    const bool inside = px < W && py < H;
    const float fx = (float)px, fy = (float)py;
    const int2 range = ranges[tile];
    int todo = range.y - range.x;
    bool done = !inside;
    float T = 1.f;
    unsigned visited = 0, live = 0;
    for (int base = range.x; todo > 0; base += UP_THREADS, todo -= UP_THREADS) {
        if (__syncthreads_count(done) == UP_THREADS) break;
        if ((int)threadIdx.x < todo) {
            const uint32_t id = point_list[base + threadIdx.x];
            const float4 r0 = rec[3 * (size_t)id], r1 = rec[3 * (size_t)id + 1];
            s_xy[threadIdx.x] = make_float2(r0.x, r0.y);
            s_co[threadIdx.x] = make_float4(r0.z, r0.w, r1.x, r1.y);
        }
        __syncthreads();
        for (int j = 0; !done && j < min(UP_THREADS, todo); ++j) {
            ++visited;
            const float dx = s_xy[j].x - fx, dy = s_xy[j].y - fy;
            const float4 co = s_co[j];
            const float power = -0.5f * (co.x * dx * dx + co.z * dy * dy) - co.y * dx * dy;
            if (power > 0.f) continue;
            const float alpha = fminf(0.99f, co.w * __expf(power));
            if (alpha < 1.f / 255.f) continue;
            const float test_T = T * (1.f - alpha);
            if (test_T < 0.0001f) { done = true; continue; }
            T = test_T;
            ++live;
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        visited += __shfl_xor_sync(0xffffffffu, visited, d);
        live += __shfl_xor_sync(0xffffffffu, live, d);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&out[0], (unsigned long long)visited);
        atomicAdd(&out[1], (unsigned long long)live);
    }
}

}  // namespace splatco

using namespace splatco;

extern "C" int splatco_blend_census(int64_t R, int H, int W, const void *geom, const void *binning, const void *image,
                                    unsigned long long *out2, void *stream) {
    SPLATCO_REQUIRE(H > 0 && W > 0 && R >= 0 && R < 0x7fffffff && image && out2, "blend_census: bad arguments");
    SPLATCO_CHECK_CUDA(cudaMemsetAsync(out2, 0, 2 * sizeof(unsigned long long), (cudaStream_t)stream));
    if (R == 0) return 0;
    SPLATCO_REQUIRE(geom && binning, "blend_census: null workspace");
    ImgWs im = img_view(const_cast<void *>(image), H, W);
    const int gx = ceil_div(W, TILE), gy = ceil_div(H, TILE);
    const uint32_t *plist = bin_view(const_cast<void *>(binning), R).vals[splatco_sorted_buffer_index(H, W)];
    blend_census_kernel<<<dim3(gx, gy), UP_THREADS, 0, (cudaStream_t)stream>>>(im.ranges, plist, reinterpret_cast<const float4 *>(geom),
                                                                              W, H, gx, out2);
    SPLATCO_CHECK_LAUNCH();
    return 0;
}

extern "C" int splatco_blend_fwd_upstream(int64_t R, int H, int W, const float *bg, const void *geom, const void *binning,
                                          void *image, float *out_color, void *stream) {
    SPLATCO_REQUIRE(H > 0 && W > 0 && R >= 0 && R < 0x7fffffff, "blend_fwd_upstream: bad sizes");
    SPLATCO_REQUIRE(bg && image && out_color && (R == 0 || (geom && binning)), "blend_fwd_upstream: null pointer");
    ImgWs im = img_view(image, H, W);
    const int gx = ceil_div(W, TILE), gy = ceil_div(H, TILE);
    const uint32_t *plist = R > 0 ? bin_view(const_cast<void *>(binning), R).vals[splatco_sorted_buffer_index(H, W)] : nullptr;
    blend_fwd_upstream_kernel<<<dim3(gx, gy), UP_THREADS, 0, (cudaStream_t)stream>>>(
        im.ranges, plist, reinterpret_cast<const float4 *>(geom), W, H, gx, bg, out_color, im.final_T, im.n_contrib);
    SPLATCO_CHECK_LAUNCH();
    return 0;
}

extern "C" int splatco_blend_bwd_upstream(int P, int64_t R, int H, int W, const float *bg, const void *geom, const void *binning,
                                          const void *image, const float *dL_dpix, float *dL_dmean2D, float *dL_dconic,
                                          float *dL_dopacity, float *dL_dcolor, void *stream) {
    SPLATCO_REQUIRE(H > 0 && W > 0 && R >= 0 && R < 0x7fffffff && P >= 0, "blend_bwd_upstream: bad sizes");
    if (P == 0 || R == 0) return 0;
    SPLATCO_REQUIRE(bg && geom && binning && image && dL_dpix && dL_dmean2D && dL_dconic && dL_dopacity && dL_dcolor,
                    "blend_bwd_upstream: null pointer");
    ImgWs im = img_view(const_cast<void *>(image), H, W);
    const int gx = ceil_div(W, TILE), gy = ceil_div(H, TILE);
    const uint32_t *plist = bin_view(const_cast<void *>(binning), R).vals[splatco_sorted_buffer_index(H, W)];
    blend_bwd_upstream_kernel<<<dim3(gx, gy), UP_THREADS, 0, (cudaStream_t)stream>>>(
        im.ranges, plist, reinterpret_cast<const float4 *>(geom), W, H, gx, bg, im.final_T, im.n_contrib, dL_dpix, dL_dmean2D,
        dL_dconic, dL_dopacity, dL_dcolor);
    SPLATCO_CHECK_LAUNCH();
    return 0;
}
