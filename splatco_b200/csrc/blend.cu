// Per-tile front-to-back alpha blend, forward and backward.  Replaces upstream renderCUDA<3> fwd/bwd
// [SURVEY.md Appendix A.4, A.5; reference call site gaussian_renderer/__init__.py:163-171 and the
// autograd backward triggered at train.py:240].
//
// One CTA per 16x16 tile, one thread per pixel; warps cover 8x4-pixel patches (not 16x2 rows) so
// that warp-wide skips (__any/__all ballots) fire more often.  Splats are staged through shared
// memory in batches of 256 (three 16-byte loads of the packed record per splat); a warp leaves the
// batch loop as soon as all its pixels are saturated and the CTA ends when every warp has.
// Backward: gradients of the 9 per-splat scalars are reduced across the warp with shuffles, then
// accumulated per CTA in shared memory, and flushed with ONE global atomic (RED) per scalar per
// (tile, splat) — 256x fewer global atomics than the per-pixel atomics of the upstream kernel.
#include "common.cuh"

namespace splatco {

constexpr int BLEND_THREADS = TILE * TILE;   // 256
constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ void pixel_of_thread(int tile_x, int tile_y, int &px, int &py) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    px = tile_x * TILE + (warp & 1) * 8 + (lane & 7);
    py = tile_y * TILE + (warp >> 1) * 4 + (lane >> 3);
}

__global__ void __launch_bounds__(BLEND_THREADS)
blend_fwd_kernel(const int2 *__restrict__ ranges, const uint32_t *__restrict__ point_list,
                 const float4 *__restrict__ rec, int W, int H, int gx, const float *__restrict__ bg,
                 float *__restrict__ out_color, float *__restrict__ final_T,
                 int32_t *__restrict__ n_contrib) {
    __shared__ float4 s_a[BLEND_THREADS];   // x, y, conA, conB
    __shared__ float4 s_b[BLEND_THREADS];   // conC, opacity, r, g
    __shared__ float s_c[BLEND_THREADS];    // b
    const int tile = blockIdx.x;
    const int tile_x = tile % gx, tile_y = tile / gx;
    int px, py;
    pixel_of_thread(tile_x, tile_y, px, py);
    const bool inside = px < W && py < H;
    const float pxf = (float)px, pyf = (float)py;
    const int2 range = ranges[tile];
    int todo = range.y - range.x;
    bool done = !inside;
    float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
    int contributor = 0, last_contributor = 0;

    for (int base = range.x; todo > 0; base += BLEND_THREADS, todo -= BLEND_THREADS) {
        if (__syncthreads_and(done)) break;
        if ((int)threadIdx.x < todo) {
            const uint32_t id = point_list[base + threadIdx.x];
            const float4 r0 = rec[3 * (size_t)id], r1 = rec[3 * (size_t)id + 1], r2 = rec[3 * (size_t)id + 2];
            s_a[threadIdx.x] = r0; s_b[threadIdx.x] = r1; s_c[threadIdx.x] = r2.x;
        }
        __syncthreads();
        const int nb = min(BLEND_THREADS, todo);
        {
            for (int j = 0; j < nb; ++j) {
                // warp-uniform early exit (j is uniform, every lane reaches the vote)
                if ((j & 3) == 0 && __all_sync(0xffffffffu, done)) break;
                if (done) continue;
                contributor = (base - range.x) + j + 1;
                const float4 a = s_a[j];
                const float4 b = s_b[j];
                const float dx = a.x - pxf, dy = a.y - pyf;
                const float power = -0.5f * (a.z * dx * dx + b.x * dy * dy) - a.w * dx * dy;
                if (power > 0.0f) continue;
                const float alpha = fminf(0.99f, b.y * ex2_approx(power * kLog2e));
                if (alpha < 1.0f / 255.0f) continue;
                const float test_T = T * (1.0f - alpha);
                if (test_T < 0.0001f) { done = true; continue; }
                const float w = alpha * T;
                C0 = fmaf(b.z, w, C0); C1 = fmaf(b.w, w, C1); C2 = fmaf(s_c[j], w, C2);
                T = test_T;
                last_contributor = contributor;
            }
        }
    }
    if (inside) {
        const size_t pid = (size_t)py * W + px, HW = (size_t)H * W;
        final_T[pid] = T;
        n_contrib[pid] = last_contributor;
        out_color[pid] = C0 + T * bg[0];
        out_color[HW + pid] = C1 + T * bg[1];
        out_color[2 * HW + pid] = C2 + T * bg[2];
    }
}

// ---- backward ---------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}

__global__ void __launch_bounds__(BLEND_THREADS)
blend_bwd_kernel(const int2 *__restrict__ ranges, const uint32_t *__restrict__ point_list,
                 const float4 *__restrict__ rec, int W, int H, int gx, const float *__restrict__ bg,
                 const float *__restrict__ final_T, const int32_t *__restrict__ n_contrib,
                 const float *__restrict__ dL_dpix, float *__restrict__ dL_dmean2D,
                 float *__restrict__ dL_dconic, float *__restrict__ dL_dopacity,
                 float *__restrict__ dL_dcolor) {
    __shared__ float4 s_a[BLEND_THREADS];
    __shared__ float4 s_b[BLEND_THREADS];
    __shared__ float s_c[BLEND_THREADS];
    __shared__ uint32_t s_id[BLEND_THREADS];
    __shared__ float s_acc[9][BLEND_THREADS];      // per-CTA accumulation of the batch's gradients
    __shared__ int s_max[BLEND_THREADS / 32];
    const int tile = blockIdx.x;
    const int tile_x = tile % gx, tile_y = tile / gx;
    int px, py;
    pixel_of_thread(tile_x, tile_y, px, py);
    const bool inside = px < W && py < H;
    const float pxf = (float)px, pyf = (float)py;
    const int2 range = ranges[tile];
    const size_t pid = (size_t)py * W + px, HW = (size_t)H * W;
    const float T_final = inside ? final_T[pid] : 0.f;
    const int last = inside ? n_contrib[pid] : 0;
    float T = T_final;
    float dp0 = 0.f, dp1 = 0.f, dp2 = 0.f;
    if (inside) { dp0 = dL_dpix[pid]; dp1 = dL_dpix[HW + pid]; dp2 = dL_dpix[2 * HW + pid]; }
    const float bg_dot = bg[0] * dp0 + bg[1] * dp1 + bg[2] * dp2;
    float ar0 = 0.f, ar1 = 0.f, ar2 = 0.f, lc0 = 0.f, lc1 = 0.f, lc2 = 0.f, last_alpha = 0.f;
    const float ddelx_dx = 0.5f * (float)W, ddely_dy = 0.5f * (float)H;

    // deepest contributor over the tile: nothing behind it receives gradient
    int m = last;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, d));
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = m;
    __syncthreads();
    int tile_last = 0;
#pragma unroll
    for (int w = 0; w < BLEND_THREADS / 32; ++w) tile_last = max(tile_last, s_max[w]);
    const int lane = threadIdx.x & 31;

    // walk positions tile_last-1 .. 0 in batches of 256, back to front
    for (int hi = tile_last; hi > 0; hi -= BLEND_THREADS) {
        const int nb = min(BLEND_THREADS, hi);
        __syncthreads();                        // previous batch fully flushed
        {
            // slot j holds position hi-1-j
            const int j = threadIdx.x;
            if (j < nb) {
                const uint32_t id = point_list[range.x + hi - 1 - j];
                const float4 r0 = rec[3 * (size_t)id], r1 = rec[3 * (size_t)id + 1], r2 = rec[3 * (size_t)id + 2];
                s_a[j] = r0; s_b[j] = r1; s_c[j] = r2.x; s_id[j] = id;
            }
#pragma unroll
            for (int q = 0; q < 9; ++q) s_acc[q][j] = 0.f;
        }
        __syncthreads();
        // warp-uniform bound: positions >= warp's deepest contributor cannot contribute
        const int warp_last = __reduce_max_sync(0xffffffffu, last);
        for (int j = 0; j < nb; ++j) {
            const int pos = hi - 1 - j;
            if (pos >= warp_last) continue;
            float g[9];
            bool live = pos < last;
            float alpha = 0.f, G = 0.f, dx = 0.f, dy = 0.f;
            const float4 a = s_a[j];
            const float4 b = s_b[j];
            if (live) {
                dx = a.x - pxf; dy = a.y - pyf;
                const float power = -0.5f * (a.z * dx * dx + b.x * dy * dy) - a.w * dx * dy;
                G = ex2_approx(power * kLog2e);
                alpha = fminf(0.99f, b.y * G);
                live = (power <= 0.0f) && (alpha >= 1.0f / 255.0f);
            }
            if (!__any_sync(0xffffffffu, live)) continue;
            if (live) {
                T = T / (1.0f - alpha);
                const float dch = alpha * T;
                const float c0 = b.z, c1 = b.w, c2 = s_c[j];
                ar0 = last_alpha * lc0 + (1.f - last_alpha) * ar0; lc0 = c0;
                ar1 = last_alpha * lc1 + (1.f - last_alpha) * ar1; lc1 = c1;
                ar2 = last_alpha * lc2 + (1.f - last_alpha) * ar2; lc2 = c2;
                float dL_dalpha = ((c0 - ar0) * dp0 + (c1 - ar1) * dp1 + (c2 - ar2) * dp2) * T;
                last_alpha = alpha;
                dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot;
                const float dL_dG = b.y * dL_dalpha;
                const float gdx = G * dx, gdy = G * dy;
                g[0] = dL_dG * (-gdx * a.z - gdy * a.w) * ddelx_dx;
                g[1] = dL_dG * (-gdy * b.x - gdx * a.w) * ddely_dy;
                g[2] = -0.5f * gdx * dx * dL_dG;
                g[3] = -0.5f * gdx * dy * dL_dG;
                g[4] = -0.5f * gdy * dy * dL_dG;
                g[5] = G * dL_dalpha;
                g[6] = dch * dp0; g[7] = dch * dp1; g[8] = dch * dp2;
            } else {
#pragma unroll
                for (int q = 0; q < 9; ++q) g[q] = 0.f;
            }
#pragma unroll
            for (int q = 0; q < 9; ++q) g[q] = warp_sum(g[q]);
            if (lane < 9) {
                float v = g[0];
#pragma unroll
                for (int q = 1; q < 9; ++q) v = (lane == q) ? g[q] : v;
                atomicAdd(&s_acc[lane][j], v);
            }
        }
        __syncthreads();
        {
            const int j = threadIdx.x;
            if (j < nb) {
                const uint32_t id = s_id[j];
                float v[9];
                bool any = false;
#pragma unroll
                for (int q = 0; q < 9; ++q) { v[q] = s_acc[q][j]; any |= (v[q] != 0.f); }
                if (any) {
                    atomicAdd(&dL_dmean2D[3 * (size_t)id], v[0]);
                    atomicAdd(&dL_dmean2D[3 * (size_t)id + 1], v[1]);
                    atomicAdd(&dL_dconic[3 * (size_t)id], v[2]);
                    atomicAdd(&dL_dconic[3 * (size_t)id + 1], v[3]);
                    atomicAdd(&dL_dconic[3 * (size_t)id + 2], v[4]);
                    atomicAdd(&dL_dopacity[id], v[5]);
                    atomicAdd(&dL_dcolor[3 * (size_t)id], v[6]);
                    atomicAdd(&dL_dcolor[3 * (size_t)id + 1], v[7]);
                    atomicAdd(&dL_dcolor[3 * (size_t)id + 2], v[8]);
                }
            }
        }
    }
}

}  // namespace splatco

using namespace splatco;

extern "C" int splatco_blend_fwd(int64_t R, int H, int W, const float *bg, const void *geom,
                                 const void *binning, void *image, float *out_color, void *stream) {
    SPLATCO_REQUIRE(H > 0 && W > 0 && R >= 0 && R < 0x7fffffff, "blend_fwd: bad sizes");
    SPLATCO_REQUIRE(bg && image && out_color, "blend_fwd: null pointer");
    SPLATCO_REQUIRE(R == 0 || (geom && binning), "blend_fwd: null workspace with R>0");
    ImgWs im = img_view(image, H, W);
    const int gx = ceil_div(W, TILE), gy = ceil_div(H, TILE);
    const uint32_t *plist = nullptr;
    const float4 *rec = nullptr;
    if (R > 0) {
        BinWs b = bin_view(const_cast<void *>(binning), R);
        plist = b.vals[splatco_sorted_buffer_index(H, W)];
        rec = reinterpret_cast<const float4 *>(geom);      // chunk 0 of the geometry workspace
    }
    blend_fwd_kernel<<<gx * gy, BLEND_THREADS, 0, (cudaStream_t)stream>>>(im.ranges, plist, rec, W, H, gx, bg,
                                                                         out_color, im.final_T, im.n_contrib);
    SPLATCO_CHECK_LAUNCH();
    return 0;
}

extern "C" int splatco_blend_bwd(int P, int64_t R, int H, int W, const float *bg, const void *geom,
                                 const void *binning, const void *image, const float *dL_dpix,
                                 float *dL_dmean2D, float *dL_dconic, float *dL_dopacity,
                                 float *dL_dcolor, void *stream) {
    SPLATCO_REQUIRE(H > 0 && W > 0 && R >= 0 && R < 0x7fffffff && P >= 0, "blend_bwd: bad sizes");
    if (P == 0 || R == 0) return 0;
    SPLATCO_REQUIRE(bg && geom && binning && image && dL_dpix && dL_dmean2D && dL_dconic && dL_dopacity && dL_dcolor,
                    "blend_bwd: null pointer");
    ImgWs im = img_view(const_cast<void *>(image), H, W);
    BinWs b = bin_view(const_cast<void *>(binning), R);
    const int gx = ceil_div(W, TILE), gy = ceil_div(H, TILE);
    blend_bwd_kernel<<<gx * gy, BLEND_THREADS, 0, (cudaStream_t)stream>>>(
        im.ranges, b.vals[splatco_sorted_buffer_index(H, W)], reinterpret_cast<const float4 *>(geom), W, H, gx, bg,
        im.final_T, im.n_contrib, dL_dpix, dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor);
    SPLATCO_CHECK_LAUNCH();
    return 0;
}
