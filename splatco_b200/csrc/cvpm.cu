// CVPM pruning mask.
// Reference being replaced: the point-cloud half of GaussianModel.compute_fast_loss_with_key_points,
// scene/gaussian_model.py:1178-1216 (called per view pair at train.py:218-234 with existing_point_cloud = get_anchor,
// distance_threshold = voxel_size; the mask goes to prune_anchor).  For cameras with translation vectors t1, t2:
//   d      = (t2 - t1) / |t2 - t1|                                   ray through the two "centres"   :1179-1182
//   dist_c = | p - (t_c + d_c * ((p - t_c) . d_c)) |,  c = 1, 2      distance to that line           :1188-1195
//   valid  = dist_1 < distance_threshold  &  dist_2 < distance_threshold                              :1197
//   close  = |p - t1| < min_cam_distance  |  |p - t2| < min_cam_distance                              :1200-1202
//   outl   = ~all(|p - mean(P)| < sigma_threshold * std(P))          per-axis, unbiased std           :1205-1207
//   mask   = valid & (close | outl), and all-false when ssim(real_1, real_2) < overall_ssim_threshold :1163-1165,1210
// The reference spends ~25 elementwise / reduction launches and a host sync on the SSIM gate; here: one reduction
// (fp64 sums about a pivot), one 1-thread finalize, one mask kernel that also counts.  44 B/anchor -> 13 B/anchor read+write.
#include "common.cuh"

namespace splatco {

// ws (doubles): [0..2] sum(x - pivot), [3..5] sum((x - pivot)^2); floats at ws+8 doubles: mean[3], 3*std[3]
__global__ void __launch_bounds__(256)
cvpm_stats_kernel(int N, const float *__restrict__ pts, double *__restrict__ ws) {
    const float px = __ldg(pts), py = __ldg(pts + 1), pz = __ldg(pts + 2);
    double s[6] = {0, 0, 0, 0, 0, 0};
    for (int i = blockIdx.x * 256 + threadIdx.x; i < N; i += gridDim.x * 256) {
        const double x = (double)(pts[3 * (size_t)i] - px), y = (double)(pts[3 * (size_t)i + 1] - py), z = (double)(pts[3 * (size_t)i + 2] - pz);
        s[0] += x; s[1] += y; s[2] += z; s[3] += x * x; s[4] += y * y; s[5] += z * z;
    }
    __shared__ double red[8][6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        double a = s[k];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) a += __shfl_xor_sync(0xffffffffu, a, d);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][k] = a;
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        double a = 0;
        for (int w = 0; w < 8; ++w) a += red[w][threadIdx.x];
        atomicAdd(ws + threadIdx.x, a);
    }
}

__global__ void cvpm_finalize_kernel(int N, const float *__restrict__ pts, double *__restrict__ ws, float sigma) {
    float *f = (float *)(ws + 8);
    for (int k = 0; k < 3; ++k) {
        const double n = (double)N, m = ws[k] / n;
        const double var = (ws[3 + k] - n * m * m) / (n - 1.0);          // unbiased (torch.std default); N = 1 -> NaN like torch
        f[k] = (float)((double)pts[k] + m);
        f[3 + k] = sigma * (float)sqrt(var > 0.0 ? var : (var == var ? 0.0 : var));
    }
}

__global__ void __launch_bounds__(256)
cvpm_mask_kernel(int N, const float *__restrict__ pts, const float *__restrict__ t1, const float *__restrict__ t2,
                 const float *__restrict__ ssim, float ssim_thr, float dist_thr, float cam_thr, const double *__restrict__ ws,
                 uint8_t *__restrict__ mask, int32_t *__restrict__ count) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    bool m = false;
    if (i < N && !(ssim && __ldg(ssim) < ssim_thr)) {
        const float *f = (const float *)(ws + 8);
        float c1[3], c2[3], d1[3], d2[3], p[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) { c1[k] = __ldg(t1 + k); c2[k] = __ldg(t2 + k); d1[k] = __fsub_rn(c2[k], c1[k]); d2[k] = __fsub_rn(c1[k], c2[k]); p[k] = pts[3 * (size_t)i + k]; }
        const float n1 = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(d1[0], d1[0]), __fmul_rn(d1[1], d1[1])), __fmul_rn(d1[2], d1[2])));
#pragma unroll
        for (int k = 0; k < 3; ++k) { d1[k] = __fdiv_rn(d1[k], n1); d2[k] = __fdiv_rn(d2[k], n1); }
        float dist[2], cam[2];
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const float *cc = c ? c2 : c1, *dd = c ? d2 : d1;
            float r[3], dot = 0.f, cs = 0.f, ds = 0.f;
#pragma unroll
            for (int k = 0; k < 3; ++k) { r[k] = __fsub_rn(p[k], cc[k]); dot = fmaf(r[k], dd[k], dot); cs = __fadd_rn(cs, __fmul_rn(r[k], r[k])); }
#pragma unroll
            for (int k = 0; k < 3; ++k) { const float q = __fsub_rn(p[k], __fadd_rn(cc[k], __fmul_rn(dd[k], dot))); ds = __fadd_rn(ds, __fmul_rn(q, q)); }
            dist[c] = __fsqrt_rn(ds); cam[c] = __fsqrt_rn(cs);
        }
        const bool valid = dist[0] < dist_thr && dist[1] < dist_thr;
        const bool close = cam[0] < cam_thr || cam[1] < cam_thr;
        bool inl = true;
#pragma unroll
        for (int k = 0; k < 3; ++k) inl = inl && (fabsf(__fsub_rn(p[k], f[k])) < f[3 + k]);
        m = valid && (close || !inl);
    }
    if (i < N) mask[i] = m ? 1 : 0;
    const uint32_t b = __ballot_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(count, __popc(b));
}

}  // namespace splatco

using namespace splatco;

extern "C" size_t splatco_cvpm_ws_bytes(void) { return align_up(8 * sizeof(double) + 8 * sizeof(float)); }

extern "C" int splatco_cvpm_mask(int N, const float *points, const float *t1, const float *t2, const float *ssim,
                                 float ssim_threshold, float distance_threshold, float sigma_threshold, float min_cam_distance,
                                 void *ws, uint8_t *mask, int32_t *count, void *stream) {
    SPLATCO_REQUIRE(N >= 0, "cvpm_mask: bad N=%d", N);
    SPLATCO_REQUIRE(count, "cvpm_mask: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    SPLATCO_CHECK_CUDA(cudaMemsetAsync(count, 0, sizeof(int32_t), st));
    if (N == 0) return 0;
    SPLATCO_REQUIRE(points && t1 && t2 && ws && mask, "cvpm_mask: null pointer");
    double *w = (double *)ws;
    SPLATCO_CHECK_CUDA(cudaMemsetAsync(w, 0, splatco_cvpm_ws_bytes(), st));
    const int grid = ceil_div(N, 256) < 148 * 4 ? ceil_div(N, 256) : 148 * 4;
    cvpm_stats_kernel<<<grid, 256, 0, st>>>(N, points, w);
    SPLATCO_CHECK_LAUNCH();
    cvpm_finalize_kernel<<<1, 1, 0, st>>>(N, points, w, sigma_threshold);
    SPLATCO_CHECK_LAUNCH();
    cvpm_mask_kernel<<<ceil_div(N, 256), 256, 0, st>>>(N, points, t1, t2, ssim, ssim_threshold, distance_threshold, min_cam_distance, w, mask, count);
    SPLATCO_CHECK_LAUNCH();
    return 0;
}
