// Plane total-variation regulariser, gradient added in place.
// Reference being replaced: PlaneGrid.total_variation_add_grad, scene/grids.py:240-250, driven by
// GaussianLearner.tv_loss (scene/gaussian_model.py:217-220) every 4th iteration after total_loss.backward()
// (train.py:242-243).  Per plane the reference builds
//     loss = w/6 * ( smooth_l1(p[:,:,1:], p[:,:,:-1], 'sum') + smooth_l1(p[:,:,:,1:], p[:,:,:,:-1], 'sum') )
// (beta = 1) and calls backward(): 2 slices + 1 smooth-L1 forward + its backward + 2 padded slice-gradients + the
// AccumulateGrad add, for each of 6 terms — ~40 launches and ~20 passes over the planes.  The gradient is a 5-point
// stencil of clamped differences,
//     dL/dp[i,j] = w/6 * ( clamp(p[i,j]-p[i-1,j]) - clamp(p[i+1,j]-p[i,j]) + clamp(p[i,j]-p[i,j-1]) - clamp(p[i,j+1]-p[i,j]) ),
// clamp to [-1, 1]; one kernel adds it to the plane's gradient buffer: 4 B read (neighbours come from L1/L2) + 8 B
// read-modify-write per texel, HBM-bound.
#include "common.cuh"

namespace splatco {

__device__ __forceinline__ float tv_clamp(float d) { return fminf(fmaxf(d, -1.f), 1.f); }

// grid: (ceil(W/128), ceil(H/TV_ROWS), C); a CTA walks TV_ROWS rows of 128 columns, carrying the row above in a register
constexpr int TV_ROWS = 8;
__global__ void __launch_bounds__(128)
tv_add_grad_kernel(int H, int W, const float *__restrict__ p, float *__restrict__ g, float coef) {
    const int j = blockIdx.x * 128 + threadIdx.x;
    if (j >= W) return;
    const int i0 = blockIdx.y * TV_ROWS;
    const size_t base = (size_t)blockIdx.z * H * W;
    const float *pc = p + base;
    float up = i0 > 0 ? __ldg(pc + (size_t)(i0 - 1) * W + j) : 0.f;
    float cur = __ldg(pc + (size_t)i0 * W + j);
    const int i1 = min(i0 + TV_ROWS, H);
    for (int i = i0; i < i1; ++i) {
        const size_t o = (size_t)i * W + j;
        const float dn = i + 1 < H ? __ldg(pc + o + W) : 0.f;
        float s = 0.f;
        if (i > 0) s += tv_clamp(cur - up);
        if (i + 1 < H) s -= tv_clamp(dn - cur);
        if (j > 0) s += tv_clamp(cur - __ldg(pc + o - 1));
        if (j + 1 < W) s -= tv_clamp(__ldg(pc + o + 1) - cur);
        g[base + o] += coef * s;
        up = cur; cur = dn;
    }
}

}  // namespace splatco

using namespace splatco;

extern "C" int splatco_tv_add_grad(int C, int H, int W, const float *plane, float *grad, float w, void *stream) {
    SPLATCO_REQUIRE(C >= 1 && C <= 65535 && H >= 1 && W >= 1 && ceil_div(H, TV_ROWS) <= 65535, "tv_add_grad: bad sizes C=%d H=%d W=%d", C, H, W);
    SPLATCO_REQUIRE(plane && grad, "tv_add_grad: null pointer");
    const dim3 grid(ceil_div(W, 128), ceil_div(H, TV_ROWS), C);
    tv_add_grad_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(H, W, plane, grad, (1.0f / 6.0f) * w);   // d(loss/6)/dterm, then * w, as autograd forms it
    SPLATCO_CHECK_LAUNCH();
    return 0;
}
