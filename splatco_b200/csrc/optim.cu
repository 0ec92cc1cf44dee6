// Fused multi-tensor Adam step ("fused Adam over planes").
// Reference being replaced: `gaussians.optimizer.step()` (train.py:310-312) with the optimizer built at
// scene/gaussian_model.py:519-572 — torch.optim.Adam(groups, lr=0.0, eps=1e-15), betas (0.9, 0.999), no weight decay,
// no amsgrad, per-group learning rates.  torch's default (foreach) implementation makes ~10 passes over every
// parameter, gradient and moment (lerp, mul, addcmul, sqrt, div, add, addcdiv, each a multi-tensor launch chain);
// the planes alone are ~0.7 GB of optimizer traffic per step at plane_size 2800.  Here every tensor is updated in one
// pass — read p, g, m, v, write p, m, v = 28 B per element, HBM-bound — and up to 64 tensors share one launch.
// Per element (torch/optim/adam.py, _single_tensor_adam):
//     m += (g - m) * (1 - beta1);   v = v * beta2 + (1 - beta2) * g * g
//     p -= (lr / (1 - beta1^t)) * m / (sqrt(v) / sqrt(1 - beta2^t) + eps)
#include <math.h>

#include "common.cuh"

namespace splatco {

constexpr int ADAM_MAX = 64;                 // tensors per launch (kernel-parameter space: 64 * 48 B + 260 B)
constexpr int ADAM_CHUNK = 256 * 4 * 4;      // elements per CTA: 256 threads x 4 iterations x float4

struct AdamBatch {
    float *p[ADAM_MAX];
    const float *g[ADAM_MAX];
    float *m[ADAM_MAX];
    float *v[ADAM_MAX];
    int64_t n[ADAM_MAX];
    float step_size[ADAM_MAX], bc2_sqrt[ADAM_MAX];
    int first_chunk[ADAM_MAX + 1];           // CTA b works on tensor t with first_chunk[t] <= b < first_chunk[t+1]
    int count;
};

__device__ __forceinline__ void adam_one(float &p, float g, float &m, float &v, float w1, float b2, float w2, float eps,
                                         float step_size, float bc2_sqrt) {
    m = m + (g - m) * w1;
    v = v * b2 + w2 * g * g;
    p = p - step_size * (m / (sqrtf(v) / bc2_sqrt + eps));
}

__global__ void __launch_bounds__(256)
adam_kernel(const __grid_constant__ AdamBatch b, float w1, float b2, float w2, float eps) {
    int t = 0;
    {   // which tensor: binary search over the chunk prefix (<= 6 steps)
        int lo = 0, hi = b.count;
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (b.first_chunk[mid] <= (int)blockIdx.x) lo = mid; else hi = mid; }
        t = lo;
    }
    const int64_t n = b.n[t];
    const int64_t base = (int64_t)(blockIdx.x - b.first_chunk[t]) * ADAM_CHUNK;
    float *__restrict__ p = b.p[t];
    const float *__restrict__ g = b.g[t];
    float *__restrict__ m = b.m[t];
    float *__restrict__ v = b.v[t];
    const float ss = b.step_size[t], bs = b.bc2_sqrt[t];
    const bool vec = ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0);
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int64_t i = base + ((int64_t)it * 256 + threadIdx.x) * 4;
        if (i >= n) break;
        if (vec && i + 4 <= n) {
            float4 P = *(float4 *)(p + i), M = *(float4 *)(m + i), V = *(float4 *)(v + i);
            const float4 G = __ldg((const float4 *)(g + i));
            adam_one(P.x, G.x, M.x, V.x, w1, b2, w2, eps, ss, bs); adam_one(P.y, G.y, M.y, V.y, w1, b2, w2, eps, ss, bs);
            adam_one(P.z, G.z, M.z, V.z, w1, b2, w2, eps, ss, bs); adam_one(P.w, G.w, M.w, V.w, w1, b2, w2, eps, ss, bs);
            *(float4 *)(p + i) = P; *(float4 *)(m + i) = M; *(float4 *)(v + i) = V;
        } else {
            for (int64_t j = i; j < n && j < i + 4; ++j) {
                float P = p[j], M = m[j], V = v[j];
                adam_one(P, g[j], M, V, w1, b2, w2, eps, ss, bs);
                p[j] = P; m[j] = M; v[j] = V;
            }
        }
    }
}

}  // namespace splatco

using namespace splatco;

extern "C" int splatco_adam_step(int n_tensors, const splatco_adam_tensor *tensors, double beta1, double beta2, double eps,
                                 void *stream) {
    SPLATCO_REQUIRE(n_tensors >= 0 && (n_tensors == 0 || tensors), "adam_step: bad tensor list");
    cudaStream_t st = (cudaStream_t)stream;
    int done = 0;
    while (done < n_tensors) {
        AdamBatch b;
        memset(&b, 0, sizeof(b));
        int chunks = 0, k = 0;
        for (; done < n_tensors && k < ADAM_MAX; ++done) {
            const splatco_adam_tensor &t = tensors[done];
            SPLATCO_REQUIRE(t.numel >= 0 && t.step >= 1, "adam_step: tensor %d has numel %lld, step %lld", done, (long long)t.numel, (long long)t.step);
            if (t.numel == 0) continue;
            SPLATCO_REQUIRE(t.param && t.grad && t.exp_avg && t.exp_avg_sq, "adam_step: tensor %d has a null pointer", done);
            const int64_t c = (t.numel + ADAM_CHUNK - 1) / ADAM_CHUNK;
            SPLATCO_REQUIRE(c < (1 << 30) - chunks, "adam_step: tensor %d too large", done);
            b.p[k] = t.param; b.g[k] = t.grad; b.m[k] = t.exp_avg; b.v[k] = t.exp_avg_sq; b.n[k] = t.numel;
            const double bc1 = 1.0 - pow(beta1, (double)t.step), bc2 = 1.0 - pow(beta2, (double)t.step);
            b.step_size[k] = (float)((double)t.lr / bc1);
            b.bc2_sqrt[k] = (float)sqrt(bc2);
            b.first_chunk[k] = chunks;
            chunks += (int)c;
            ++k;
        }
        b.first_chunk[k] = chunks;
        b.count = k;
        if (k == 0) continue;
        // 1 - beta in double, then rounded: torch passes `1 - beta2` as a double scalar (1 - 0.999f would be off by 1.3e-5)
        adam_kernel<<<chunks, 256, 0, st>>>(b, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), (float)eps);
        SPLATCO_CHECK_LAUNCH();
    }
    return 0;
}
