// Densification statistics of one view, fused.
// Reference being replaced: GaussianModel.training_statis, scene/gaussian_model.py:761-782, called with the LAST
// view's render outputs at train.py:264-266.  The reference runs ~15 masked / indexed torch ops over [N*K] tensors
// (two boolean scatters build `combined_mask`); here one thread per visible anchor walks its K offsets:
//   opacity_accum[i] += sum_k max(neural_opacity[v,k], 0)      anchor_demon[i] += 1
//   for every offset k that survived the opacity mask (compacted index j) and is on screen (update_filter[j]):
//       offset_gradient_accum[i*K+k] += |viewspace_grad[j, :2]|    offset_denom[i*K+k] += 1
// Every accumulator element has a single owner thread, so plain read-modify-writes are exact and the integer-valued
// counters (anchor_demon, offset_denom) match the reference bit for bit.
#include "common.cuh"

namespace splatco {

__global__ void __launch_bounds__(256)
training_statis_kernel(int V, int K, const int32_t *__restrict__ vis_idx, const float *__restrict__ neural_opacity,
                       const uint8_t *__restrict__ sel, const int32_t *__restrict__ sel_excl,
                       const uint8_t *__restrict__ update_filter, const float *__restrict__ grad /*[M,3]*/,
                       float *__restrict__ opacity_accum, float *__restrict__ anchor_demon,
                       float *__restrict__ offset_gradient_accum, float *__restrict__ offset_denom) {
    const int v = blockIdx.x * 256 + threadIdx.x;
    if (v >= V) return;
    const size_t i = (size_t)vis_idx[v];
    float s = 0.f;
    for (int k = 0; k < K; ++k) {
        const size_t t = (size_t)v * K + k;
        s += fmaxf(neural_opacity[t], 0.f);
        if (sel[t]) {
            const int j = sel_excl[t];
            if (update_filter[j]) {
                const float gx = grad[3 * (size_t)j], gy = grad[3 * (size_t)j + 1];
                offset_gradient_accum[i * K + k] += sqrtf(gx * gx + gy * gy);
                offset_denom[i * K + k] += 1.0f;
            }
        }
    }
    opacity_accum[i] += s;
    anchor_demon[i] += 1.0f;
}

}  // namespace splatco

using namespace splatco;

extern "C" int splatco_training_statis(int V, int K, const int32_t *vis_idx, const float *neural_opacity,
                                       const uint8_t *selection_mask, const int32_t *selection_excl,
                                       const uint8_t *update_filter, const float *viewspace_grad, float *opacity_accum,
                                       float *anchor_demon, float *offset_gradient_accum, float *offset_denom,
                                       void *stream) {
    SPLATCO_REQUIRE(V >= 0 && K >= 1, "training_statis: bad sizes V=%d K=%d", V, K);
    if (V == 0) return 0;
    SPLATCO_REQUIRE(vis_idx && neural_opacity && selection_mask && selection_excl && opacity_accum && anchor_demon &&
                    offset_gradient_accum && offset_denom, "training_statis: null pointer");
    training_statis_kernel<<<ceil_div(V, 256), 256, 0, (cudaStream_t)stream>>>(
        V, K, vis_idx, neural_opacity, selection_mask, selection_excl, update_filter, viewspace_grad, opacity_accum,
        anchor_demon, offset_gradient_accum, offset_denom);
    SPLATCO_CHECK_LAUNCH();
    return 0;
}
