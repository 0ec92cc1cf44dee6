// Anchor growing (densification) — SURVEY.md §8 row f4.
// Reference being replaced: GaussianModel.anchor_growing, scene/gaussian_model.py:832-925 (one pass of its
// `for i in range(self.update_depth)` loop per call sequence below), called from adjust_anchor (:929-947).
//
// What one pass of the reference computes (line numbers in scene/gaussian_model.py):
//   candidate_mask = (grads >= thr) & offset_mask & (rand > cut)                                   :839-846
//   all_xyz        = anchor[:,None] + offset * scaling[:, :3][:,None]                              :855
//   grid_coords    = round(anchor / cur_size).int()          (every existing anchor)               :862
//   sel_coords     = round(all_xyz[candidate_mask] / cur_size).int()                               :864-865
//   uniq, inverse  = unique(sel_coords, dim=0)                (rows sorted lexicographically)      :867
//   keep           = ~any(uniq == grid_coords)                (chunked O(U*N) compare)             :871-884
//   candidate_anchor = uniq[keep] * cur_size                                                       :885
//   new_feat       = scatter_max(anchor_feat of the candidates, inverse)[keep]                     :895-897
// The reference materialises an [N*K, 32] repeat of the features, a dim-0 unique over int triples (sort of all
// candidates) and an O(U*N/4096) chain of broadcast compares.  Here the voxel set is a hash table of candidate
// indices (keys are compared through the coordinate array, so they can be any 3x int32), every existing anchor probes
// it once, only the surviving voxels are sorted (x, y, z signed-lexicographic = torch.unique's row order) and the
// per-voxel feature maximum is a float atomic max straight into the output rows.  All integer work: the new anchors
// and their features equal the reference's bit for bit (max is order-independent).
//
// Division by cur_size: torch's CUDA kernel for `tensor / python_scalar` multiplies by the fp32 reciprocal
// (ATen BinaryDivTrueKernel.cu, "compute a * reciprocal(b)"), its CPU kernel divides; `div_mode` selects which
// (0 = reciprocal, what the reference does on the GPU; 1 = true division, what the CPU-generated fixtures hold).
#include "common.cuh"

namespace splatco {

constexpr uint32_t GROW_EMPTY = 0xffffffffu;

struct GrowWs {
    uint32_t *slot;      // [n]  offset slot (anchor * K + k) of candidate c
    int32_t *xyz;        // [3n] grid coordinates of candidate c
    uint32_t *rep;       // [n]  representative candidate of c's voxel
    uint32_t *row;       // [n]  output row of a surviving representative
    uint8_t *removed;    // [n]  representative's voxel already holds an anchor
    uint32_t *table;     // [tsize]
    void *bin;           // binning workspace of capacity n (keys / values ping-pong + histograms)
};

static inline uint32_t grow_table_size(int64_t n) {
    uint32_t t = 1024;
    while ((int64_t)t < 2 * n) t <<= 1;
    return t;
}

static size_t grow_offsets(int64_t n, size_t off[8]) {
    const size_t m = (size_t)(n > 0 ? n : 0);
    size_t o = 0;
    off[0] = o; o += align_up(m * sizeof(uint32_t));
    off[1] = o; o += align_up(3 * m * sizeof(int32_t));
    off[2] = o; o += align_up(m * sizeof(uint32_t));
    off[3] = o; o += align_up(m * sizeof(uint32_t));
    off[4] = o; o += align_up(m);
    off[5] = o; o += align_up((size_t)grow_table_size(n) * sizeof(uint32_t));
    size_t boff[7];
    off[6] = o; o += bin_offsets(n, boff);
    off[7] = o;
    return o;
}

static GrowWs grow_view(void *base, int64_t n) {
    size_t off[8];
    grow_offsets(n, off);
    char *b = (char *)base;
    GrowWs w;
    w.slot = (uint32_t *)(b + off[0]); w.xyz = (int32_t *)(b + off[1]); w.rep = (uint32_t *)(b + off[2]);
    w.row = (uint32_t *)(b + off[3]); w.removed = (uint8_t *)(b + off[4]); w.table = (uint32_t *)(b + off[5]);
    w.bin = (void *)(b + off[6]);
    return w;
}

struct GrowSelect {           // the candidate decision of gaussian_model.py:839-846, evaluated per offset slot
    const uint8_t *cand_mask; // optional precomputed mask; when null the three tests below are applied
    const float *grads; float thr;
    const uint8_t *offset_mask;
    const float *rand; float cut;
};

__device__ __forceinline__ bool grow_is_candidate(const GrowSelect &s, uint32_t t) {
    if (s.cand_mask) return s.cand_mask[t] != 0;
    return s.grads[t] >= s.thr && s.offset_mask[t] != 0 && s.rand[t] > s.cut;
}

__device__ __forceinline__ int32_t grow_coord(float a, float cur_size, float inv, int div_mode) {
    const float q = div_mode ? __fdiv_rn(a, cur_size) : __fmul_rn(a, inv);
    return (int32_t)rintf(q);       // torch.round is round-half-to-even; .int() of an integral float
}

__device__ __forceinline__ uint32_t grow_hash(int32_t x, int32_t y, int32_t z) {
    uint32_t h = (uint32_t)x * 73856093u ^ (uint32_t)y * 19349663u ^ (uint32_t)z * 83492791u;
    h ^= h >> 16; h *= 0x7feb352du; h ^= h >> 15; h *= 0x846ca68bu; h ^= h >> 16;
    return h;
}

// warp-aggregated append: returns this lane's position in the list (only meaningful where `take`)
__device__ __forceinline__ uint32_t warp_append(bool take, uint32_t *counter) {
    const uint32_t m = __ballot_sync(0xffffffffu, take);
    uint32_t base = 0;
    if (lane_id() == 0 && m) base = atomicAdd(counter, (uint32_t)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    return base + __popc(m & lanemask_lt());
}

__global__ void __launch_bounds__(256)
grow_count_kernel(uint32_t n_stat, GrowSelect sel, uint32_t *__restrict__ counts) {
    const uint32_t t = blockIdx.x * 256u + threadIdx.x;
    const bool c = t < n_stat && grow_is_candidate(sel, t);
    const uint32_t m = __ballot_sync(0xffffffffu, c);
    __shared__ uint32_t s_n;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    if (lane_id() == 0 && m) atomicAdd(&s_n, (uint32_t)__popc(m));
    __syncthreads();
    if (threadIdx.x == 0 && s_n) atomicAdd(counts, s_n);
}

// one thread per offset slot: candidates append (slot, grid coordinates of anchor + offset * scaling)
__global__ void __launch_bounds__(256)
grow_mark_kernel(uint32_t n_stat, int K, GrowSelect sel, const float *__restrict__ anchors, const float *__restrict__ offsets,
                 const float *__restrict__ scaling, int scale_stride, float cur_size, float inv, int div_mode,
                 uint32_t cap, GrowWs w, uint32_t *__restrict__ cursor) {
    const uint32_t t = blockIdx.x * 256u + threadIdx.x;
    const bool c = t < n_stat && grow_is_candidate(sel, t);
    const uint32_t pos = warp_append(c, cursor);
    if (!c || pos >= cap) return;
    const uint32_t a = t / (uint32_t)K;
    w.slot[pos] = t;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        // anchor + offset * scaling, two separately rounded operations as in the reference's eager torch ops
        const float p = __fadd_rn(anchors[3 * (size_t)a + d], __fmul_rn(offsets[3 * (size_t)t + d], scaling[(size_t)a * scale_stride + d]));
        w.xyz[3 * (size_t)pos + d] = grow_coord(p, cur_size, inv, div_mode);
    }
}

// voxel set: the table holds candidate indices; equal voxels share the first index that claimed the slot
__global__ void __launch_bounds__(256)
grow_insert_kernel(uint32_t n, uint32_t tmask, GrowWs w) {
    const uint32_t c = blockIdx.x * 256u + threadIdx.x;
    if (c >= n) return;
    const int32_t x = w.xyz[3 * (size_t)c], y = w.xyz[3 * (size_t)c + 1], z = w.xyz[3 * (size_t)c + 2];
    uint32_t s = grow_hash(x, y, z) & tmask;
    for (;;) {
        uint32_t cur = w.table[s];
        if (cur == GROW_EMPTY) cur = atomicCAS(&w.table[s], GROW_EMPTY, c);
        if (cur == GROW_EMPTY) { w.rep[c] = c; return; }
        if (w.xyz[3 * (size_t)cur] == x && w.xyz[3 * (size_t)cur + 1] == y && w.xyz[3 * (size_t)cur + 2] == z) { w.rep[c] = cur; return; }
        s = (s + 1) & tmask;
    }
}

// every existing anchor looks its own voxel up and strikes it from the candidates (gaussian_model.py:862,871-884)
__global__ void __launch_bounds__(256)
grow_existing_kernel(int N, const float *__restrict__ anchors, float cur_size, float inv, int div_mode, uint32_t tmask, GrowWs w) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= N) return;
    const int32_t x = grow_coord(anchors[3 * (size_t)i], cur_size, inv, div_mode);
    const int32_t y = grow_coord(anchors[3 * (size_t)i + 1], cur_size, inv, div_mode);
    const int32_t z = grow_coord(anchors[3 * (size_t)i + 2], cur_size, inv, div_mode);
    uint32_t s = grow_hash(x, y, z) & tmask;
    for (;;) {
        const uint32_t cur = w.table[s];
        if (cur == GROW_EMPTY) return;
        if (w.xyz[3 * (size_t)cur] == x && w.xyz[3 * (size_t)cur + 1] == y && w.xyz[3 * (size_t)cur + 2] == z) { w.removed[cur] = 1; return; }
        s = (s + 1) & tmask;
    }
}

// surviving voxels -> (key = z, value = representative) pairs for the first sort
__global__ void __launch_bounds__(256)
grow_survivors_kernel(uint32_t n, GrowWs w, uint64_t *__restrict__ keys, uint32_t *__restrict__ vals, uint32_t *__restrict__ count) {
    const uint32_t c = blockIdx.x * 256u + threadIdx.x;
    const bool take = c < n && w.rep[c] == c && !w.removed[c];
    const uint32_t pos = warp_append(take, count);
    if (!take) return;
    keys[pos] = (uint64_t)((uint32_t)w.xyz[3 * (size_t)c + 2] ^ 0x80000000u);
    vals[pos] = c;
}

// second sort key: (x, y) of the representatives in their z-sorted order (LSD: least significant column first)
__global__ void __launch_bounds__(256)
grow_rekey_kernel(uint32_t n, const int32_t *__restrict__ xyz, const uint32_t *__restrict__ vals, uint64_t *__restrict__ keys) {
    const uint32_t i = blockIdx.x * 256u + threadIdx.x;
    if (i >= n) return;
    const uint32_t c = vals[i];
    keys[i] = ((uint64_t)((uint32_t)xyz[3 * (size_t)c] ^ 0x80000000u) << 32) | (uint64_t)((uint32_t)xyz[3 * (size_t)c + 1] ^ 0x80000000u);
}

__global__ void __launch_bounds__(256)
grow_finalize_kernel(uint32_t n_new, int F, const uint32_t *__restrict__ sorted_rep, GrowWs w, float cur_size,
                     float *__restrict__ new_anchor, float *__restrict__ new_feat) {
    const uint32_t u = blockIdx.x * 256u + threadIdx.x;
    if (u >= n_new) return;
    const uint32_t c = sorted_rep[u];
    w.row[c] = u;
#pragma unroll
    for (int d = 0; d < 3; ++d) new_anchor[3 * (size_t)u + d] = __fmul_rn((float)w.xyz[3 * (size_t)c + d], cur_size);
    for (int f = 0; f < F; ++f) new_feat[(size_t)u * F + f] = __int_as_float(0xff800000);   // -inf
}

__device__ __forceinline__ void atomic_max_float(float *addr, float v) {
    if (!(__float_as_uint(v) >> 31)) atomicMax((int *)addr, __float_as_int(v));
    else atomicMin((unsigned int *)addr, __float_as_uint(v));
}

// per-voxel, per-channel maximum of the candidates' anchor features (torch_scatter.scatter_max, :895-897); warp per candidate
__global__ void __launch_bounds__(256)
grow_featmax_kernel(uint32_t n, int K, int F, GrowWs w, const float *__restrict__ anchor_feat, float *__restrict__ new_feat) {
    const uint32_t c = (blockIdx.x * 256u + threadIdx.x) >> 5;
    if (c >= n) return;
    const uint32_t r = w.rep[c];
    if (w.removed[r]) return;
    const size_t row = w.row[r], a = w.slot[c] / (uint32_t)K;
    for (int f = lane_id(); f < F; f += 32) atomic_max_float(new_feat + row * F + f, __ldg(anchor_feat + a * F + f));
}

}  // namespace splatco

using namespace splatco;

static GrowSelect make_select(const uint8_t *cand_mask, const float *grads, float thr, const uint8_t *offset_mask, const float *rand,
                              float cut) {
    GrowSelect s; s.cand_mask = cand_mask; s.grads = grads; s.thr = thr; s.offset_mask = offset_mask; s.rand = rand; s.cut = cut;
    return s;
}

static int grow_check_select(int64_t n_stat, const uint8_t *cand_mask, const float *grads, const uint8_t *offset_mask, const float *rand) {
    SPLATCO_REQUIRE(n_stat >= 0 && n_stat < 0x7fffffff, "anchor growing: bad slot count %lld", (long long)n_stat);
    SPLATCO_REQUIRE(n_stat == 0 || cand_mask || (grads && offset_mask && rand), "anchor growing: need cand_mask or (grads, offset_mask, rand)");
    return 0;
}

extern "C" size_t splatco_grow_ws_bytes(int64_t n_cand) { size_t off[8]; return grow_offsets(n_cand, off); }

extern "C" int splatco_grow_count(int64_t n_stat, const uint8_t *cand_mask, const float *grads, float threshold,
                                  const uint8_t *offset_mask, const float *rand, float rand_cut, uint32_t *counts, void *stream) {
    if (grow_check_select(n_stat, cand_mask, grads, offset_mask, rand)) return -1;
    SPLATCO_REQUIRE(counts, "grow_count: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    SPLATCO_CHECK_CUDA(cudaMemsetAsync(counts, 0, 4 * sizeof(uint32_t), st));
    if (n_stat == 0) return 0;
    grow_count_kernel<<<(uint32_t)ceil_div64(n_stat, 256), 256, 0, st>>>((uint32_t)n_stat, make_select(cand_mask, grads, threshold, offset_mask, rand, rand_cut), counts);
    SPLATCO_CHECK_LAUNCH();
    return 0;
}

extern "C" int splatco_grow_unique(int N, int K, int64_t n_stat, const float *anchors, const float *offsets, const float *scaling,
                                   int scale_stride, const uint8_t *cand_mask, const float *grads, float threshold,
                                   const uint8_t *offset_mask, const float *rand, float rand_cut, float cur_size, int div_mode,
                                   int64_t n_cand, void *ws, uint32_t *counts, void *stream) {
    if (grow_check_select(n_stat, cand_mask, grads, offset_mask, rand)) return -1;
    SPLATCO_REQUIRE(N >= 0 && K >= 1 && n_stat <= (int64_t)N * K && n_cand >= 0 && n_cand <= n_stat, "grow_unique: bad sizes N=%d K=%d slots=%lld candidates=%lld",
                    N, K, (long long)n_stat, (long long)n_cand);
    SPLATCO_REQUIRE(cur_size > 0.f && scale_stride >= 3, "grow_unique: bad cur_size / scale_stride");
    SPLATCO_REQUIRE(counts, "grow_unique: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    SPLATCO_CHECK_CUDA(cudaMemsetAsync(counts + 1, 0, 3 * sizeof(uint32_t), st));
    if (n_cand == 0) return 0;
    SPLATCO_REQUIRE(anchors && offsets && scaling && ws, "grow_unique: null pointer");
    GrowWs w = grow_view(ws, n_cand);
    const uint32_t tsize = grow_table_size(n_cand), n = (uint32_t)n_cand;
    const float inv = 1.0f / cur_size;
    SPLATCO_CHECK_CUDA(cudaMemsetAsync(w.table, 0xff, (size_t)tsize * sizeof(uint32_t), st));
    SPLATCO_CHECK_CUDA(cudaMemsetAsync(w.removed, 0, (size_t)n, st));
    grow_mark_kernel<<<(uint32_t)ceil_div64(n_stat, 256), 256, 0, st>>>((uint32_t)n_stat, K, make_select(cand_mask, grads, threshold, offset_mask, rand, rand_cut),
                                                                        anchors, offsets, scaling, scale_stride, cur_size, inv, div_mode, n, w, counts + 1);
    SPLATCO_CHECK_LAUNCH();
    grow_insert_kernel<<<ceil_div((int)n, 256), 256, 0, st>>>(n, tsize - 1, w);
    SPLATCO_CHECK_LAUNCH();
    if (N > 0) {
        grow_existing_kernel<<<ceil_div(N, 256), 256, 0, st>>>(N, anchors, cur_size, inv, div_mode, tsize - 1, w);
        SPLATCO_CHECK_LAUNCH();
    }
    BinWs b = bin_view(w.bin, n_cand);
    grow_survivors_kernel<<<ceil_div((int)n, 256), 256, 0, st>>>(n, w, b.keys[0], b.vals[0], counts + 2);
    SPLATCO_CHECK_LAUNCH();
    return 0;
}

extern "C" int splatco_grow_emit(int K, int F, float cur_size, int64_t n_cand, int64_t n_new, void *ws, const float *anchor_feat,
                                 float *new_anchor, float *new_feat, void *stream) {
    SPLATCO_REQUIRE(K >= 1 && F >= 1 && n_cand >= 0 && n_new >= 0 && n_new <= n_cand, "grow_emit: bad sizes");
    if (n_new == 0) return 0;
    SPLATCO_REQUIRE(ws && anchor_feat && new_anchor && new_feat, "grow_emit: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    GrowWs w = grow_view(ws, n_cand);
    BinWs b = bin_view(w.bin, n_cand);
    const uint32_t n = (uint32_t)n_new;
    int cur = 0;
    if (radix_sort_pairs(b, n, 32, &cur, st)) return -2;                       // by z
    grow_rekey_kernel<<<ceil_div((int)n, 256), 256, 0, st>>>(n, w.xyz, b.vals[cur], b.keys[cur]);
    SPLATCO_CHECK_LAUNCH();
    if (radix_sort_pairs(b, n, 64, &cur, st)) return -2;                       // then (stably) by (x, y)
    grow_finalize_kernel<<<ceil_div((int)n, 256), 256, 0, st>>>(n, F, b.vals[cur], w, cur_size, new_anchor, new_feat);
    SPLATCO_CHECK_LAUNCH();
    grow_featmax_kernel<<<(uint32_t)ceil_div64(n_cand * 32, 256), 256, 0, st>>>((uint32_t)n_cand, K, F, w, anchor_feat, new_feat);
    SPLATCO_CHECK_LAUNCH();
    return 0;
}
