// Tile binning: duplicateWithKeys, a hand-written stable LSD radix sort of (tile|depth) 64-bit keys
// with 32-bit payloads, and identifyTileRanges.  Replaces upstream duplicateWithKeys /
// cub::DeviceRadixSort::SortPairs / identifyTileRanges [SURVEY.md Appendix A.3].
//
// All integer work, HBM-bound (emit 12 B·R, sort p·24 B·R + 8 B·R, ranges 8 B·R + 8 B·T).
// Sort = p passes of 8 bits over bits [0, 32 + bits(T-1)); each pass is
//   (1) per-CTA digit histogram (2048 pairs / CTA),
//   (2) per-digit exclusive scan across CTAs,
//   (3) stable scatter: warp-level match_any ranking, re-ordered through shared memory so that a
//       CTA writes each digit run with consecutive threads (coalesced 8-byte / 4-byte stores).
// Stability (ties keep Gaussian-index order, and (y,x) emission order) is what makes the resulting
// permutation bit-identical to the reference's stable SortPairs.
#include "common.cuh"

namespace splatco {

__device__ __forceinline__ void rect_from_rec(float px, float py, int radius, int gx, int gy, int &r0x,
                                              int &r0y, int &r1x, int &r1y) {
    const float rf = (float)radius;
    r0x = min(gx, max(0, __float2int_rz(__fmul_rn(__fsub_rn(px, rf), 0.0625f))));
    r0y = min(gy, max(0, __float2int_rz(__fmul_rn(__fsub_rn(py, rf), 0.0625f))));
    r1x = min(gx, max(0, __float2int_rz(__fmul_rn(__fadd_rn(__fadd_rn(px, rf), 15.0f), 0.0625f))));
    r1y = min(gy, max(0, __float2int_rz(__fmul_rn(__fadd_rn(__fadd_rn(py, rf), 15.0f), 0.0625f))));
}

// One Gaussian per thread, same 256-wide partition as preprocess so the CTA-local exclusive scan of
// tiles_touched plus block_offsets[blockIdx] reproduces the global offsets without materialising them.
__global__ void __launch_bounds__(PRE_THREADS)
duplicate_with_keys_kernel(int P, const int32_t *__restrict__ radii, const float4 *__restrict__ rec,
                           const uint32_t *__restrict__ tiles, const uint32_t *__restrict__ block_offsets,
                           int gx, int gy, uint64_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    __shared__ uint32_t s_warp[PRE_THREADS / 32];
    const int i = blockIdx.x * PRE_THREADS + threadIdx.x;
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    const uint32_t t = i < P ? tiles[i] : 0u;
    uint32_t inc = t;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t n = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= (uint32_t)d) inc += n; }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    uint32_t woff = 0;
#pragma unroll
    for (int w = 0; w < PRE_THREADS / 32; ++w) woff += (w < (int)warp) ? s_warp[w] : 0u;
    if (t == 0) return;
    uint32_t off = block_offsets[blockIdx.x] + woff + inc - t;
    const float4 r0 = rec[3 * (size_t)i];
    const float4 r2 = rec[3 * (size_t)i + 2];
    int r0x, r0y, r1x, r1y;
    rect_from_rec(r0.x, r0.y, radii[i], gx, gy, r0x, r0y, r1x, r1y);
    const uint64_t depth_bits = (uint64_t)__float_as_uint(r2.y);
    for (int y = r0y; y < r1y; ++y)
        for (int x = r0x; x < r1x; ++x) {
            const uint64_t key = ((uint64_t)(uint32_t)(y * gx + x) << 32) | depth_bits;
            keys[off] = key;
            vals[off] = (uint32_t)i;
            ++off;
        }
}

// ---- radix sort ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SORT_THREADS)
sort_hist_kernel(const uint64_t *__restrict__ keys, uint32_t n, int shift, uint32_t mask,
                 uint32_t *__restrict__ block_hist, int nblocks) {
    __shared__ uint32_t s_hist[256];
    s_hist[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x * SORT_TILE;
#pragma unroll
    for (int it = 0; it < SORT_ITEMS; ++it) {
        const uint32_t idx = base + it * SORT_THREADS + threadIdx.x;
        const bool valid = idx < n;
        const uint32_t digit = valid ? (uint32_t)(keys[idx] >> shift) & mask : 0xffffffffu;
        const uint32_t peers = __match_any_sync(0xffffffffu, digit);
        if (valid && (peers & lanemask_lt()) == 0) atomicAdd(&s_hist[digit], __popc(peers));
    }
    __syncthreads();
    block_hist[(size_t)threadIdx.x * nblocks + blockIdx.x] = s_hist[threadIdx.x];
}

// grid = 256 CTAs (one per digit): exclusive scan of that digit's per-CTA counts, in place.
__global__ void __launch_bounds__(256)
sort_scan_kernel(uint32_t *__restrict__ block_hist, int nblocks, uint32_t *__restrict__ bin_totals) {
    __shared__ uint32_t s_warp[8];
    __shared__ uint32_t s_carry;
    uint32_t *row = block_hist + (size_t)blockIdx.x * nblocks;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    for (int base = 0; base < nblocks; base += 256) {
        const int i = base + threadIdx.x;
        const uint32_t v = i < nblocks ? row[i] : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t n = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= (uint32_t)d) inc += n; }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        uint32_t woff = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) woff += (w < (int)warp) ? s_warp[w] : 0u;
        const uint32_t excl = s_carry + woff + inc - v;
        if (i < nblocks) row[i] = excl;
        __syncthreads();
        if (threadIdx.x == 255) s_carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) bin_totals[blockIdx.x] = s_carry;
}

__global__ void __launch_bounds__(SORT_THREADS)
sort_scatter_kernel(const uint64_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                    uint64_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, uint32_t n,
                    int shift, uint32_t mask, const uint32_t *__restrict__ block_hist, int nblocks,
                    const uint32_t *__restrict__ bin_totals) {
    __shared__ uint32_t s_warp_cnt[SORT_THREADS / 32][256];   // 8 KB
    __shared__ uint32_t s_digit_start[256];
    __shared__ uint32_t s_gofs[256];
    __shared__ uint32_t s_scan[8];
    __shared__ uint64_t s_keys[SORT_TILE];                    // 16 KB
    __shared__ uint32_t s_vals[SORT_TILE];                    //  8 KB
    const uint32_t tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
#pragma unroll
    for (int w = 0; w < SORT_THREADS / 32; ++w) s_warp_cnt[w][tid] = 0;

    // global base of each digit = exclusive scan of bin_totals (256 values, recomputed per CTA)
    const uint32_t tot = bin_totals[tid];
    uint32_t inc = tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= (uint32_t)d) inc += v; }
    if (lane == 31) s_scan[warp] = inc;
    __syncthreads();
    uint32_t woff = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) woff += (w < (int)warp) ? s_scan[w] : 0u;
    const uint32_t bin_base = woff + inc - tot;
    const uint32_t cta_prefix = block_hist[(size_t)tid * nblocks + blockIdx.x];

    // load (warp-striped: warp w owns pairs [w*256, w*256+256) of the CTA tile, item it at +it*32+lane)
    const uint32_t wbase = blockIdx.x * SORT_TILE + warp * (32 * SORT_ITEMS);
    uint64_t key[SORT_ITEMS];
    uint32_t val[SORT_ITEMS], rank[SORT_ITEMS];
#pragma unroll
    for (int it = 0; it < SORT_ITEMS; ++it) {
        const uint32_t idx = wbase + it * 32 + lane;
        const bool valid = idx < n;
        key[it] = valid ? keys_in[idx] : ~0ull;
        val[it] = valid ? vals_in[idx] : 0u;
    }
    // stable rank inside the warp, item by item
#pragma unroll
    for (int it = 0; it < SORT_ITEMS; ++it) {
        const uint32_t digit = (uint32_t)(key[it] >> shift) & mask;
        const uint32_t peers = __match_any_sync(0xffffffffu, digit);
        const uint32_t pre = s_warp_cnt[warp][digit];
        const uint32_t r = __popc(peers & lanemask_lt());
        rank[it] = pre + r;
        __syncwarp();
        if (r == 0) s_warp_cnt[warp][digit] = pre + __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    // per digit (= tid): exclusive scan over warps, then over digits
    uint32_t dsum = 0;
#pragma unroll
    for (int w = 0; w < SORT_THREADS / 32; ++w) { const uint32_t c = s_warp_cnt[w][tid]; s_warp_cnt[w][tid] = dsum; dsum += c; }
    uint32_t dinc = dsum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, dinc, d); if (lane >= (uint32_t)d) dinc += v; }
    __syncthreads();          // everyone is done reading s_scan from the first use
    if (lane == 31) s_scan[warp] = dinc;
    __syncthreads();
    uint32_t doff = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) doff += (w < (int)warp) ? s_scan[w] : 0u;
    const uint32_t dstart = doff + dinc - dsum;
    s_digit_start[tid] = dstart;
    s_gofs[tid] = bin_base + cta_prefix - dstart;     // global index = s_gofs[digit] + local position
    __syncthreads();
#pragma unroll
    for (int it = 0; it < SORT_ITEMS; ++it) {
        const uint32_t digit = (uint32_t)(key[it] >> shift) & mask;
        const uint32_t pos = s_digit_start[digit] + s_warp_cnt[warp][digit] + rank[it];
        s_keys[pos] = key[it];
        s_vals[pos] = val[it];
    }
    __syncthreads();
    const uint32_t cta_base = blockIdx.x * SORT_TILE;
    const uint32_t count = min((uint32_t)SORT_TILE, n - cta_base);
    for (uint32_t k = tid; k < count; k += SORT_THREADS) {
        const uint64_t kk = s_keys[k];
        const uint32_t digit = (uint32_t)(kk >> shift) & mask;
        const uint32_t dst = s_gofs[digit] + k;
        keys_out[dst] = kk;
        vals_out[dst] = s_vals[k];
    }
}

__global__ void __launch_bounds__(256)
identify_tile_ranges_kernel(uint32_t n, const uint64_t *__restrict__ keys, int2 *__restrict__ ranges) {
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const uint32_t t = (uint32_t)(keys[i] >> 32);
    if (i == 0) ranges[t].x = 0;
    else {
        const uint32_t tp = (uint32_t)(keys[i - 1] >> 32);
        if (tp != t) { ranges[tp].y = (int)i; ranges[t].x = (int)i; }
    }
    if (i == n - 1) ranges[t].y = (int)n;
}

}  // namespace splatco

using namespace splatco;

static int check_R(int64_t R) {
    SPLATCO_REQUIRE(R >= 0 && R < (int64_t)0x7fffffff, "instance count R=%lld out of range", (long long)R);
    return 0;
}

extern "C" int splatco_duplicate_with_keys(int P, int64_t R, int H, int W, const int32_t *radii,
                                           const void *geom, void *binning, void *stream) {
    if (check_R(R)) return -1;
    if (P == 0 || R == 0) return 0;
    SPLATCO_REQUIRE(radii && geom && binning, "duplicate_with_keys: null pointer");
    GeomWs g = geom_view(const_cast<void *>(geom), P);
    BinWs b = bin_view(binning, R);
    const int gx = ceil_div(W, TILE), gy = ceil_div(H, TILE);
    duplicate_with_keys_kernel<<<ceil_div(P, PRE_THREADS), PRE_THREADS, 0, (cudaStream_t)stream>>>(
        P, radii, g.rec, g.tiles, g.block_offsets, gx, gy, b.keys[0], b.vals[0]);
    SPLATCO_CHECK_LAUNCH();
    return 0;
}

extern "C" int splatco_sort_pairs(int64_t R, int H, int W, void *binning, void *stream) {
    if (check_R(R)) return -1;
    if (R == 0) return 0;
    SPLATCO_REQUIRE(binning, "sort_pairs: null pointer");
    BinWs b = bin_view(binning, R);
    cudaStream_t st = (cudaStream_t)stream;
    const int T = ceil_div(W, TILE) * ceil_div(H, TILE);
    const int bits = 32 + tile_bits(T);
    const int passes = (bits + 7) / 8;
    const int nblocks = (int)ceil_div64(R, SORT_TILE);
    int cur = 0;
    for (int p = 0; p < passes; ++p) {
        const int shift = 8 * p;
        const int nb = bits - shift < 8 ? bits - shift : 8;
        const uint32_t mask = (1u << nb) - 1u;
        sort_hist_kernel<<<nblocks, SORT_THREADS, 0, st>>>(b.keys[cur], (uint32_t)R, shift, mask, b.hist, nblocks);
        SPLATCO_CHECK_LAUNCH();
        sort_scan_kernel<<<256, 256, 0, st>>>(b.hist, nblocks, b.bin_totals);
        SPLATCO_CHECK_LAUNCH();
        sort_scatter_kernel<<<nblocks, SORT_THREADS, 0, st>>>(b.keys[cur], b.vals[cur], b.keys[cur ^ 1], b.vals[cur ^ 1],
                                                              (uint32_t)R, shift, mask, b.hist, nblocks, b.bin_totals);
        SPLATCO_CHECK_LAUNCH();
        cur ^= 1;
    }
    return 0;
}

extern "C" int splatco_sorted_buffer_index(int H, int W) { return sort_passes(H, W) & 1; }

extern "C" int splatco_identify_tile_ranges(int64_t R, int H, int W, const void *binning, void *image,
                                            void *stream) {
    if (check_R(R)) return -1;
    SPLATCO_REQUIRE(image, "identify_tile_ranges: null pointer");
    ImgWs im = img_view(image, H, W);
    cudaStream_t st = (cudaStream_t)stream;
    const int T = ceil_div(W, TILE) * ceil_div(H, TILE);
    SPLATCO_CHECK_CUDA(cudaMemsetAsync(im.ranges, 0, sizeof(int2) * (size_t)T, st));
    if (R == 0) return 0;
    SPLATCO_REQUIRE(binning, "identify_tile_ranges: null pointer");
    BinWs b = bin_view(const_cast<void *>(binning), R);
    const int s = splatco_sorted_buffer_index(H, W);
    identify_tile_ranges_kernel<<<(unsigned)ceil_div64(R, 256), 256, 0, st>>>((uint32_t)R, b.keys[s], im.ranges);
    SPLATCO_CHECK_LAUNCH();
    return 0;
}

extern "C" int splatco_binning(int P, int64_t R, int H, int W, const int32_t *radii, const void *geom,
                               void *binning, void *image, void *stream) {
    int rc = splatco_duplicate_with_keys(P, R, H, W, radii, geom, binning, stream);
    if (rc) return rc;
    rc = splatco_sort_pairs(R, H, W, binning, stream);
    if (rc) return rc;
    return splatco_identify_tile_ranges(R, H, W, binning, image, stream);
}
