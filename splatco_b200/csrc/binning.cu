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


// =====================================================================================================
// Tile-segmented binning (the path splatco_binning takes).  The 64-bit key is (tile << 32 | depth): the
// tile part only says WHICH segment of the list an instance belongs to.  So instead of six global radix
// passes over all R pairs, ONE multi-way partition by tile and a per-tile sort in shared memory:
//   tile_hist     chunks of 4096+ Gaussians, one CTA each: per-tile instance counts in a shared-memory histogram;
//                 each non-empty (chunk, tile) count claims its slice of the tile's segment with one global atomic
//                 on the tile total, the returned bases are written out as one row per chunk
//   tile_scan     one CTA: exclusive scan of the T totals -> ranges
//   tile_scatter  same chunks: slot = range start + chunk base + shared-memory atomic; stores (depth << 32 | id)
//   tile_sort_*   one CTA per tile: LSD radix sort of the segment on the depth bits in shared memory
//                 (constant digits skipped), equal-depth runs then ordered by Gaussian id
// The arrival order inside a segment is arbitrary, but (depth, id) is unique within a tile, so the
// result is exactly what a stable sort of (tile | depth) keys in emission order produces: ties in depth
// keep ascending Gaussian id, the reference's emission order [SURVEY.md Appendix A.3].
// Traffic: 8 B*R scattered + 8 B*R read + 12 B*R written, L2-resident at the BASELINE sizes.
// =====================================================================================================
constexpr int TCHUNK = 4096;                       // Gaussians per CTA of tile_hist / tile_scatter (doubled until the
                                                   // [chunks][tiles] histogram matrix fits the spare value buffer: tile_chunk())
constexpr int TCHUNK_MAX = 1 << 17;
constexpr int TCHUNK_THREADS = 512;

__device__ __forceinline__ bool tile_rect(int i, int P, const int32_t *__restrict__ radii, const float4 *__restrict__ rec,
                                          int gx, int gy, int &r0x, int &r0y, int &r1x, int &r1y) {
    if (i >= P) return false;
    const int radius = radii[i];
    if (radius <= 0) return false;
    const float4 r0 = rec[3 * (size_t)i];
    rect_from_rec(r0.x, r0.y, radius, gx, gy, r0x, r0y, r1x, r1y);
    return true;
}

__global__ void __launch_bounds__(TCHUNK_THREADS)
tile_hist_kernel(int P, int T, int chunk, const int32_t *__restrict__ radii, const float4 *__restrict__ rec, int gx, int gy,
                 uint32_t *__restrict__ chunk_base /*[nchunks][T]*/, uint32_t *__restrict__ tile_count /*[T], zeroed*/,
                 uint32_t *__restrict__ work /*[8]: list lengths [0..2], queue heads [4..6]*/) {
    extern __shared__ uint32_t s_hist[];
    for (int t = threadIdx.x; t < T; t += TCHUNK_THREADS) s_hist[t] = 0;
    __syncthreads();
    const int base = blockIdx.x * chunk;
#pragma unroll 1
    for (int k = 0; k < chunk / TCHUNK_THREADS; ++k) {
        const int i = base + k * TCHUNK_THREADS + threadIdx.x;
        int r0x, r0y, r1x, r1y;
        if (!tile_rect(i, P, radii, rec, gx, gy, r0x, r0y, r1x, r1y)) continue;
        for (int y = r0y; y < r1y; ++y)
            for (int x = r0x; x < r1x; ++x) atomicAdd(&s_hist[y * gx + x], 1u);
    }
    __syncthreads();
    // this chunk's slice of every tile's segment: claimed with one atomic per non-empty (chunk, tile).  The order of the
    // chunks inside a segment is whatever order the atomics arrive in -- the per-tile sort that follows makes the list
    // independent of it ((depth, id) is unique within a tile) -- and no second pass over the [chunks][tiles] matrix is
    // needed to turn counts into offsets.
    if (blockIdx.x == 0 && threadIdx.x < 8) work[threadIdx.x] = 0;
    uint32_t *row = chunk_base + (size_t)blockIdx.x * T;
    for (int t = threadIdx.x; t < T; t += TCHUNK_THREADS) {
        const uint32_t c = s_hist[t];
        row[t] = c ? atomicAdd(&tile_count[t], c) : 0u;
    }
}

// one CTA; ranges[t] = [start, end) (zero for empty tiles, as identifyTileRanges leaves them), cursor[t] = start
__global__ void __launch_bounds__(1024)
tile_scan_kernel(int T, const uint32_t *__restrict__ tile_count, int2 *__restrict__ ranges, uint32_t *__restrict__ cursor,
                 uint32_t capacity, uint32_t *__restrict__ work, uint32_t *__restrict__ lists /*[3][T]*/, uint32_t lim_s,
                 uint32_t lim_l) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    for (int base = 0; base < T; base += 1024) {
        const int t = base + threadIdx.x;
        const uint32_t v = t < T ? tile_count[t] : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t n = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= (uint32_t)d) inc += n; }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = s_warp[lane], winc = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { uint32_t n = __shfl_up_sync(0xffffffffu, winc, d); if (lane >= (uint32_t)d) winc += n; }
            s_warp[lane] = winc - w;
        }
        __syncthreads();
        const uint32_t excl = s_carry + s_warp[warp] + inc - v;
        if (t < T) {
            // clamped to the buffers' capacity: a caller that sized them from a guess gets a truncated but
            // in-bounds list and re-runs the stage once it knows R
            const uint32_t s0 = min(excl, capacity), e0 = min(excl + v, capacity);
            ranges[t] = e0 > s0 ? make_int2((int)s0, (int)e0) : make_int2(0, 0);
            cursor[t] = excl;
            // work list of the sort kernel that handles this segment length (order inside a list is irrelevant)
            const uint32_t n = e0 - s0;
            if (n > 1) {
                const int cls = n <= lim_s ? 0 : (n <= lim_l ? 1 : 2);
                lists[(size_t)cls * T + atomicAdd(&work[cls], 1u)] = (uint32_t)t;
            } else if (n == 1) {
                lists[(size_t)0 * T + atomicAdd(&work[0], 1u)] = (uint32_t)t;
            }
        }
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = excl + v;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(TCHUNK_THREADS)
tile_scatter_kernel(int P, int T, int chunk, const int32_t *__restrict__ radii, const float4 *__restrict__ rec, int gx, int gy,
                    const uint32_t *__restrict__ chunk_base /*[nchunks][T]*/, const uint32_t *__restrict__ tile_start,
                    uint64_t *__restrict__ entries, uint32_t capacity) {
    extern __shared__ uint32_t s_next[];
    const uint32_t *row = chunk_base + (size_t)blockIdx.x * T;
    for (int t = threadIdx.x; t < T; t += TCHUNK_THREADS) s_next[t] = tile_start[t] + row[t];
    __syncthreads();
    const int base = blockIdx.x * chunk;
#pragma unroll 1
    for (int k = 0; k < chunk / TCHUNK_THREADS; ++k) {
        const int i = base + k * TCHUNK_THREADS + threadIdx.x;
        int r0x, r0y, r1x, r1y;
        if (!tile_rect(i, P, radii, rec, gx, gy, r0x, r0y, r1x, r1y)) continue;
        const uint64_t e = ((uint64_t)__float_as_uint(rec[3 * (size_t)i + 2].y) << 32) | (uint32_t)i;
        for (int y = r0y; y < r1y; ++y)
            for (int x = r0x; x < r1x; ++x) {
                const uint32_t slot = atomicAdd(&s_next[y * gx + x], 1u);
                if (slot < capacity) entries[slot] = e;
            }
    }
}

// ---- per-tile sort, segments that fit in shared memory: LSD radix on the 32 depth bits ----------------------
// ITEMS consecutive-by-32 elements per lane, warp w owns the contiguous slice [w*32*ITEMS, (w+1)*32*ITEMS): the
// usual stable ranking (match_any inside the warp, running per-(warp, digit) counters across its items).
template <int NT, int ITEMS>
__global__ void __launch_bounds__(NT)
tile_sort_radix_kernel(const int2 *__restrict__ ranges, const uint64_t *__restrict__ entries,
                       uint64_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, const uint32_t *__restrict__ list,
                       uint32_t *__restrict__ work, int cls) {
    constexpr int NW = NT / 32, CAP = NT * ITEMS;
    extern __shared__ uint32_t s_dyn[];
    uint32_t *s_cnt = s_dyn + 4 * CAP;                         // [NW][256]; before it: [key CAP | val CAP] x 2
    __shared__ uint32_t s_scan[NW];
    __shared__ int s_skip;
    __shared__ uint32_t s_item, s_vary[2];
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t count = work[cls];
  for (;;) {                                                   // persistent CTA: next tile of this size class
    __syncthreads();
    if (tid == 0) s_item = atomicAdd(&work[4 + cls], 1u);
    __syncthreads();
    if (s_item >= count) break;
    const uint32_t tile = list[s_item];
    const int2 rg = ranges[tile];
    const uint32_t n = (uint32_t)(rg.y - rg.x);
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // The lowest depth byte is left out of the radix passes: after the passes on bytes 1..3 only keys that agree in their
    // top 24 bits can still be out of order -- a handful per tile (2048 depths spread over >= 2^20 values) -- and the
    // run-fixing step below orders those by (full key, id).  A tile whose depths collide massively in the top 24 bits
    // is sorted again with all four passes.
    int first_pass = n >= 64u ? 1 : 0;
    uint32_t *key, *val;
  for (;;) {
    uint32_t k_or = 0u, k_and = 0xffffffffu;
    for (uint32_t i = tid; i < n; i += NT) {
        const uint64_t e = entries[rg.x + i];
        const uint32_t kd = (uint32_t)(e >> 32);
        s_dyn[i] = kd;
        s_dyn[CAP + i] = (uint32_t)e;
        k_or |= kd; k_and &= kd;
    }
    // bits that differ between any two keys of the tile: digits made of constant bits need no pass (depths of one
    // tile share their sign/exponent byte almost always)
    k_or = __reduce_or_sync(0xffffffffu, k_or);
    k_and = __reduce_and_sync(0xffffffffu, k_and);
    if (tid == 0) { s_vary[0] = 0u; s_vary[1] = 0xffffffffu; }
    __syncthreads();
    if (lane == 0) { atomicOr(&s_vary[0], k_or); atomicAnd(&s_vary[1], k_and); }
    __syncthreads();
    const uint32_t varying = s_vary[0] & ~s_vary[1];
    int cur = 0;
    const uint32_t wbase = warp * 32 * ITEMS;
#pragma unroll 1
    for (int pass = first_pass; pass < 4; ++pass) {
        const int shift = 8 * pass;
        if (((varying >> shift) & 255u) == 0u) continue;       // uniform across the CTA
        __syncthreads();                                       // the previous scatter (or the load) is complete
        for (uint32_t i = tid; i < NW * 256; i += NT) s_cnt[i] = 0;
        __syncthreads();
        const uint32_t *key = s_dyn + cur * 2 * CAP, *val = key + CAP;
        uint32_t k[ITEMS], rank[ITEMS];
#pragma unroll
        for (int it = 0; it < ITEMS; ++it) {
            const uint32_t idx = wbase + it * 32 + lane;
            const bool valid = idx < n;
            k[it] = valid ? key[idx] : 0xffffffffu;
            const uint32_t digit = valid ? (k[it] >> shift) & 255u : 256u;
            const uint32_t peers = __match_any_sync(0xffffffffu, digit);
            const uint32_t r = __popc(peers & lanemask_lt());
            uint32_t pre = 0;
            if (valid) pre = s_cnt[warp * 256 + digit];
            rank[it] = pre + r;
            __syncwarp();
            if (valid && r == 0) s_cnt[warp * 256 + digit] = pre + __popc(peers);
            __syncwarp();
        }
        __syncthreads();
        // digit d = tid (< 256): exclusive scan over warps, then over digits
        uint32_t dsum = 0, dinc = 0;
        if (tid < 256) {
#pragma unroll
            for (int w = 0; w < NW; ++w) { const uint32_t c = s_cnt[w * 256 + tid]; s_cnt[w * 256 + tid] = dsum; dsum += c; }
            dinc = dsum;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, dinc, d); if (lane >= (uint32_t)d) dinc += v; }
            if (lane == 31) s_scan[warp] = dinc;
            if (tid == 0) s_skip = 0;
        }
        __syncthreads();
        if (tid < 256) {
            uint32_t doff = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) doff += (w < (int)warp) ? s_scan[w] : 0u;
            const uint32_t dstart = doff + dinc - dsum;
#pragma unroll
            for (int w = 0; w < NW; ++w) s_cnt[w * 256 + tid] += dstart;
            if (dsum == n) s_skip = 1;                         // every key has this digit: the pass is the identity
        }
        __syncthreads();
        if (s_skip) continue;
        uint32_t *okey = s_dyn + (cur ^ 1) * 2 * CAP, *oval = okey + CAP;
#pragma unroll
        for (int it = 0; it < ITEMS; ++it) {
            const uint32_t idx = wbase + it * 32 + lane;
            if (idx < n) {
                const uint32_t pos = s_cnt[warp * 256 + ((k[it] >> shift) & 255u)] + rank[it];
                okey[pos] = k[it];
                oval[pos] = val[idx];
            }
        }
        cur ^= 1;
    }
    __syncthreads();
    // runs of keys that agree above the bits the passes covered (equal depths, or equal top 24 bits when the low byte was
    // left out): order by (key, Gaussian id); the thread at the head of a run sorts it
    key = s_dyn + cur * 2 * CAP; val = key + CAP;
    const int cs = 8 * first_pass;
    int bad = 0;
    // two phases, so that no thread reads keys while a run head permutes its run: first every run head notes where its run
    // ends (in the inactive key buffer), then the heads sort
    uint32_t *run_end = s_dyn + (cur ^ 1) * 2 * CAP;
    for (uint32_t i = tid; i < n; i += NT) {
        uint32_t e = 0;
        if (i + 1 < n) {
            const uint32_t ki = key[i] >> cs;
            if (ki == (key[i + 1] >> cs) && (i == 0 || (key[i - 1] >> cs) != ki)) {
                e = i + 1;
                while (e + 1 < n && (key[e + 1] >> cs) == ki) ++e;
            }
        }
        run_end[i] = e;
    }
    __syncthreads();
    for (uint32_t i = tid; i + 1 < n; i += NT) {
        const uint32_t e = run_end[i];
        if (e == 0) continue;
        if (first_pass && e - i >= 32u) { bad = 1; continue; }
        for (uint32_t a = i + 1; a <= e; ++a) {                // insertion sort of (key, val)[i..e]
            const uint32_t kk = key[a], v = val[a];
            uint32_t b = a;
            while (b > i && (key[b - 1] > kk || (key[b - 1] == kk && val[b - 1] > v))) { key[b] = key[b - 1]; val[b] = val[b - 1]; --b; }
            key[b] = kk; val[b] = v;
        }
    }
    if (!__syncthreads_or(bad)) break;                        // (also orders the run fixing before the copy-out)
    first_pass = 0;
  }
    const uint64_t hi = (uint64_t)tile << 32;
    for (uint32_t i = tid; i < n; i += NT) {
        keys_out[rg.x + i] = hi | key[i];
        vals_out[rg.x + i] = val[i];
    }
  }
}

// ---- per-tile sort, segments larger than shared memory: in-place sorting network in global memory -----------
// Every comparator ascending (first step of each merge mirrors, the rest are half-cleaners), so a segment of any
// length n behaves as if padded with +inf: comparators touching an index >= n are no-ops and are skipped.
constexpr int TSORT_THREADS = 512;

__device__ __forceinline__ void tile_sort_network(uint64_t *buf, uint32_t n) {
    for (uint32_t k = 2; (k >> 1) < n; k <<= 1) {
        for (uint32_t p = threadIdx.x;; p += TSORT_THREADS) {
            const uint32_t blk = p / (k >> 1), r = p - blk * (k >> 1);
            const uint32_t i = blk * k + r, l = blk * k + k - 1 - r;
            if (i >= n) break;
            if (l < n) { const uint64_t a = buf[i], b = buf[l]; if (a > b) { buf[i] = b; buf[l] = a; } }
        }
        __syncthreads();
        for (uint32_t j = k >> 2; j > 0; j >>= 1) {
            for (uint32_t p = threadIdx.x;; p += TSORT_THREADS) {
                const uint32_t i = ((p & ~(j - 1)) << 1) | (p & (j - 1)), l = i + j;
                if (i >= n) break;
                if (l < n) { const uint64_t a = buf[i], b = buf[l]; if (a > b) { buf[i] = b; buf[l] = a; } }
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(TSORT_THREADS)
tile_sort_global_kernel(const int2 *__restrict__ ranges, uint64_t *__restrict__ entries, uint64_t *__restrict__ keys_out,
                        uint32_t *__restrict__ vals_out, const uint32_t *__restrict__ list, uint32_t *__restrict__ work) {
    __shared__ uint32_t s_item;
    const uint32_t count = work[2];
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_item = atomicAdd(&work[6], 1u);
        __syncthreads();
        if (s_item >= count) break;
        const uint32_t tile = list[s_item];
        const int2 rg = ranges[tile];
        const uint32_t n = (uint32_t)(rg.y - rg.x);
        uint64_t *seg = entries + rg.x;
        tile_sort_network(seg, n);
        const uint64_t hi = (uint64_t)tile << 32;
        for (uint32_t i = threadIdx.x; i < n; i += TSORT_THREADS) {
            const uint64_t e = seg[i];
            keys_out[rg.x + i] = hi | (e >> 32);
            vals_out[rg.x + i] = (uint32_t)e;
        }
    }
}

// ---- launch order of the blend CTAs: longest tiles first ------------------------------------------------------
// The blend kernels run one CTA per tile and a tile's cost is ~ its instance count; started in raster order the
// few very long tiles that happen to sit late in the grid finish alone at the end of the kernel.  A counting sort
// of the tiles on (count / 8, capped) in descending order is enough (one CTA, O(T)).
constexpr int ORDER_BUCKETS = 2048;
__global__ void __launch_bounds__(1024)
tile_order_kernel(int T, const int2 *__restrict__ ranges, uint32_t *__restrict__ order) {
    __shared__ uint32_t s_cnt[ORDER_BUCKETS];
    __shared__ uint32_t s_warp[32];
    for (int i = threadIdx.x; i < ORDER_BUCKETS; i += 1024) s_cnt[i] = 0;
    __syncthreads();
    auto bucket = [&](int t) {
        const int2 r = ranges[t];
        const int c = (r.y - r.x) >> 3;
        return ORDER_BUCKETS - 1 - min(c, ORDER_BUCKETS - 1);      // bucket 0 = longest
    };
    for (int t = threadIdx.x; t < T; t += 1024) atomicAdd(&s_cnt[bucket(t)], 1u);
    __syncthreads();
    // exclusive scan of the 2048 counters: two per thread
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    const uint32_t a = s_cnt[2 * threadIdx.x], b = s_cnt[2 * threadIdx.x + 1];
    uint32_t inc = a + b;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t n = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= (uint32_t)d) inc += n; }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = s_warp[lane], winc = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t n = __shfl_up_sync(0xffffffffu, winc, d); if (lane >= (uint32_t)d) winc += n; }
        s_warp[lane] = winc - w;
    }
    __syncthreads();
    const uint32_t excl = s_warp[warp] + inc - (a + b);
    s_cnt[2 * threadIdx.x] = excl;
    s_cnt[2 * threadIdx.x + 1] = excl + a;
    __syncthreads();
    for (int t = threadIdx.x; t < T; t += 1024) order[atomicAdd(&s_cnt[bucket(t)], 1u)] = (uint32_t)t;
}

// size classes of the shared-memory sort: 256 threads x 8 (2048 elements, 40 KB: five CTAs per SM) and
// 512 threads x 16 (8192 elements, 144 KB)
constexpr int TSORT_NT_S = 256, TSORT_ITEMS_S = 8, TSORT_NT_L = 512, TSORT_ITEMS_L = 16;
constexpr size_t tsort_smem(int nt, int items) { return (size_t)(4 * nt * items + (nt / 32) * 256) * sizeof(uint32_t); }

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

// Stable LSD radix sort of (64-bit key, 32-bit value) pairs on key bits [0, bits): 8-bit digits, ping-pong between the
// two key/value buffers of a binning workspace.  Returns the index of the buffer holding the result in *cur.
// Shared with the densification path (csrc/densify.cu), which sorts voxel coordinates with it.
int splatco::radix_sort_pairs(const BinWs &b, uint32_t n, int bits, int *cur_io, cudaStream_t st) {
    const int passes = (bits + 7) / 8;
    const int nblocks = (int)ceil_div64((int64_t)n, SORT_TILE);
    int cur = *cur_io;
    for (int p = 0; p < passes && n > 0; ++p) {
        const int shift = 8 * p;
        const int nb = bits - shift < 8 ? bits - shift : 8;
        const uint32_t mask = (1u << nb) - 1u;
        sort_hist_kernel<<<nblocks, SORT_THREADS, 0, st>>>(b.keys[cur], n, shift, mask, b.hist, nblocks);
        SPLATCO_CHECK_LAUNCH();
        sort_scan_kernel<<<256, 256, 0, st>>>(b.hist, nblocks, b.bin_totals);
        SPLATCO_CHECK_LAUNCH();
        sort_scatter_kernel<<<nblocks, SORT_THREADS, 0, st>>>(b.keys[cur], b.vals[cur], b.keys[cur ^ 1], b.vals[cur ^ 1],
                                                              n, shift, mask, b.hist, nblocks, b.bin_totals);
        SPLATCO_CHECK_LAUNCH();
        cur ^= 1;
    }
    *cur_io = cur;
    return 0;
}

extern "C" int splatco_sort_pairs(int64_t R, int H, int W, void *binning, void *stream) {
    if (check_R(R)) return -1;
    if (R == 0) return 0;
    SPLATCO_REQUIRE(binning, "sort_pairs: null pointer");
    BinWs b = bin_view(binning, R);
    const int T = ceil_div(W, TILE) * ceil_div(H, TILE);
    int cur = 0;
    return radix_sort_pairs(b, (uint32_t)R, 32 + tile_bits(T), &cur, (cudaStream_t)stream);
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

// Legacy composition of the three upstream stages (global LSD radix sort); kept for stage-level parity checks.
extern "C" int splatco_binning_radix(int P, int64_t R, int H, int W, const int32_t *radii, const void *geom,
                                     void *binning, void *image, void *stream) {
    int rc = splatco_duplicate_with_keys(P, R, H, W, radii, geom, binning, stream);
    if (rc) return rc;
    rc = splatco_sort_pairs(R, H, W, binning, stream);
    if (rc) return rc;
    rc = splatco_identify_tile_ranges(R, H, W, binning, image, stream);
    if (rc) return rc;
    ImgWs im = img_view(image, H, W);
    tile_order_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(ceil_div(W, TILE) * ceil_div(H, TILE), im.ranges, im.order);
    SPLATCO_CHECK_LAUNCH();
    return 0;
}

// 1 when splatco_binning(P, R, H, W) takes the tile-segmented path, whose R may be a CAPACITY (every write is clamped
// to it); 0 when it falls back to the radix composition, which needs the exact instance count.
// Gaussians per CTA of tile_hist / tile_scatter: the smallest power-of-two multiple of TCHUNK whose [chunks][tiles]
// histogram matrix fits the 4 B*R value buffer that holds it (4K images have 32 k tiles: 4096-Gaussian chunks of a
// 15 M-Gaussian view would need 118 M counters for 54 M instances; 16384-Gaussian chunks need 30 M).  0: no fit.
static int tile_chunk(int P, int64_t R, int H, int W) {
    const int T = ceil_div(W, TILE) * ceil_div(H, TILE);
    const size_t hist_bytes = (size_t)T * sizeof(uint32_t);
    if (hist_bytes > 200 * 1024) return 0;
    for (int chunk = TCHUNK; chunk <= TCHUNK_MAX; chunk *= 2)
        if ((size_t)ceil_div(P, chunk) * hist_bytes <= (size_t)R * sizeof(uint32_t)) return chunk;
    return 0;
}
static bool tile_path(int P, int64_t R, int H, int W) { return tile_chunk(P, R, H, W) != 0; }
extern "C" int splatco_binning_accepts_capacity(int P, int64_t R, int H, int W) {
    return (P > 0 && R > 0 && H > 0 && W > 0 && tile_path(P, R, H, W)) ? 1 : 0;
}

extern "C" int splatco_binning(int P, int64_t R, int H, int W, const int32_t *radii, const void *geom,
                               void *binning, void *image, void *stream) {
    if (check_R(R)) return -1;
    SPLATCO_REQUIRE(image && H > 0 && W > 0 && P >= 0, "binning: bad arguments");
    ImgWs im = img_view(image, H, W);
    cudaStream_t st = (cudaStream_t)stream;
    const int gx = ceil_div(W, TILE), gy = ceil_div(H, TILE);
    const int T = gx * gy;
    if (P == 0 || R == 0) {
        SPLATCO_CHECK_CUDA(cudaMemsetAsync(im.ranges, 0, sizeof(int2) * (size_t)T, st));
        return 0;
    }
    SPLATCO_REQUIRE(radii && geom && binning, "binning: null pointer");
    const int chunk = tile_chunk(P, R, H, W);
    const int nchunks = chunk ? ceil_div(P, chunk) : 0;
    const size_t hist_bytes = (size_t)T * sizeof(uint32_t);
    BinWs b = bin_view(binning, R);
    const int s = splatco_sorted_buffer_index(H, W);
    // the per-chunk histogram rows live in the value buffer that is not the sorted output (4 B*R): fall back to the
    // radix composition when they do not fit there or the tile histogram does not fit in shared memory
    // (that path treats R as the EXACT instance count: callers that pass a capacity must ask
    // splatco_binning_accepts_capacity first -- diff_gaussian_rasterization does)
    if (!chunk)
        return splatco_binning_radix(P, R, H, W, radii, geom, binning, image, stream);
    GeomWs g = geom_view(const_cast<void *>(geom), P);
    uint32_t *chunk_hist = b.vals[s ^ 1];
    static unsigned char attr_dev[64];       // cudaFuncSetAttribute is per device
    const int attr_i = current_device() & 63;
    if (!attr_dev[attr_i]) {
        SPLATCO_CHECK_CUDA(cudaFuncSetAttribute(tile_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        SPLATCO_CHECK_CUDA(cudaFuncSetAttribute(tile_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        SPLATCO_CHECK_CUDA(cudaFuncSetAttribute(tile_sort_radix_kernel<TSORT_NT_S, TSORT_ITEMS_S>,
                                                cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)tsort_smem(TSORT_NT_S, TSORT_ITEMS_S)));
        SPLATCO_CHECK_CUDA(cudaFuncSetAttribute(tile_sort_radix_kernel<TSORT_NT_L, TSORT_ITEMS_L>,
                                                cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)tsort_smem(TSORT_NT_L, TSORT_ITEMS_L)));
        attr_dev[attr_i] = 1;
    }
    SPLATCO_CHECK_CUDA(cudaMemsetAsync(im.tile_count, 0, (size_t)T * sizeof(uint32_t), st));
    tile_hist_kernel<<<nchunks, TCHUNK_THREADS, hist_bytes, st>>>(P, T, chunk, radii, g.rec, gx, gy, chunk_hist, im.tile_count, im.work);
    SPLATCO_CHECK_LAUNCH();
    tile_scan_kernel<<<1, 1024, 0, st>>>(T, im.tile_count, im.ranges, im.cursor, (uint32_t)R, im.work, im.lists,
                                         (uint32_t)(TSORT_NT_S * TSORT_ITEMS_S), (uint32_t)(TSORT_NT_L * TSORT_ITEMS_L));
    SPLATCO_CHECK_LAUNCH();
    tile_scatter_kernel<<<nchunks, TCHUNK_THREADS, hist_bytes, st>>>(P, T, chunk, radii, g.rec, gx, gy, chunk_hist, im.cursor,
                                                                    b.keys[s ^ 1], (uint32_t)R);
    SPLATCO_CHECK_LAUNCH();
    // persistent sort CTAs pull tiles of their size class from the lists tile_scan built
    tile_sort_radix_kernel<TSORT_NT_S, TSORT_ITEMS_S><<<min(T, 148 * 5), TSORT_NT_S, tsort_smem(TSORT_NT_S, TSORT_ITEMS_S), st>>>(
        im.ranges, b.keys[s ^ 1], b.keys[s], b.vals[s], im.lists, im.work, 0);
    SPLATCO_CHECK_LAUNCH();
    tile_sort_radix_kernel<TSORT_NT_L, TSORT_ITEMS_L><<<min(T, 148), TSORT_NT_L, tsort_smem(TSORT_NT_L, TSORT_ITEMS_L), st>>>(
        im.ranges, b.keys[s ^ 1], b.keys[s], b.vals[s], im.lists + T, im.work, 1);
    SPLATCO_CHECK_LAUNCH();
    tile_sort_global_kernel<<<min(T, 148), TSORT_THREADS, 0, st>>>(im.ranges, b.keys[s ^ 1], b.keys[s], b.vals[s],
                                                                   im.lists + 2 * (size_t)T, im.work);
    SPLATCO_CHECK_LAUNCH();
    tile_order_kernel<<<1, 1024, 0, st>>>(T, im.ranges, im.order);
    SPLATCO_CHECK_LAUNCH();
    return 0;
}
