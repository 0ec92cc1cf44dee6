// Shared helpers for libsplatco_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/splatco_b200.h"

namespace splatco {

constexpr int TILE = SPLATCO_TILE;           // 16x16 pixel tiles
constexpr int PRE_THREADS = 256;             // per-Gaussian kernels: 1 Gaussian / thread
constexpr int SORT_THREADS = 256;
constexpr int SORT_ITEMS = 8;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;   // 2048 pairs per CTA per pass
constexpr size_t ALIGN = 256;

inline size_t align_up(size_t x) { return (x + ALIGN - 1) / ALIGN * ALIGN; }
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int current_device() { int d = 0; cudaGetDevice(&d); return d; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

void set_error(const char *fmt, ...);
void count_launch(int n = 1);          // process-wide count of kernels this library launched

#define SPLATCO_CHECK_CUDA(expr)                                                           \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            splatco::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,        \
                               cudaGetErrorString(_e));                                    \
            return -2;                                                                     \
        }                                                                                  \
    } while (0)

#define SPLATCO_CHECK_LAUNCH()                                                              \
    do { splatco::count_launch(); SPLATCO_CHECK_CUDA(cudaGetLastError()); } while (0)

#define SPLATCO_REQUIRE(cond, ...)                                                         \
    do {                                                                                   \
        if (!(cond)) { splatco::set_error(__VA_ARGS__); return -1; }                       \
    } while (0)

// ---- workspace layouts (mirrored by splatco_*_layout) ------------------------------------------
struct GeomWs {
    float4 *rec;              // [3P]  (x,y,conA,conB | conC,opacity,r,g | b,depth,radius,0)
    float *depths;            // [P]
    uint32_t *tiles;          // [P]
    uint32_t *block_sums;     // [nb]
    uint32_t *block_offsets;  // [nb]
    uint32_t *total;          // [1]
};
struct BinWs {
    uint64_t *keys[2];
    uint32_t *vals[2];
    uint32_t *hist;           // [256 * nsb]
    uint32_t *bin_totals;     // [256]
};
struct ImgWs {
    int2 *ranges;             // [T]
    float *final_T;           // [HW]
    int32_t *n_contrib;       // [HW]
    uint32_t *tile_count;     // [T]  instances per tile (tile-segmented binning)
    uint32_t *cursor;         // [T]  start of each tile's segment
    uint32_t *lists;          // [3][T] tiles by segment-length class (per-tile sort work lists)
    uint32_t *work;           // [8]  list lengths [0..2], queue heads [4..6]
    uint32_t *order;          // [T]  tile handled by blend CTA b (longest tiles first)
};

size_t geom_offsets(int P, size_t off[7]);
size_t bin_offsets(int64_t R, size_t off[7]);
size_t img_offsets(int H, int W, size_t off[9]);
GeomWs geom_view(void *base, int P);
BinWs bin_view(void *base, int64_t R);
ImgWs img_view(void *base, int H, int W);
int radix_sort_pairs(const BinWs &b, uint32_t n, int bits, int *cur_io, cudaStream_t st);   // csrc/binning.cu

inline int tile_bits(int T) { int b = 1; while ((1 << b) < T) ++b; return b; }   // bits to hold T-1
inline int sort_passes(int H, int W) {
    const int T = ceil_div(W, TILE) * ceil_div(H, TILE);
    return (32 + tile_bits(T) + 7) / 8;
}

// ---- device helpers ------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
    float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
}

// ---- exclusive scan of the block sums (single CTA; nb = ceil(P/256) <= ~80k for 20 M Gaussians) --
static __global__ void __launch_bounds__(1024)
scan_block_sums_kernel(int nb, const uint32_t *__restrict__ block_sums,
                       uint32_t *__restrict__ block_offsets, uint32_t *__restrict__ total) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    for (int base = 0; base < nb; base += 1024) {
        const int i = base + threadIdx.x;
        const uint32_t v = i < nb ? block_sums[i] : 0u;
        uint32_t inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { uint32_t n = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= (uint32_t)d) inc += n; }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = s_warp[lane], winc = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { uint32_t n = __shfl_up_sync(0xffffffffu, winc, d); if (lane >= (uint32_t)d) winc += n; }
            s_warp[lane] = winc - w;   // exclusive warp offsets
        }
        __syncthreads();
        const uint32_t carry = s_carry;
        const uint32_t excl = carry + s_warp[warp] + inc - v;
        if (i < nb) block_offsets[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = s_carry;
}


}  // namespace splatco
