// Anchor decode v2 (included by decode.cu after the v1 kernels, whose fold / scan / bilinear helpers it reuses).
//
// The reference's decode (gaussian_renderer/__init__.py:18-116, scene/gaussian_model.py:149-169,316-337) has NO
// nonlinearity between the geo features and the heads' first Linear:
//     geo = [ Lin(BN(P)) | CLin(CBN(g)) ]            (64)        BatchNorm in train mode = an affine map per view
//     h   = relu( [feat | dir,dist | geo] W1 + b1 )  (96)
// so with BN folded into the Linears (dec_fold_kernel) the whole path from the gathered row to the hidden layer is ONE
// affine map, evaluated here as one GEMM on the row  u = [ g (71) | 1 | dir,dist (4) | P (DP) ]:
//     h = relu( u Wc1 ),   Wc1 = [Wc' W1g_c + W1f ; b1 + bgeo W1g ; W1d ; Wp' W1g_p]        (dec2_combine_kernel)
//     z = h W2 (block diagonal: three K = 32 products) + b2 -> tanh | raw | sigmoid
// Two tcgen05 stages instead of three, no geo / x100 tensors, and in the backward ONE weight-gradient product
// G = u^T dH from which every BN / Linear / W1 gradient and both BatchNorm-backward sums follow (dec2_expand_kernel):
// the "1" column carries the bias forward and the column sums of dH backward.
//
// Tensors between kernels are stored in the tensor-core TILE layout [tile][chunk][128 rows][4 floats] (tc.cuh), written
// and read with 16-byte accesses that are contiguous across the 32 rows of a warp:
//   XT  u rows            19 + ceil(DP/4) chunks     HT  hidden, 24 chunks
//   ZT  head outputs after activation, 32 chunks in BLOCK order [opacity 16 | cov 80 | colour 32]
//   DUT dL/du (+ the direct anchor / offset / scaling gradients of the post-processing), same chunks as XT
#pragma once
#include "tc.cuh"

namespace splatco {

constexpr int D2_ROWS = 128;
constexpr int D2_WORKERS = 512;                      // 16 worker warps: row = t & 127 (TMEM lane), column group = t >> 7
constexpr int D2_THREADS = D2_WORKERS + 32;          // + one control warp (TMA + MMA issue)
constexpr int D2_DIR_CH = 18;                        // chunk of (dir.xyz, dist)
constexpr int D2_P_CH0 = 19;                         // first plane chunk
constexpr int D2_ONE = GD;                           // u column 71: constant 1
constexpr int D2_UDIR = 72, D2_UP0 = 76;             // u columns of dir/dist and of plane column 0
constexpr int D2_MAX_NK = 36;                        // chunks of the stage-1 operand (even)
constexpr uint32_t D2_CHUNK = 2048;
constexpr int D2_ZCH = 32;                           // ZT chunks: block columns [0,16) opacity, [16,96) cov, [96,128) colour
constexpr int D2_ZCOV = 16, D2_ZCOL = 96;
constexpr int D2_RCOV = 16, D2_RCOL = 88, D2_RCH = 30;   // dZ operand columns (backward): [opacity 16 | cov 72 | colour 32]

// phase trace (measurement aid, splatco_decode_profile(2)): CTA 0's worker thread 0 stamps clock64 at the phase boundaries
// of its first tiles; [0] = forward kernel, [1] = backward kernel
__device__ unsigned long long g_d2_trace[2][64];
#define D2_TRACE(which, slot)                                                                         \
    do { if (a.trace && blockIdx.x == 0 && tid == 0 && (slot) < 64) g_d2_trace[which][slot] = clock64(); } while (0)

struct D2Dims { int V, rc, level, DP, LDX, npc, nch, nk, NB, ntiles; };
inline D2Dims d2_dims(int V, int rc, int level) {
    D2Dims d;
    d.V = V; d.rc = rc; d.level = level;
    d.DP = rc * (level == 0 ? 6 : (level == 1 ? 9 : 12));
    d.LDX = ru4(d.DP + GD);
    d.npc = (d.DP + 3) >> 2;
    d.nch = D2_P_CH0 + d.npc;
    d.nk = (d.nch + 1) & ~1;
    d.NB = (4 * d.nk + 15) & ~15;
    d.ntiles = (V + D2_ROWS - 1) / D2_ROWS;
    return d;
}

// The stage-1 weights travel through the ring as three K slices (all 96 outputs each): chunks [c0, c0 + nc) of slice s
__host__ __device__ inline void d2_kslice(int nk, int s, int &c0, int &nc) {
    const int ks = nk / 2, q = ks / 3, r = ks % 3;
    c0 = 2 * (s * q + (s < r ? s : r));
    nc = 2 * (q + (s < r ? 1 : 0));
}

// row count known on the host (p == nullptr) or left on the device by the prefilter (splatco_decode_desc::V_dev)
__device__ __forceinline__ int d2_count(const int32_t *p, int v) { return p ? __ldg(p) : v; }

__device__ __forceinline__ int d2_zcol_op(int k) { return k; }
__device__ __forceinline__ int d2_zcol_cov(int k, int q) { return D2_ZCOV + 7 * k + q; }
__device__ __forceinline__ int d2_zcol_col(int k, int q) { return D2_ZCOL + 3 * k + q; }
// float index of column c of row `row` inside a tile (chunk stride = 128 cells x 4 floats)
__device__ __forceinline__ int d2_tile_idx(int c) { return (c >> 2) * 512 + (c & 3); }

__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add_v2(float *addr, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}

// Column sums of 8 per-lane values over the 32 lanes of a warp (transposing butterfly: 9 shuffles for 8 columns).
// On return the lanes with (lane & 3) == 0 hold the sum of column `idx`.
__device__ __forceinline__ float warp_colsum8(const float (&g)[8], int lane, int &idx) {
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
    float w4[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) w4[i] = (b4 ? g[i + 4] : g[i]) + __shfl_xor_sync(0xffffffffu, b4 ? g[i] : g[i + 4], 16);
    float w2[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) w2[i] = (b3 ? w4[i + 2] : w4[i]) + __shfl_xor_sync(0xffffffffu, b3 ? w4[i] : w4[i + 2], 8);
    float r = (b2 ? w2[1] : w2[0]) + __shfl_xor_sync(0xffffffffu, b2 ? w2[0] : w2[1], 4);
    r += __shfl_xor_sync(0xffffffffu, r, 2);
    r += __shfl_xor_sync(0xffffffffu, r, 1);
    idx = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
    return r;
}

// =======================================================================================================
// Gather: one thread per visible anchor, one CTA per 128-row tile.  Every XT store is a 16-byte cell and the 32 lanes
// of a warp write 512 contiguous bytes; texels of channel-last planes are fetched as one 32-byte sector.  BatchNorm
// batch statistics: transposing warp reductions -> per-CTA sums -> one fp64 atomic per column and CTA.
// =======================================================================================================
template <int LEVEL, int RC, bool PACKED>
__global__ void __launch_bounds__(D2_ROWS)
dec2_gather_kernel(DecPtrs p, int V_host, const int32_t *__restrict__ Vdev, int LDX, float4 *__restrict__ XT,
                   double *__restrict__ stats) {
    constexpr int NS = LEVEL == 0 ? 6 : (LEVEL == 1 ? 9 : 12);
    constexpr int DP = NS * RC, NPC = (DP + 3) / 4, NCH = D2_P_CH0 + NPC;
    __shared__ float s_part[4][2][DEC_MAX_DP + GD + 9];
    const int V = d2_count(Vdev, V_host);
    if ((int)blockIdx.x * D2_ROWS >= V) return;          // (grid sized for an upper bound of V)
    const int r = threadIdx.x, lane = r & 31, warp = r >> 5;
    const int v = blockIdx.x * D2_ROWS + r;
    const bool valid = v < V;
    float4 *xt = XT + (size_t)blockIdx.x * NCH * D2_ROWS + r;
    auto stat8 = [&](const float (&x)[8], int col0, int nvalid) {
        float q[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) q[i] = x[i] * x[i];
        int idx;
        const float s = warp_colsum8(x, lane, idx);
        const float s2 = warp_colsum8(q, lane, idx);
        if ((lane & 3) == 0 && idx < nvalid) { s_part[warp][0][col0 + idx] = s; s_part[warp][1][col0 + idx] = s2; }
    };
    const int i = valid ? p.vis[v] : 0;
    float ax = 0.f, ay = 0.f, az = 0.f;
    // ---- context block g = [feat 32 | anchor 3 | offsets 30 | scaling 6], then the constant 1 ------------------------
    {
        const float4 *fp = reinterpret_cast<const float4 *>(p.anchor_feat + (size_t)i * FD);
        float4 f[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) f[q] = valid ? __ldg(fp + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        float ga[40];
        if (valid) {
            ax = __ldg(p.anchor + 3 * (size_t)i); ay = __ldg(p.anchor + 3 * (size_t)i + 1); az = __ldg(p.anchor + 3 * (size_t)i + 2);
            ga[0] = ax; ga[1] = ay; ga[2] = az;
            const float2 *op = reinterpret_cast<const float2 *>(p.offset + (size_t)i * 3 * KO);
#pragma unroll
            for (int q = 0; q < 15; ++q) { const float2 t = __ldg(op + q); ga[3 + 2 * q] = t.x; ga[4 + 2 * q] = t.y; }
            const float2 *sp = reinterpret_cast<const float2 *>(p.scaling + (size_t)i * 6);
#pragma unroll
            for (int q = 0; q < 3; ++q) { const float2 t = __ldg(sp + q); ga[33 + 2 * q] = t.x; ga[34 + 2 * q] = t.y; }
            ga[39] = 1.f;
        } else {
#pragma unroll
            for (int q = 0; q < 40; ++q) ga[q] = 0.f;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) xt[q * D2_ROWS] = f[q];
#pragma unroll
        for (int q = 0; q < 10; ++q) xt[(8 + q) * D2_ROWS] = make_float4(ga[4 * q], ga[4 * q + 1], ga[4 * q + 2], ga[4 * q + 3]);
        // (dir, dist)   gaussian_renderer/__init__.py:34-38
        float4 dd = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) {
            const float vx = ax - __ldg(p.cam), vy = ay - __ldg(p.cam + 1), vz = az - __ldg(p.cam + 2);
            const float dist = sqrtf(vx * vx + vy * vy + vz * vz);
            dd = make_float4(vx / dist, vy / dist, vz / dist, dist);
        }
        xt[D2_DIR_CH * D2_ROWS] = dd;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float x8[8] = {f[2 * q].x, f[2 * q].y, f[2 * q].z, f[2 * q].w, f[2 * q + 1].x, f[2 * q + 1].y, f[2 * q + 1].z, f[2 * q + 1].w};
            stat8(x8, DP + 8 * q, 8);
        }
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            const float x8[8] = {ga[8 * q], ga[8 * q + 1], ga[8 * q + 2], ga[8 * q + 3], ga[8 * q + 4], ga[8 * q + 5], ga[8 * q + 6], ga[8 * q + 7]};
            stat8(x8, DP + FD + 8 * q, q == 4 ? 7 : 8);      // the last value of the last group is the constant
        }
    }
    // ---- plane block: NS bilinear samples of RC channels --------------------------------------------------------------
    float pv[NPC * 4];
#pragma unroll
    for (int q = 0; q < NPC * 4; ++q) pv[q] = 0.f;
    if (valid) {
        float ind[3];
        norm_coords(p, ax, ay, az, ind);
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const int lvl = s < 6 ? 0 : (s < 9 ? 1 : 2);
            const int pl = s < 6 ? (s >> 1) : (s < 9 ? s - 6 : s - 9);
            const bool att = s < 6 && (s & 1);
            const float *base = att ? p.att[pl] : p.plane[lvl][pl];
            const int E = p.E[lvl];
            float u, w;
            plane_axes(pl, ind, u, w);
            const Bilin b = bilin_setup(u, w, E);
            const int ti[4] = {b.i00, b.i01, b.i10, b.i11};
            const float tw[4] = {b.w00, b.w01, b.w10, b.w11};
            float acc[RC];
#pragma unroll
            for (int ch = 0; ch < RC; ++ch) acc[ch] = 0.f;
            if (PACKED) {
                float4 t4[4];
                float t1[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const bool in = ti[t] >= 0;
                    const float *tp = base + (size_t)(in ? ti[t] : 0) * 8;
                    t4[t] = in ? __ldg(reinterpret_cast<const float4 *>(tp)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    t1[t] = (RC > 4 && in) ? __ldg(tp + 4) : 0.f;
                }
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const float tv[5] = {t4[t].x, t4[t].y, t4[t].z, t4[t].w, t1[t]};
#pragma unroll
                    for (int ch = 0; ch < RC; ++ch) acc[ch] = fmaf(tv[ch], tw[t], acc[ch]);
                }
            } else {
                const size_t cs = (size_t)E * E;
                float tv[4][RC];
#pragma unroll
                for (int t = 0; t < 4; ++t)
#pragma unroll
                    for (int ch = 0; ch < RC; ++ch) tv[t][ch] = ti[t] >= 0 ? __ldg(base + ch * cs + ti[t]) : 0.f;
#pragma unroll
                for (int t = 0; t < 4; ++t)
#pragma unroll
                    for (int ch = 0; ch < RC; ++ch) acc[ch] = fmaf(tv[t][ch], tw[t], acc[ch]);
            }
#pragma unroll
            for (int ch = 0; ch < RC; ++ch) {
                const int c = s * RC + ch;
                float val = acc[ch];
                if (s >= 6) {       // training-time plane-feature noise of the non-attended levels (scene/grids.py:159-164)
                    if (p.noise) val += __ldg(p.noise + (size_t)v * (DP - 6 * RC) + (c - 6 * RC));
                    else if (p.noise_q != 0.f) val = fmaf(uniform_pm_half(p.noise_seed, (unsigned long long)v * DP + c), p.noise_q, val);
                }
                pv[c] = val;
            }
        }
    }
#pragma unroll
    for (int q = 0; q < NPC; ++q) xt[(D2_P_CH0 + q) * D2_ROWS] = make_float4(pv[4 * q], pv[4 * q + 1], pv[4 * q + 2], pv[4 * q + 3]);
#pragma unroll
    for (int q = 0; q < (DP + 7) / 8; ++q) {
        float x8[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) x8[e] = 8 * q + e < NPC * 4 ? pv[8 * q + e] : 0.f;
        stat8(x8, 8 * q, DP - 8 * q < 8 ? DP - 8 * q : 8);
    }
    __syncthreads();
    for (int c = r; c < DP + GD; c += D2_ROWS) {
        const double s = (double)s_part[0][0][c] + (double)s_part[1][0][c] + (double)s_part[2][0][c] + (double)s_part[3][0][c];
        const double q = (double)s_part[0][1][c] + (double)s_part[1][1][c] + (double)s_part[2][1][c] + (double)s_part[3][1][c];
        atomicAdd(&stats[c], s);
        atomicAdd(&stats[LDX + c], q);
    }
}

// =======================================================================================================
// Combined weights (after dec_fold_kernel): Wc1 = the affine map u -> hidden pre-activation, packed hi/lo into
//   W1S  forward B tiles, three K slices (d2_kslice) of all 96 outputs: [slice][hi | lo][chunks of the slice][96 rows]
//   W1R  backward B tile (dU = dH Wc1^T): [hi | lo][24 chunks][NB rows]   (row 71, the bias row, is zero)
//   W2B  forward block tiles [hi | lo]{[8][16] opacity, [8][80] cov, [8][32] colour};  b2blk[128]
//   W2R  backward block tiles (dH_h = dZ_h W2_h^T): [hi | lo]{[4][32], [18][32], [8][32]}
// =======================================================================================================
constexpr uint32_t D2_W2B_HALF = 8 * (16 + 80 + 32) * 16;           // 16384
constexpr uint32_t D2_W2R_HALF = (4 + 18 + 8) * 32 * 16;            // 15360
constexpr uint32_t D2_RING_SLOT = 12 * HD * 16 * 2;                 // 36864: the largest K slice (6 K steps = 12 chunks x 96 rows, hi + lo)
static_assert(2 * D2_W2B_HALF <= D2_RING_SLOT, "W2 blocks travel through a ring slot");

__device__ __forceinline__ void d2_put(uint8_t *dst, uint32_t half, size_t cell, int sub, float w) {
    const float h = tc::tf32_hi(w);
    reinterpret_cast<float *>(dst + cell * 16)[sub] = h;
    reinterpret_cast<float *>(dst + half + cell * 16)[sub] = w - h;
}

__global__ void __launch_bounds__(256)
dec2_combine_kernel(int DP, int nk, int NB, const float *__restrict__ WpT, const float *__restrict__ WcT,
                    const float *__restrict__ bgeo, const float *__restrict__ W1T, const float *__restrict__ b1e,
                    const float *__restrict__ W2T, const float *__restrict__ b2, uint8_t *__restrict__ W1S,
                    uint8_t *__restrict__ W1R, uint8_t *__restrict__ W2B, float *__restrict__ b2blk,
                    uint8_t *__restrict__ W2R) {
    const int tid = blockIdx.x * 256 + threadIdx.x, nthr = gridDim.x * 256;
    const uint32_t r_half = 24u * NB * 16;
    for (int e = tid; e < NB * HD; e += nthr) {
        const int uc = e / HD, n = e - uc * HD;
        float w = 0.f;
        if (uc < GD) {
            for (int o = 0; o < 32; ++o) w = fmaf(WcT[uc * 32 + o], W1T[(68 + o) * HD + n], w);
            if (uc < FD) w += W1T[uc * HD + n];
        } else if (uc == D2_ONE) {
            for (int o = 0; o < 64; ++o) w = fmaf(bgeo[o], W1T[(36 + o) * HD + n], w);
            w += b1e[n];
        } else if (uc < D2_UP0) {
            w = W1T[(FD + uc - D2_UDIR) * HD + n];
        } else if (uc < D2_UP0 + DP) {
            const int c = uc - D2_UP0;
            for (int o = 0; o < 32; ++o) w = fmaf(WpT[c * 32 + o], W1T[(36 + o) * HD + n], w);
        }
        if (uc < 4 * nk) {
            int sl = 0, c0, nc;
            d2_kslice(nk, 1, c0, nc);
            if ((uc >> 2) >= c0) sl = 1;
            d2_kslice(nk, 2, c0, nc);
            if ((uc >> 2) >= c0) sl = 2;
            d2_kslice(nk, sl, c0, nc);
            d2_put(W1S + (size_t)c0 * HD * 32, (uint32_t)nc * HD * 16, (size_t)((uc >> 2) - c0) * HD + n, uc & 3, w);
        }
        d2_put(W1R, r_half, (size_t)(n >> 2) * NB + uc, n & 3, uc == D2_ONE ? 0.f : w);
    }
    // W2 blocks.  compact output index j: [0,10) opacity, [10,80) cov, [80,110) colour
    for (int e = tid; e < 128 * 32; e += nthr) {
        const int jb = e >> 5, ii = e & 31;                  // block column, hidden index within the head
        const int h = jb < D2_ZCOV ? 0 : (jb < D2_ZCOL ? 1 : 2);
        const int jj = jb - (h == 0 ? 0 : (h == 1 ? D2_ZCOV : D2_ZCOL));
        const int nout = h == 0 ? KO : (h == 1 ? 7 * KO : 3 * KO), j0 = h == 0 ? 0 : (h == 1 ? KO : 8 * KO);
        const int rows = h == 0 ? 16 : (h == 1 ? 80 : 32);
        const float w = jj < nout ? W2T[(32 * h + ii) * ZD + j0 + jj] : 0.f;
        const size_t fbase = h == 0 ? 0 : (h == 1 ? 8 * 16 : 8 * (16 + 80));
        d2_put(W2B, D2_W2B_HALF, fbase + (size_t)(ii >> 2) * rows + jj, ii & 3, w);
        // backward: rows = hidden ii, K = jj (chunks of 4 outputs); cov has 18 chunks (72 columns), jj < 80 only 72 used
        const int kc = h == 0 ? 4 : (h == 1 ? 18 : 8);
        if (jj < 4 * kc) {
            const size_t rbase = h == 0 ? 0 : (h == 1 ? 4 * 32 : (4 + 18) * 32);
            d2_put(W2R, D2_W2R_HALF, rbase + (size_t)(jj >> 2) * 32 + ii, jj & 3, w);
        }
        if (ii == 0) b2blk[jb] = jj < nout ? b2[j0 + jj] : 0.f;
    }
}

// =======================================================================================================
// Forward MLP: persistent, one CTA per SM, 128 anchors per tile.
//   workers (16 warps): split the TMA-loaded u tile hi/lo in place, epilogues;  control warp: bulk copies + tcgen05.mma.
//   stage 1: H[128,96] = u Wc1 in three K slices whose weights stream through a 2-slot ring (the u operand and the full
//            weight set do not fit shared memory together; a slice is fetched while the previous one multiplies).  K
//            slices, not N slices: every tcgen05.mma reads its 128 x 8 A block (4 KB) from shared memory whatever its N,
//            and at N = 32 that read -- not the tensor pipe -- set the pace (phase trace: 51 MMAs = 1.4 us)
//   stage 2: Z = H W2, three block products (K = 32 each), weights = the ring's fourth item
//   the next tile's u rows are fetched during epilogue 2, the next tile's first two slices during epilogues 1 and 2.
// Shared memory: U hi [0, 73728) | U lo [73728, 147456) | ring 2 x 36864;  H hi / lo reuse the U regions.
// TMEM: H accumulator columns [0,96), Z accumulator [96,224).
// =======================================================================================================
constexpr uint32_t D2F_ULO = D2_MAX_NK * D2_CHUNK;                  // 73728
constexpr uint32_t D2F_RING = 2 * D2F_ULO;                          // 147456
constexpr uint32_t D2F_SMEM = D2F_RING + 2 * D2_RING_SLOT;          // 221184

struct D2Fwd {
    int V, nch, nk, trace;
    const int32_t *Vdev;
    const float4 *XT;
    const uint8_t *W1S, *W2B;
    const float *b2blk;
    float4 *HT, *ZT;
    float *nopac;
    uint8_t *mask_out;
    uint32_t *maskbits, *block_sums;
};

__device__ __forceinline__ void d2_bar_sync_all() { asm volatile("bar.sync 0, %0;" ::"n"(D2_THREADS) : "memory"); }

__global__ void __launch_bounds__(D2_THREADS, 1)
dec2_mlp_fwd_kernel(D2Fwd a) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t barU, barM1, barM2, full[2], empty[2];
    __shared__ uint32_t tmem_s;
    __shared__ float s_b2[128];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tc::tmem_alloc<256>(&tmem_s);
    if (tid == 0) {
        tc::mbar_init(&barU, 1); tc::mbar_init(&barM1, 1); tc::mbar_init(&barM2, 1);
        tc::mbar_init(&full[0], 1); tc::mbar_init(&full[1], 1); tc::mbar_init(&empty[0], 1); tc::mbar_init(&empty[1], 1);
        tc::fence_barrier_init();
    }
    if (tid < 128) s_b2[tid] = a.b2blk[tid];
    // chunks [nch, nk) of the operand are never written by the copies: zero them once (both halves)
    for (int e = tid; e < (a.nk - a.nch) * D2_ROWS; e += D2_THREADS) {
        *reinterpret_cast<float4 *>(sm + (size_t)(a.nch * D2_ROWS + e) * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
        *reinterpret_cast<float4 *>(sm + D2F_ULO + (size_t)(a.nch * D2_ROWS + e) * 16) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_s;
    const uint32_t u_hi = tc::smem_u32(sm), u_lo = u_hi + D2F_ULO, ring = u_hi + D2F_RING;
    const uint32_t tile_bytes = (uint32_t)a.nch * D2_CHUNK;
    int sl_c0[3], sl_nc[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) d2_kslice(a.nk, q, sl_c0[q], sl_nc[q]);
    auto slice_src = [&](int q) { return a.W1S + (size_t)sl_c0[q] * HD * 32; };
    auto slice_bytes = [&](int q) { return (uint32_t)sl_nc[q] * HD * 32; };
    const int Vn = d2_count(a.Vdev, a.V), ntiles = (Vn + D2_ROWS - 1) / D2_ROWS;
    const int ntl = ntiles > (int)blockIdx.x ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;     // tiles of this CTA

    if (warp == D2_WORKERS / 32) {
        // =========================== control warp ===========================
        uint32_t pf[2] = {0, 0}, pe[2] = {0, 0};
        bool used[2] = {false, false};
        auto load_item = [&](int slot, const uint8_t *src, uint32_t bytes) {
            if (used[slot]) { tc::mbar_wait(&empty[slot], pe[slot]); pe[slot] ^= 1; }
            used[slot] = true;
            if (lane == 0) {
                tc::mbar_arrive_expect_tx(&full[slot], bytes);
                tc::bulk_g2s(sm + D2F_RING + slot * D2_RING_SLOT, src, bytes, &full[slot]);
            }
        };
        auto wait_item = [&](int slot) { tc::mbar_wait(&full[slot], pf[slot]); pf[slot] ^= 1; tc::tc_fence_after(); };
        constexpr uint32_t id32 = tc::make_idesc_tf32(128, 32), id16 = tc::make_idesc_tf32(128, 16), id80 = tc::make_idesc_tf32(128, 80),
                           id96 = tc::make_idesc_tf32(128, HD);
        if (ntl > 0) {
            if (lane == 0) {
                tc::mbar_arrive_expect_tx(&barU, tile_bytes);
                tc::bulk_g2s(sm, a.XT + (size_t)blockIdx.x * a.nch * D2_ROWS, tile_bytes, &barU);
            }
            load_item(0, slice_src(0), slice_bytes(0));
            load_item(1, slice_src(1), slice_bytes(1));
        }
        for (int it = 0; it < ntl; ++it) {
            const int tile = blockIdx.x + it * gridDim.x;
            d2_bar_sync_all();                                  // A: operand split done
            tc::tc_fence_after();
#define D2_CTRACE(slot) do { if (a.trace && blockIdx.x == 0 && lane == 0 && it < 4) g_d2_trace[0][32 + 8 * it + (slot)] = clock64(); } while (0)
            D2_CTRACE(0);
            for (int s = 0; s < 3; ++s) {
                const int slot = s & 1;
                wait_item(slot);
                D2_CTRACE(1 + 2 * s);
                if (lane == 0) {
                    const uint32_t b = ring + slot * D2_RING_SLOT;
                    tc::issue_3xtf32_lbo(tmem, u_hi, u_lo, D2_CHUNK, sl_c0[s], b, b + (uint32_t)sl_nc[s] * HD * 16, HD * 16, 0, sl_nc[s] / 2, id96, s > 0);
                    tc::mma_commit(&empty[slot]);
                    if (s == 2) tc::mma_commit(&barM1);
                }
                __syncwarp();
                D2_CTRACE(2 + 2 * s);
                // refill the slots behind the running products
                if (s == 1) load_item(0, slice_src(2), slice_bytes(2));
                if (s == 2) load_item(1, a.W2B, 2 * D2_W2B_HALF);
            }
            D2_CTRACE(7);
            d2_bar_sync_all();                                  // B: H operand written
            tc::tc_fence_after();
            wait_item(1);
            if (lane == 0) {
                const uint32_t b = ring + D2_RING_SLOT;
                tc::issue_3xtf32(tmem + 96, u_hi, u_lo, D2_ROWS, 0, b, b + D2_W2B_HALF, 16, 0, 4, id16, false);
                tc::issue_3xtf32(tmem + 96 + D2_ZCOV, u_hi, u_lo, D2_ROWS, 8, b + 8 * 16 * 16, b + D2_W2B_HALF + 8 * 16 * 16, 80, 0, 4, id80, false);
                tc::issue_3xtf32(tmem + 96 + D2_ZCOL, u_hi, u_lo, D2_ROWS, 16, b + 8 * 96 * 16, b + D2_W2B_HALF + 8 * 96 * 16, 32, 0, 4, id32, false);
                tc::mma_commit(&empty[1]);
                tc::mma_commit(&barM2);
            }
            __syncwarp();
            if (it + 1 < ntl) {
                load_item(0, slice_src(0), slice_bytes(0));     // slot 0: slice 2 has been consumed
                load_item(1, slice_src(1), slice_bytes(1));     // waits for stage 2 => the H operand is dead
                if (lane == 0) {
                    tc::mbar_arrive_expect_tx(&barU, tile_bytes);
                    tc::bulk_g2s(sm, a.XT + (size_t)(tile + gridDim.x) * a.nch * D2_ROWS, tile_bytes, &barU);
                }
            }
        }
    } else {
        // =========================== workers ===========================
        const int r = tid & (D2_ROWS - 1), grp = tid >> 7;
        const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        const int ncell = a.nch * D2_ROWS;
        for (int it = 0; it < ntl; ++it) {
            const int tile = blockIdx.x + it * gridDim.x;
            const int row = tile * D2_ROWS + r;
            const bool valid = row < Vn;
            const uint32_t par = it & 1;
            D2_TRACE(0, 8 * it + 0);
            tc::mbar_wait(&barU, par);
            D2_TRACE(0, 8 * it + 1);
            for (int e = tid; e < ncell; e += D2_WORKERS) {
                float4 *ph = reinterpret_cast<float4 *>(sm + (size_t)e * 16);
                const float4 x = *ph;
                const float4 h = make_float4(tc::tf32_hi(x.x), tc::tf32_hi(x.y), tc::tf32_hi(x.z), tc::tf32_hi(x.w));
                *ph = h;
                *reinterpret_cast<float4 *>(sm + D2F_ULO + (size_t)e * 16) = make_float4(x.x - h.x, x.y - h.y, x.z - h.z, x.w - h.w);
            }
            tc::fence_proxy_async();
            tc::tc_fence_before();
            D2_TRACE(0, 8 * it + 2);
            d2_bar_sync_all();                                  // A
            tc::mbar_wait(&barM1, par);
            tc::tc_fence_after();
            D2_TRACE(0, 8 * it + 3);
            // ---- epilogue 1: H = relu(acc) -> HT (for the backward) and the stage-2 operand --------------------------
            {
                float4 *ht = a.HT + ((size_t)tile * 24 + 6 * grp) * D2_ROWS + r;
#pragma unroll
                for (int n0 = 0; n0 < 24; n0 += 8) {
                    float v[8];
                    tc::tmem_ld8(tlane + 24 * grp + n0, v);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 8; ++q) v[q] = fmaxf(v[q], 0.f);
                    const int c = 6 * grp + (n0 >> 2);
                    ht[(n0 >> 2) * D2_ROWS] = make_float4(v[0], v[1], v[2], v[3]);
                    ht[((n0 >> 2) + 1) * D2_ROWS] = make_float4(v[4], v[5], v[6], v[7]);
#pragma unroll
                    for (int hq = 0; hq < 2; ++hq) {
                        const float *x = v + 4 * hq;
                        const float4 h = make_float4(tc::tf32_hi(x[0]), tc::tf32_hi(x[1]), tc::tf32_hi(x[2]), tc::tf32_hi(x[3]));
                        const uint32_t off = (uint32_t)(c + hq) * D2_CHUNK + (uint32_t)r * 16u;
                        *reinterpret_cast<float4 *>(sm + off) = h;
                        *reinterpret_cast<float4 *>(sm + D2F_ULO + off) = make_float4(x[0] - h.x, x[1] - h.y, x[2] - h.z, x[3] - h.w);
                    }
                }
            }
            tc::fence_proxy_async();
            tc::tc_fence_before();
            D2_TRACE(0, 8 * it + 4);
            d2_bar_sync_all();                                  // B
            tc::mbar_wait(&barM2, par);
            tc::tc_fence_after();
            D2_TRACE(0, 8 * it + 5);
            // ---- epilogue 2: bias, tanh / sigmoid, mask bits, survivor counts -> ZT ----------------------------------
            uint32_t bits = 0;
            {
                float4 *zt = a.ZT + ((size_t)tile * D2_ZCH + 8 * grp) * D2_ROWS + r;
#pragma unroll
                for (int n0 = 0; n0 < 32; n0 += 8) {
                    float v[8];
                    tc::tmem_ld8(tlane + 96 + 32 * grp + n0, v);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const int j = 32 * grp + n0 + q;
                        float z = v[q] + s_b2[j];
                        if (j < KO) {
                            z = tanhf(z);
                            if (valid) {
                                a.nopac[(size_t)row * KO + j] = z;
                                a.mask_out[(size_t)row * KO + j] = z > 0.f ? 1 : 0;
                            }
                            bits |= z > 0.f ? (1u << j) : 0u;
                        } else if (j >= D2_ZCOL) {
                            z = 1.f / (1.f + expf(-z));
                        }
                        v[q] = z;
                    }
                    zt[(n0 >> 2) * D2_ROWS] = make_float4(v[0], v[1], v[2], v[3]);
                    zt[((n0 >> 2) + 1) * D2_ROWS] = make_float4(v[4], v[5], v[6], v[7]);
                }
            }
            if (grp == 0) {
                if (!valid) bits = 0;
                if (valid) a.maskbits[row] = bits;
                uint32_t cnt = __popc(bits);
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
                if (lane == 0 && cnt) atomicAdd(&a.block_sums[(tile * D2_ROWS + warp * 32) / 256], cnt);
            }
            D2_TRACE(0, 8 * it + 6);
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<256>(tmem);
}

// compaction + post-processing (gaussian_renderer/__init__.py:96-111), one thread per (anchor, offset)
__global__ void __launch_bounds__(256)
dec2_compact_kernel(int V_host, const int32_t *__restrict__ Vdev, int nch, const float4 *__restrict__ XT, const float4 *__restrict__ ZT,
                    const uint32_t *__restrict__ maskbits, const uint32_t *__restrict__ offs, float *__restrict__ xyz,
                    float *__restrict__ color, float *__restrict__ opacity, float *__restrict__ scaling, float *__restrict__ rot) {
    const int t = blockIdx.x * 256 + threadIdx.x;
    const int V = d2_count(Vdev, V_host);
    if (t >= V * KO) return;
    const int v = t / KO, k = t - v * KO;
    const uint32_t bits = maskbits[v];
    if (!((bits >> k) & 1u)) return;
    const size_t j = offs[v] + __popc(bits & ((1u << k) - 1u));
    const int tile = v >> 7, row = v & 127;
    const float *g = reinterpret_cast<const float *>(XT + (size_t)tile * nch * D2_ROWS + row);
    const float *z = reinterpret_cast<const float *>(ZT + (size_t)tile * D2_ZCH * D2_ROWS + row);
    auto G = [&](int c) { return __ldg(g + d2_tile_idx(c)); };
    auto Z = [&](int c) { return __ldg(z + d2_tile_idx(c)); };
    const int cs = FD + 3 + 3 * KO, co = FD + 3 + 3 * k;
    const float s0 = G(cs), s1 = G(cs + 1), s2 = G(cs + 2), s3 = G(cs + 3), s4 = G(cs + 4), s5 = G(cs + 5);
    xyz[3 * j] = G(FD) + G(co) * s0;
    xyz[3 * j + 1] = G(FD + 1) + G(co + 1) * s1;
    xyz[3 * j + 2] = G(FD + 2) + G(co + 2) * s2;
    color[3 * j] = Z(d2_zcol_col(k, 0)); color[3 * j + 1] = Z(d2_zcol_col(k, 1)); color[3 * j + 2] = Z(d2_zcol_col(k, 2));
    opacity[j] = Z(d2_zcol_op(k));
    float sr[7];
#pragma unroll
    for (int q = 0; q < 7; ++q) sr[q] = Z(d2_zcol_cov(k, q));
    scaling[3 * j] = s3 / (1.f + expf(-sr[0]));
    scaling[3 * j + 1] = s4 / (1.f + expf(-sr[1]));
    scaling[3 * j + 2] = s5 / (1.f + expf(-sr[2]));
    const float n = fmaxf(sqrtf(sr[3] * sr[3] + sr[4] * sr[4] + sr[5] * sr[5] + sr[6] * sr[6]), 1e-12f);
    rot[4 * j] = sr[3] / n; rot[4 * j + 1] = sr[4] / n; rot[4 * j + 2] = sr[5] / n; rot[4 * j + 3] = sr[6] / n;
}

}  // namespace splatco

namespace splatco {

// =======================================================================================================
// Backward, after the MLP kernel: parameter gradients and the BatchNorm-backward sums from the two
// weight-gradient products   gW2b[j][i] = sum_v dZ[v][j] H[v][i]   (j in the dZ operand's column order)
// and   GT[n][uc] = sum_v dH[v][n] u[v][uc].
// =======================================================================================================
constexpr int D2_GW2_LD = 112;                                // gW2b row: 96 hidden columns + column 96 = sum_v dZ[v][j] (ones row of H^T)
constexpr int D2_PART = D2_ROWS * D2_GW2_LD + D2_ROWS * 144;  // floats per CTA partial: gW2b [128][112] | GT [128][144]

// sum of the per-CTA partials (deterministic; the MLP kernel writes its TMEM accumulators once per CTA)
__global__ void __launch_bounds__(256)
dec2_reduce_kernel(int nparts, const float *__restrict__ part, float *__restrict__ red) {
    const int e = blockIdx.x * 256 + threadIdx.x;
    if (e >= D2_PART) return;
    // same summation order as a plain loop; the loads of eight partials are issued together (the kernel is a chain of
    // dependent-looking L2 reads otherwise: ncu long_scoreboard 30 warps per issue)
    float s = 0.f;
    int p = 0;
    for (; p + 8 <= nparts; p += 8) {
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = __ldg(part + (size_t)(p + q) * D2_PART + e);
#pragma unroll
        for (int q = 0; q < 8; ++q) s += v[q];
    }
    for (; p < nparts; ++p) s += __ldg(part + (size_t)p * D2_PART + e);
    red[e] = s;
}

// red -> the quantities dec_bwd_fold_kernel consumes (same definitions as the three-stage chain produced):
//   gb1 = colsum dH = GT[:, one]        S0[o]  = colsum dgeo = sum_n W1g[o][n] gb1[n]
//   S1raw[o][c] = sum_v dgeo[v][o] x[v][c] = sum_n W1g[o][n] GT[n][u(c)]
//   gW1T[k][n]: feat / dir rows straight from GT, geo rows = sum_c Wgeo'[c][o] GT[n][u(c)] + bgeo[o] gb1[n]
__global__ void __launch_bounds__(256)
dec2_expand_kernel(int DP, int LDX, const float *__restrict__ red,
                   const float *__restrict__ WpT, const float *__restrict__ WcT, const float *__restrict__ bgeo,
                   const float *__restrict__ W1T, float *__restrict__ S1, float *__restrict__ S0,
                   float *__restrict__ gW1T, float *__restrict__ gb1, float *__restrict__ gW2T, float *__restrict__ gb2) {
    const int tid = blockIdx.x * 256 + threadIdx.x, nthr = gridDim.x * 256;
    const float *gW2b = red, *GT = red + D2_ROWS * D2_GW2_LD;
    auto gt = [&](int n, int uc) { return GT[n * 144 + uc]; };
    for (int n = tid; n < HD; n += nthr) gb1[n] = gt(n, D2_ONE);
    for (int o = tid; o < 64; o += nthr) {
        float s = 0.f;
        for (int n = 0; n < HD; ++n) s = fmaf(W1T[(36 + o) * HD + n], gt(n, D2_ONE), s);
        S0[o] = s;
    }
    for (int e = tid; e < 32 * LDX; e += nthr) {
        const int o = e / LDX, c = e - o * LDX;
        float s = 0.f;
        if (c < DP) { for (int n = 0; n < HD; ++n) s = fmaf(W1T[(36 + o) * HD + n], gt(n, D2_UP0 + c), s); }
        else if (c < DP + GD) { for (int n = 0; n < HD; ++n) s = fmaf(W1T[(68 + o) * HD + n], gt(n, c - DP), s); }
        S1[e] = s;
    }
    for (int e = tid; e < XI * HD; e += nthr) {
        const int k = e / HD, n = e - k * HD;
        float s;
        if (k < FD) s = gt(n, k);
        else if (k < 36) s = gt(n, D2_UDIR + k - FD);
        else if (k < 68) {
            const int o = k - 36;
            s = bgeo[o] * gt(n, D2_ONE);
            for (int c = 0; c < DP; ++c) s = fmaf(WpT[c * 32 + o], gt(n, D2_UP0 + c), s);
        } else {
            const int o = k - 68;
            s = bgeo[32 + o] * gt(n, D2_ONE);
            for (int g = 0; g < GD; ++g) s = fmaf(WcT[g * 32 + o], gt(n, g), s);
        }
        gW1T[e] = s;
    }
    for (int e = tid; e < HD * ZD; e += nthr) {
        const int i = e / ZD, j = e - i * ZD;
        const int jb = j < KO ? j : (j < 8 * KO ? D2_RCOV + (j - KO) : (j < 11 * KO ? D2_RCOL + (j - 8 * KO) : -1));
        gW2T[e] = jb >= 0 ? gW2b[jb * D2_GW2_LD + i] : 0.f;
    }
    for (int j = tid; j < ZD; j += nthr) {
        const int jb = j < KO ? j : (j < 8 * KO ? D2_RCOV + (j - KO) : (j < 11 * KO ? D2_RCOL + (j - 8 * KO) : -1));
        gb2[j] = jb >= 0 ? gW2b[jb * D2_GW2_LD + HD] : 0.f;
    }
}

// Final input gradients.  dU already holds the MLP-path gradient of every u column plus the direct gradients of the
// post-processing (anchor / offsets / scaling); BatchNorm's backward through the folded branches is
//     dx = dU - rstd (m1 + xhat m2)        (m1, m2 from dec_bwd_fold_kernel; columns in its [P | g] order)
// One thread per visible anchor: 16-byte coalesced tile reads, 16-byte vector REDs into the feature rows and into
// channel-last plane gradients (one or two REDs per texel instead of one per channel).
template <int LEVEL, int RC, bool PACKED>
__global__ void __launch_bounds__(D2_ROWS)
dec2_bwd_inputs_kernel(DecPtrs p, DecInputGrads gi, int V, const float4 *__restrict__ XT, const float4 *__restrict__ DUT,
                       const float *__restrict__ mu, const float *__restrict__ rstd, const float *__restrict__ m1,
                       const float *__restrict__ m2) {
    constexpr int NS = LEVEL == 0 ? 6 : (LEVEL == 1 ? 9 : 12);
    constexpr int DP = NS * RC, NPC = (DP + 3) / 4, NCH = D2_P_CH0 + NPC;
    const int r = threadIdx.x;
    const int v = blockIdx.x * D2_ROWS + r;
    if (v >= V) return;
    const int i = p.vis[v];
    const float4 *xt = XT + (size_t)blockIdx.x * NCH * D2_ROWS + r;
    const float4 *du = DUT + (size_t)blockIdx.x * NCH * D2_ROWS + r;
    auto bn = [&](float x, float d, int c) {            // c: column in [P | g] order
        const float rs = __ldg(rstd + c);
        return d - rs * (__ldg(m1 + c) + (x - __ldg(mu + c)) * rs * __ldg(m2 + c));
    };
    // ---- context columns ------------------------------------------------------------------------------------------------
    {
        float *gf = gi.anchor_feat + (size_t)i * FD;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float4 x = __ldg(xt + q * D2_ROWS), d = __ldg(du + q * D2_ROWS);
            red_add_v4(gf + 4 * q, bn(x.x, d.x, DP + 4 * q), bn(x.y, d.y, DP + 4 * q + 1), bn(x.z, d.z, DP + 4 * q + 2),
                       bn(x.w, d.w, DP + 4 * q + 3));
        }
        float ga[40];
#pragma unroll
        for (int q = 0; q < 10; ++q) {
            const float4 x = __ldg(xt + (8 + q) * D2_ROWS), d = __ldg(du + (8 + q) * D2_ROWS);
            ga[4 * q] = bn(x.x, d.x, DP + FD + 4 * q);
            ga[4 * q + 1] = bn(x.y, d.y, DP + FD + 4 * q + 1);
            ga[4 * q + 2] = bn(x.z, d.z, DP + FD + 4 * q + 2);
            ga[4 * q + 3] = q < 9 ? bn(x.w, d.w, DP + FD + 4 * q + 3) : 0.f;      // column 71 is the constant
        }
        // direction / distance -> anchor
        const float4 dd = __ldg(xt + D2_DIR_CH * D2_ROWS), gd = __ldg(du + D2_DIR_CH * D2_ROWS);
        const float dot = dd.x * gd.x + dd.y * gd.y + dd.z * gd.z;
        ga[0] += (gd.x - dd.x * dot) / dd.w + gd.w * dd.x;
        ga[1] += (gd.y - dd.y * dot) / dd.w + gd.w * dd.y;
        ga[2] += (gd.z - dd.z * dot) / dd.w + gd.w * dd.z;
        float *g3 = gi.anchor + 3 * (size_t)i;
        atomicAdd(g3, ga[0]); atomicAdd(g3 + 1, ga[1]); atomicAdd(g3 + 2, ga[2]);
        float *go = gi.offset + (size_t)i * 3 * KO;       // 120-byte rows: 8-byte aligned
#pragma unroll
        for (int q = 0; q < 15; ++q) red_add_v2(go + 2 * q, ga[3 + 2 * q], ga[4 + 2 * q]);
        float *gs = gi.scaling + (size_t)i * 6;
#pragma unroll
        for (int q = 0; q < 3; ++q) red_add_v2(gs + 2 * q, ga[33 + 2 * q], ga[34 + 2 * q]);
    }
    // ---- plane columns: bilinear scatter ---------------------------------------------------------------------------------
    float pg[NPC * 4];
#pragma unroll
    for (int q = 0; q < NPC; ++q) {
        const float4 x = __ldg(xt + (D2_P_CH0 + q) * D2_ROWS), d = __ldg(du + (D2_P_CH0 + q) * D2_ROWS);
        pg[4 * q] = 4 * q < DP ? bn(x.x, d.x, 4 * q) : 0.f;
        pg[4 * q + 1] = 4 * q + 1 < DP ? bn(x.y, d.y, 4 * q + 1) : 0.f;
        pg[4 * q + 2] = 4 * q + 2 < DP ? bn(x.z, d.z, 4 * q + 2) : 0.f;
        pg[4 * q + 3] = 4 * q + 3 < DP ? bn(x.w, d.w, 4 * q + 3) : 0.f;
    }
    // (the plane-feature noise is additive: it shifts x but d x / d plane is unchanged)
    const float ax = __ldg(p.anchor + 3 * (size_t)i), ay = __ldg(p.anchor + 3 * (size_t)i + 1), az = __ldg(p.anchor + 3 * (size_t)i + 2);
    float ind[3];
    norm_coords(p, ax, ay, az, ind);
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        const int lvl = s < 6 ? 0 : (s < 9 ? 1 : 2);
        const int pl = s < 6 ? (s >> 1) : (s < 9 ? s - 6 : s - 9);
        const bool att = s < 6 && (s & 1);
        float *base = att ? gi.att[pl] : gi.plane[lvl][pl];
        if (!base) continue;
        const int E = p.E[lvl];
        float u, w;
        plane_axes(pl, ind, u, w);
        const Bilin b = bilin_setup(u, w, E);
        const int ti[4] = {b.i00, b.i01, b.i10, b.i11};
        const float tw[4] = {b.w00, b.w01, b.w10, b.w11};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            if (ti[t] < 0) continue;
            if (PACKED) {
                float *tp = base + (size_t)ti[t] * 8;
                const float g0 = pg[s * RC] * tw[t], g1 = RC > 1 ? pg[s * RC + 1] * tw[t] : 0.f, g2 = RC > 2 ? pg[s * RC + 2] * tw[t] : 0.f,
                            g3 = RC > 3 ? pg[s * RC + 3] * tw[t] : 0.f;
                red_add_v4(tp, g0, g1, g2, g3);
                if (RC > 4) atomicAdd(tp + 4, pg[s * RC + 4] * tw[t]);
            } else {
                const size_t cs = (size_t)E * E;
#pragma unroll
                for (int ch = 0; ch < RC; ++ch) atomicAdd(base + ch * cs + ti[t], pg[s * RC + ch] * tw[t]);
            }
        }
    }
}

}  // namespace splatco
