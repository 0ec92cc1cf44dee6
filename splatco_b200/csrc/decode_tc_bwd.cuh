// Fused anchor-decode BACKWARD row chain on tensor cores (tcgen05, TMEM, 3xTF32); included by decode.cu
// after decode_tc.cuh.  Per 128-anchor tile, three chained GEMM stages (the transposes of the forward):
//   1  dH [128, 96] = (dZ[128,112] * W2) .* [H > 0]
//   2  dX [128,100] = dH * W1                      (columns 36..99 = dgeo)
//   3  dxhat[128, DP | 71] = dgeo_p * (gamma W)_planes | dgeo_c * (gamma W)_context
// and, in the epilogues, the three column sums the parameter gradients need (gb2 = colsum dZ,
// gb1 = colsum dH, S0 = colsum dgeo), reduced with a warp butterfly and added with fp32 REDs.
// The reductions over anchors that produce weight gradients (H^T dZ, X100^T dH, dgeo^T X) stay in the
// split-K SGEMM (decode.cu): their operands would have to be transposed through shared memory.
//
// Shared memory map: A hi [0, 57344) 28 chunks | A lo [57344, 114688) | weights [114688, 200704)
#pragma once
#include "tc.cuh"

namespace splatco {

constexpr int TCB_ACH = 28;                                   // dZ: 112 columns
constexpr uint32_t TCB_AREG = TCB_ACH * TC_CHUNK;             // 57344
constexpr uint32_t TCB_OFF_B = 2 * TCB_AREG;                  // 114688
constexpr uint32_t TCB_W2R_HALF = 28 * HD * 16;               // 43008: rows = hidden i (96), K = output j (112)
constexpr uint32_t TCB_W1R_HALF = 24 * ZD * 16;               // 43008: rows = x100 index (100 -> 112), K = hidden (96)
constexpr int TCB_NP = 64, TCB_NC = 80;                       // stage-3 output widths (DP <= 64, 71 -> 80)
constexpr uint32_t TCB_WPC_HALF = 8 * (TCB_NP + TCB_NC) * 16; // 18432: [WpR 64 rows | WcR 80 rows] x 8 chunks (K = 32)
constexpr uint32_t TCB_SMEM = TCB_OFF_B + 2 * TCB_W2R_HALF;   // 200704

// weight tiles of the backward chain (hi/lo canonical), packed once per view next to the forward ones
__global__ void __launch_bounds__(256)
dec_tc_pack_bwd_kernel(int DP, const float *__restrict__ W2T, const float *__restrict__ W1T,
                       const float *__restrict__ WpG, const float *__restrict__ WcG, uint8_t *__restrict__ W2R,
                       uint8_t *__restrict__ W1R, uint8_t *__restrict__ WPCR) {
    const int tid = threadIdx.x;
    auto put = [](uint8_t *dst, uint32_t half, size_t e, const float *w) {
        const float4 h = make_float4(tc::tf32_hi(w[0]), tc::tf32_hi(w[1]), tc::tf32_hi(w[2]), tc::tf32_hi(w[3]));
        *reinterpret_cast<float4 *>(dst + e * 16) = h;
        *reinterpret_cast<float4 *>(dst + half + e * 16) = make_float4(w[0] - h.x, w[1] - h.y, w[2] - h.z, w[3] - h.w);
    };
    for (int e = tid; e < 28 * HD; e += 256) {                // B[i][j] = W2T[i][j]
        const int c = e / HD, i = e - c * HD;
        float w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) w[q] = W2T[i * ZD + 4 * c + q];
        put(W2R, TCB_W2R_HALF, e, w);
    }
    for (int e = tid; e < 24 * ZD; e += 256) {                // B[n][k] = W1T[n][k], rows >= 100 zero
        const int c = e / ZD, n = e - c * ZD;
        float w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) w[q] = n < XI ? W1T[n * HD + 4 * c + q] : 0.f;
        put(W1R, TCB_W1R_HALF, e, w);
    }
    for (int e = tid; e < 8 * TCB_NP; e += 256) {             // B[c][o] = WpG[o][c]
        const int c8 = e / TCB_NP, n = e - c8 * TCB_NP;
        float w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) w[q] = n < DP ? WpG[(4 * c8 + q) * DP + n] : 0.f;
        put(WPCR, TCB_WPC_HALF, e, w);
    }
    for (int e = tid; e < 8 * TCB_NC; e += 256) {             // B[g][o] = WcG[o][g]
        const int c8 = e / TCB_NC, n = e - c8 * TCB_NC;
        float w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) w[q] = n < GD ? WcG[(4 * c8 + q) * GD + n] : 0.f;
        put(WPCR, TCB_WPC_HALF, (size_t)8 * TCB_NP + e, w);
    }
}

__device__ __forceinline__ void tcb_store_split(uint8_t *sm, int chunk, int row, const float *v4) {
    const float4 h = make_float4(tc::tf32_hi(v4[0]), tc::tf32_hi(v4[1]), tc::tf32_hi(v4[2]), tc::tf32_hi(v4[3]));
    const uint32_t off = (uint32_t)chunk * TC_CHUNK + (uint32_t)row * 16u;
    *reinterpret_cast<float4 *>(sm + off) = h;
    *reinterpret_cast<float4 *>(sm + TCB_AREG + off) = make_float4(v4[0] - h.x, v4[1] - h.y, v4[2] - h.z, v4[3] - h.w);
}

// column sums of 8 per-row values over the 32 rows of a warp, added to dst[0..7] (fp32 RED)
__device__ __forceinline__ void colsum8_add(const float (&g)[8], int lane, float *dst, int ncols_valid) {
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
    const int idx = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
    if ((lane & 3) == 0 && idx < ncols_valid && r != 0.f) atomicAdd(dst + idx, r);
}

__global__ void __launch_bounds__(TC_ROWS, 1)
dec_tc_bwd_kernel(int V, int DP, int LDX, const float *__restrict__ DZ, const float *__restrict__ Hs,
                  const uint8_t *__restrict__ W2R, const uint8_t *__restrict__ W1R, const uint8_t *__restrict__ WPCR,
                  float *__restrict__ DH, float *__restrict__ DX, float *__restrict__ DXH, float *__restrict__ gb2,
                  float *__restrict__ gb1, float *__restrict__ S0) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t barL, barM;
    __shared__ uint32_t tmem_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tc::tmem_alloc<256>(&tmem_s);
    if (tid == 0) { tc::mbar_init(&barL, 1); tc::mbar_init(&barM, 1); tc::fence_barrier_init(); }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_s;
    const uint32_t tlane = tmem + ((uint32_t)(warp * 32) << 16);
    const uint32_t a_hi = tc::smem_u32(sm), a_lo = a_hi + TCB_AREG, b_hi = a_hi + TCB_OFF_B;
    constexpr uint32_t idesc96 = tc::make_idesc_tf32(128, HD), idesc112 = tc::make_idesc_tf32(128, ZD),
                       idesc64 = tc::make_idesc_tf32(128, TCB_NP), idesc80 = tc::make_idesc_tf32(128, TCB_NC);
    uint32_t phL = 0, phM = 0;
    const int ntiles = (V + TC_ROWS - 1) / TC_ROWS;

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int row = tile * TC_ROWS + tid;
        const bool valid = row < V;
        // ---- stage 1 operands: W2R by bulk copy, dZ rows split by the threads -----------------------------------
        if (tid == 0) {
            tc::mbar_arrive_expect_tx(&barL, 2 * TCB_W2R_HALF);
            tc::bulk_g2s(sm + TCB_OFF_B, W2R, 2 * TCB_W2R_HALF, &barL);
        }
        const float4 *zrow = reinterpret_cast<const float4 *>(DZ + (size_t)row * ZD);
#pragma unroll 1
        for (int c = 0; c < TCB_ACH; c += 2) {
            float v[8];
            if (valid) {
                const float4 x0 = zrow[c], x1 = zrow[c + 1];
                v[0] = x0.x; v[1] = x0.y; v[2] = x0.z; v[3] = x0.w; v[4] = x1.x; v[5] = x1.y; v[6] = x1.z; v[7] = x1.w;
            } else {
#pragma unroll
                for (int q = 0; q < 8; ++q) v[q] = 0.f;
            }
            tcb_store_split(sm, c, tid, v);
            tcb_store_split(sm, c + 1, tid, v + 4);
            colsum8_add(v, lane, gb2 + 4 * c, ZD - 4 * c);
        }
        tc::fence_proxy_async();
        tc::tc_fence_before();
        __syncthreads();
        tc::tc_fence_after();
        tc::mbar_wait(&barL, phL); phL ^= 1;
        if (tid == 0) {
            tc::issue_3xtf32(tmem, a_hi, a_lo, TC_ROWS, 0, b_hi, b_hi + TCB_W2R_HALF, HD, 0, ZD / 8, idesc96, false);
            tc::mma_commit(&barM);
        }
        tc::mbar_wait(&barM, phM); phM ^= 1;
        tc::tc_fence_after();
        if (tid == 0) {
            tc::mbar_arrive_expect_tx(&barL, 2 * TCB_W1R_HALF);
            tc::bulk_g2s(sm + TCB_OFF_B, W1R, 2 * TCB_W1R_HALF, &barL);
        }
        // ---- epilogue 1: ReLU gate, dH -> global + stage-2 operand chunks 0..23, gb1 ------------------------------
        const float4 *hrow = reinterpret_cast<const float4 *>(Hs + (size_t)row * HD);
#pragma unroll 1
        for (int n0 = 0; n0 < HD; n0 += 8) {
            float v[8];
            tc::tmem_ld8(tlane + n0, v);
            tc::tmem_ld_wait();
            if (valid) {
                const float4 h0 = hrow[n0 / 4], h1 = hrow[n0 / 4 + 1];
                v[0] = h0.x > 0.f ? v[0] : 0.f; v[1] = h0.y > 0.f ? v[1] : 0.f; v[2] = h0.z > 0.f ? v[2] : 0.f; v[3] = h0.w > 0.f ? v[3] : 0.f;
                v[4] = h1.x > 0.f ? v[4] : 0.f; v[5] = h1.y > 0.f ? v[5] : 0.f; v[6] = h1.z > 0.f ? v[6] : 0.f; v[7] = h1.w > 0.f ? v[7] : 0.f;
                float4 *dst = reinterpret_cast<float4 *>(DH + (size_t)row * HD + n0);
                dst[0] = make_float4(v[0], v[1], v[2], v[3]);
                dst[1] = make_float4(v[4], v[5], v[6], v[7]);
            } else {
#pragma unroll
                for (int q = 0; q < 8; ++q) v[q] = 0.f;
            }
            tcb_store_split(sm, n0 / 4, tid, v);
            tcb_store_split(sm, n0 / 4 + 1, tid, v + 4);
            colsum8_add(v, lane, gb1 + n0, 8);
        }
        tc::fence_proxy_async();
        tc::tc_fence_before();
        __syncthreads();
        tc::tc_fence_after();
        tc::mbar_wait(&barL, phL); phL ^= 1;
        if (tid == 0) {
            tc::issue_3xtf32(tmem, a_hi, a_lo, TC_ROWS, 0, b_hi, b_hi + TCB_W1R_HALF, ZD, 0, HD / 8, idesc112, false);
            tc::mma_commit(&barM);
        }
        tc::mbar_wait(&barM, phM); phM ^= 1;
        tc::tc_fence_after();
        if (tid == 0) {
            tc::mbar_arrive_expect_tx(&barL, 2 * TCB_WPC_HALF);
            tc::bulk_g2s(sm + TCB_OFF_B, WPCR, 2 * TCB_WPC_HALF, &barL);
        }
        // ---- epilogue 2: dX -> global (100 columns); dgeo = columns 36..99 -> stage-3 operand chunks 0..15, S0 ------
#pragma unroll 1
        for (int n0 = 0; n0 < 104; n0 += 8) {
            float v[8];
            tc::tmem_ld8(tlane + n0, v);
            tc::tmem_ld_wait();
            if (!valid) {
#pragma unroll
                for (int q = 0; q < 8; ++q) v[q] = 0.f;
            }
            if (valid) {
                float4 *dst = reinterpret_cast<float4 *>(DX + (size_t)row * XI + n0);
                dst[0] = make_float4(v[0], v[1], v[2], v[3]);
                if (n0 + 4 < XI) dst[1] = make_float4(v[4], v[5], v[6], v[7]);
            }
            // dgeo chunk index of column n: (n - 36) / 4
            if (n0 + 4 >= 36 && n0 + 4 < XI) tcb_store_split(sm, (n0 + 4 - 36) / 4, tid, v + 4);
            if (n0 >= 36) tcb_store_split(sm, (n0 - 36) / 4, tid, v);
            // S0[o] = colsum(dgeo[:, o]), o = n - 36
            if (n0 >= 36) {
                colsum8_add(v, lane, S0 + (n0 - 36), min(8, XI - n0));
            } else if (n0 + 8 > 36) {                       // n0 = 32: columns 36..39 are v[4..7]
                const float t[8] = {v[4], v[5], v[6], v[7], 0.f, 0.f, 0.f, 0.f};
                colsum8_add(t, lane, S0, 4);
            }
        }
        tc::fence_proxy_async();
        tc::tc_fence_before();
        __syncthreads();
        tc::tc_fence_after();
        tc::mbar_wait(&barL, phL); phL ^= 1;
        if (tid == 0) {
            tc::issue_3xtf32(tmem, a_hi, a_lo, TC_ROWS, 0, b_hi, b_hi + TCB_WPC_HALF, TCB_NP, 0, 4, idesc64, false);
            tc::issue_3xtf32(tmem + TCB_NP, a_hi, a_lo, TC_ROWS, 8, b_hi + 8 * TCB_NP * 16, b_hi + TCB_WPC_HALF + 8 * TCB_NP * 16,
                             TCB_NC, 0, 4, idesc80, false);
            tc::mma_commit(&barM);
        }
        tc::mbar_wait(&barM, phM); phM ^= 1;
        tc::tc_fence_after();
        // ---- epilogue 3: dxhat rows -> global [V, LDX] ---------------------------------------------------------------
        {
            float *drow = DXH + (size_t)row * LDX;
#pragma unroll 1
            for (int n0 = 0; n0 < TCB_NP + TCB_NC; n0 += 8) {
                float v[8];
                tc::tmem_ld8(tlane + n0, v);          // warp-collective: every lane executes it, stores are predicated
                tc::tmem_ld_wait();
                const int base = n0 < TCB_NP ? n0 : DP + (n0 - TCB_NP);
                const int lim = n0 < TCB_NP ? DP : DP + GD;
                if (valid) {
#pragma unroll
                    for (int q = 0; q < 8; ++q)
                        if (base + q < lim) drow[base + q] = v[q];
                }
            }
        }
        tc::tc_fence_before();
        __syncthreads();
        tc::tc_fence_after();
    }
    if (warp == 0) tc::tmem_dealloc<256>(tmem);
}

// ---- grouped split-K  C_p[M,N] += A_p^T B_p  (weight-gradient reductions over anchors) ------------------------
// All products of one backward pass share K = V, so they are launched as ONE grid: blockIdx.y walks the
// 64x64 output tiles of every problem, blockIdx.x the K slices.  Each problem alone is a single wave of
// latency-bound CTAs; together they keep ~10 CTAs per SM in flight.  4x4 register blocks.
struct TnProblem { const float *A; const float *B; float *C; int M, N, lda, ldb, ldc, tile0, tiles_n; };
constexpr int TN_MAX_PROBLEMS = 8;
struct TnGroup { TnProblem p[TN_MAX_PROBLEMS]; int count; };

__global__ void __launch_bounds__(256)
sgemm_tn64_grouped_kernel(TnGroup g, int K, int kchunk) {
    __shared__ __align__(16) float As[16][68];
    __shared__ __align__(16) float Bs[16][68];
    int pi = 0;
#pragma unroll
    for (int i = 1; i < TN_MAX_PROBLEMS; ++i)
        if (i < g.count && (int)blockIdx.y >= g.p[i].tile0) pi = i;
    const TnProblem &pr = g.p[pi];
    const int t = blockIdx.y - pr.tile0;
    const int m0 = (t / pr.tiles_n) * 64, n0 = (t % pr.tiles_n) * 64;
    const int M = pr.M, N = pr.N, lda = pr.lda, ldb = pr.ldb;
    const float *__restrict__ A = pr.A;
    const float *__restrict__ B = pr.B;
    const int tid = threadIdx.x;
    const int kbeg = blockIdx.x * kchunk, kend = min(K, kbeg + kchunk);
    const int ty = tid >> 4, tx = tid & 15;
    float acc[4][4] = {};
    for (int k0 = kbeg; k0 < kend; k0 += 16) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = tid + i * 256, c = e & 63, k = e >> 6;
            const int gk = k0 + k;
            As[k][c] = (gk < kend && m0 + c < M) ? A[(size_t)gk * lda + m0 + c] : 0.f;
            Bs[k][c] = (gk < kend && n0 + c < N) ? B[(size_t)gk * ldb + n0 + c] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const float4 a = *reinterpret_cast<const float4 *>(&As[k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4 *>(&Bs[k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
    float *__restrict__ C = pr.C;
    const int ldc = pr.ldc;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gm = m0 + ty * 4 + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gn = n0 + tx * 4 + j;
            if (gn < N && acc[i][j] != 0.f) atomicAdd(&C[(size_t)gm * ldc + gn], acc[i][j]);
        }
    }
}

}  // namespace splatco
