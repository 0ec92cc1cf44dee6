// Fused anchor-decode BACKWARD row chain on tensor cores (tcgen05, TMEM, 3xTF32); included by decode.cu
// after decode_tc.cuh.  Per 128-anchor tile, three chained GEMM stages (the transposes of the forward):
//   1  dH [128, 96] = (dZ[128,112] * W2) .* [H > 0]
//   2  dX [128,100] = dH * W1                      (columns 36..99 = dgeo)
//   3  dxhat[128, DP | 71] = dgeo_p * (gamma W)_planes | dgeo_c * (gamma W)_context
// and, in the epilogues, the three column sums the parameter gradients need (gb2 = colsum dZ,
// gb1 = colsum dH, S0 = colsum dgeo), reduced with a warp butterfly and added with fp32 REDs.
// The reductions over anchors that produce weight gradients (H^T dZ, X100^T dH, dgeo^T X) run in
// dec_wgrad_kernel below (each CTA keeps a whole output in registers); on tensor cores they need MN-major operand
// tiles of both the forward and the backward activations, which do not fit next to this chain's operands.
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
    const int tid = blockIdx.x * 256 + threadIdx.x, nthr = gridDim.x * 256;
    auto put = [](uint8_t *dst, uint32_t half, size_t e, const float *w) {
        const float4 h = make_float4(tc::tf32_hi(w[0]), tc::tf32_hi(w[1]), tc::tf32_hi(w[2]), tc::tf32_hi(w[3]));
        *reinterpret_cast<float4 *>(dst + e * 16) = h;
        *reinterpret_cast<float4 *>(dst + half + e * 16) = make_float4(w[0] - h.x, w[1] - h.y, w[2] - h.z, w[3] - h.w);
    };
    for (int e = tid; e < 28 * HD; e += nthr) {                // B[i][j] = W2T[i][j]
        const int c = e / HD, i = e - c * HD;
        float w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) w[q] = W2T[i * ZD + 4 * c + q];
        put(W2R, TCB_W2R_HALF, e, w);
    }
    for (int e = tid; e < 24 * ZD; e += nthr) {                // B[n][k] = W1T[n][k], rows >= 100 zero
        const int c = e / ZD, n = e - c * ZD;
        float w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) w[q] = n < XI ? W1T[n * HD + 4 * c + q] : 0.f;
        put(W1R, TCB_W1R_HALF, e, w);
    }
    for (int e = tid; e < 8 * TCB_NP; e += nthr) {             // B[c][o] = WpG[o][c]
        const int c8 = e / TCB_NP, n = e - c8 * TCB_NP;
        float w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) w[q] = n < DP ? WpG[(4 * c8 + q) * DP + n] : 0.f;
        put(WPCR, TCB_WPC_HALF, e, w);
    }
    for (int e = tid; e < 8 * TCB_NC; e += nthr) {             // B[g][o] = WcG[o][g]
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

// 512 threads as in the forward: warp w <-> rows 32 (w & 3) .. +31 (its TMEM lane quarter), column group w >> 2.
__global__ void __launch_bounds__(TC_THREADS, 1)
dec_tc_bwd_kernel(int V, int DP, int LDX, const float *__restrict__ DZ, const float *__restrict__ Hs,
                  const uint8_t *__restrict__ W2R, const uint8_t *__restrict__ W1R, const uint8_t *__restrict__ WPCR,
                  float *__restrict__ DH, float *__restrict__ DX, float *__restrict__ DXH, float *__restrict__ gb2,
                  float *__restrict__ gb1, float *__restrict__ S0) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t barL, barM;
    __shared__ uint32_t tmem_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r = tid & (TC_ROWS - 1), grp = tid >> 7;
    if (warp == 0) tc::tmem_alloc<256>(&tmem_s);
    if (tid == 0) { tc::mbar_init(&barL, 1); tc::mbar_init(&barM, 1); tc::fence_barrier_init(); }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_s;
    const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t a_hi = tc::smem_u32(sm), a_lo = a_hi + TCB_AREG, b_hi = a_hi + TCB_OFF_B;
    // 8-column groups of each stage dealt to the four column groups: 14 -> 4,4,3,3 | 12 -> 3,3,3,3 | 13 -> 4,3,3,3 | 18 -> 5,5,4,4
    const int z_beg = (grp < 2 ? grp * 4 : 8 + (grp - 2) * 3) * 8, z_end = z_beg + (grp < 2 ? 32 : 24);
    const int h_beg = grp * 24, h_end = h_beg + 24;
    const int x_beg = (grp == 0 ? 0 : 4 + (grp - 1) * 3) * 8, x_end = x_beg + (grp == 0 ? 32 : 24);
    const int o_beg = (grp < 2 ? grp * 5 : 10 + (grp - 2) * 4) * 8, o_end = o_beg + (grp < 2 ? 40 : 32);
    static_assert(ZD == 112 && HD == 96 && TCB_NP + TCB_NC == 144, "column-group split of the decode backward epilogues");
    constexpr uint32_t idesc96 = tc::make_idesc_tf32(128, HD), idesc112 = tc::make_idesc_tf32(128, ZD),
                       idesc64 = tc::make_idesc_tf32(128, TCB_NP), idesc80 = tc::make_idesc_tf32(128, TCB_NC);
    uint32_t phL = 0, phM = 0;
    const int ntiles = (V + TC_ROWS - 1) / TC_ROWS;

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int row = tile * TC_ROWS + r;
        const bool valid = row < V;
        // ---- stage 1 operands: W2R by bulk copy, dZ rows split by the threads -----------------------------------
        if (tid == 0) {
            tc::mbar_arrive_expect_tx(&barL, 2 * TCB_W2R_HALF);
            tc::bulk_g2s(sm + TCB_OFF_B, W2R, 2 * TCB_W2R_HALF, &barL);
        }
        const float4 *zrow = reinterpret_cast<const float4 *>(DZ + (size_t)row * ZD);
        // the ReLU gate of epilogue 1 as 96 bits, so that its H row loads are issued here, all independent,
        // instead of one dependent round trip per 8 columns inside the TMEM read-out loop
        uint32_t hbits = 0u;                                     // bit q <-> hidden column h_beg + q (this thread's 24 columns)
        if (valid) {
            const float4 *hrow = reinterpret_cast<const float4 *>(Hs + (size_t)row * HD + h_beg);
            float4 h[6];
#pragma unroll
            for (int q = 0; q < 6; ++q) h[q] = __ldg(hrow + q);
#pragma unroll
            for (int q = 0; q < 6; ++q)
                hbits |= (h[q].x > 0.f ? 1u : 0u) << (4 * q) | (h[q].y > 0.f ? 2u : 0u) << (4 * q) |
                         (h[q].z > 0.f ? 4u : 0u) << (4 * q) | (h[q].w > 0.f ? 8u : 0u) << (4 * q);
        }
        {                                                        // this thread's 6 or 8 chunks of the dZ row, all loads in flight
            float4 x[8];
            const int cb = z_beg >> 2, nc = (z_end - z_beg) >> 2;
#pragma unroll
            for (int q = 0; q < 8; ++q) x[q] = (valid && q < nc) ? __ldg(zrow + cb + q) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int q = 0; q < 8; q += 2) {
                if (q < nc) {                                    // warp-uniform
                    const int c = cb + q;
                    const float v[8] = {x[q].x, x[q].y, x[q].z, x[q].w, x[q + 1].x, x[q + 1].y, x[q + 1].z, x[q + 1].w};
                    tcb_store_split(sm, c, r, v);
                    tcb_store_split(sm, c + 1, r, v + 4);
                    colsum8_add(v, lane, gb2 + 4 * c, ZD - 4 * c);
                }
            }
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
#pragma unroll
        for (int n0 = h_beg; n0 < h_end; n0 += 8) {
            float v[8];
            tc::tmem_ld8(tlane + n0, v);
            tc::tmem_ld_wait();
            if (valid) {
                const uint32_t gate = hbits >> (n0 - h_beg);
#pragma unroll
                for (int q = 0; q < 8; ++q) v[q] = (gate >> q) & 1u ? v[q] : 0.f;
                float4 *dst = reinterpret_cast<float4 *>(DH + (size_t)row * HD + n0);
                dst[0] = make_float4(v[0], v[1], v[2], v[3]);
                dst[1] = make_float4(v[4], v[5], v[6], v[7]);
            } else {
#pragma unroll
                for (int q = 0; q < 8; ++q) v[q] = 0.f;
            }
            tcb_store_split(sm, n0 / 4, r, v);
            tcb_store_split(sm, n0 / 4 + 1, r, v + 4);
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
        for (int n0 = x_beg; n0 < x_end; n0 += 8) {
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
            if (n0 + 4 >= 36 && n0 + 4 < XI) tcb_store_split(sm, (n0 + 4 - 36) / 4, r, v + 4);
            if (n0 >= 36) tcb_store_split(sm, (n0 - 36) / 4, r, v);
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
            for (int n0 = o_beg; n0 < o_end; n0 += 8) {
                float v[8];
                tc::tmem_ld8(tlane + n0, v);          // warp-collective: every lane executes it, stores are predicated
                tc::tmem_ld_wait();
                const int base = n0 < TCB_NP ? n0 : DP + (n0 - TCB_NP);
                const int lim = n0 < TCB_NP ? DP : DP + GD;
                if (valid) {
                    // rows are 16-byte aligned (LDX % 4 == 0); whole groups that start on a 16-byte boundary (always in the
                    // plane block; in the context block when DP = 6/9/12 rc is a multiple of 4) go out as two 16-byte
                    // stores (scalar stores of this epilogue were 144 of the ~200 store instructions per row and showed
                    // up as lg_throttle stalls, ncu r1z)
                    if (base + 8 <= lim && (base & 3) == 0) {
                        float4 *dst = reinterpret_cast<float4 *>(drow + base);
                        dst[0] = make_float4(v[0], v[1], v[2], v[3]);
                        dst[1] = make_float4(v[4], v[5], v[6], v[7]);
                    } else {
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            if (base + q < lim) drow[base + q] = v[q];
                    }
                }
            }
        }
        tc::tc_fence_before();
        __syncthreads();
        tc::tc_fence_after();
    }
    if (warp == 0) tc::tmem_dealloc<256>(tmem);
}

// ---- weight-gradient reductions over anchors:  C_p[M,N] += A_p[:, :M]^T B_p[:, :N]  ---------------------------
// Every product of a backward pass has a SMALL output (at most 100 x 96) and a long reduction (K = V visible
// anchors), so one CTA keeps a whole output matrix in registers (16 x 16 threads, TM x TN each) and walks a
// slice of the rows: operands are read exactly once, staged 32 rows at a time through shared memory with the
// next block's global loads in flight during the FMAs, and the partial result is added with fp32 REDs.
// CTAs are dealt to the problems in proportion to their FMA count (blockIdx.x -> problem via cta0[]).
struct WgProblem { const float *A; const float *B; float *C; int M, N, lda, ldb, ldc, cta0, nslices, shape; };
constexpr int WG_MAX_PROBLEMS = 8;
struct WgGroup { WgProblem p[WG_MAX_PROBLEMS]; int count; };
constexpr int WG_KB = 32;                        // rows per shared-memory block
constexpr int WG_SMEM_FLOATS = WG_KB * (128 + 96);

template <int TM, int TN>
__device__ __forceinline__ void wgrad_body(const WgProblem &pr, int kbeg, int kend, float *smem) {
    constexpr int MP = TM * 16, NP = TN * 16;
    constexpr int NA = (WG_KB * MP + 255) / 256, NB = (WG_KB * NP + 255) / 256;
    float *As = smem, *Bs = smem + WG_KB * MP;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const float *__restrict__ A = pr.A;
    const float *__restrict__ B = pr.B;
    const int M = pr.M, N = pr.N, lda = pr.lda, ldb = pr.ldb;
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    float ra[NA], rb[NB];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int q = 0; q < NA; ++q) {
            const int e = tid + q * 256, r = e / MP, c = e - r * MP;
            ra[q] = (e < WG_KB * MP && c < M && k0 + r < kend) ? __ldg(A + (size_t)(k0 + r) * lda + c) : 0.f;
        }
#pragma unroll
        for (int q = 0; q < NB; ++q) {
            const int e = tid + q * 256, r = e / NP, c = e - r * NP;
            rb[q] = (e < WG_KB * NP && c < N && k0 + r < kend) ? __ldg(B + (size_t)(k0 + r) * ldb + c) : 0.f;
        }
    };
    fetch(kbeg);
    for (int k0 = kbeg; k0 < kend; k0 += WG_KB) {
#pragma unroll
        for (int q = 0; q < NA; ++q) { const int e = tid + q * 256; if (e < WG_KB * MP) As[e] = ra[q]; }
#pragma unroll
        for (int q = 0; q < NB; ++q) { const int e = tid + q * 256; if (e < WG_KB * NP) Bs[e] = rb[q]; }
        __syncthreads();
        if (k0 + WG_KB < kend) fetch(k0 + WG_KB);
#pragma unroll 4
        for (int k = 0; k < WG_KB; ++k) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = As[k * MP + ty * TM + i];
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = Bs[k * NP + tx * TN + j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    float *__restrict__ C = pr.C;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = ty * TM + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = tx * TN + j;
            if (n < N && acc[i][j] != 0.f) atomicAdd(&C[(size_t)m * pr.ldc + n], acc[i][j]);
        }
    }
}

// output-shape classes (TM x TN per thread): 0: 32 x <=16   1: 32 x <=32   2: 32 x <=64   3: 32 x <=80   4: <=128 x <=96
__global__ void __launch_bounds__(256)
dec_wgrad_kernel(WgGroup g, int K) {
    __shared__ __align__(16) float smem[WG_SMEM_FLOATS];
    int pi = 0;
#pragma unroll
    for (int i = 1; i < WG_MAX_PROBLEMS; ++i)
        if (i < g.count && (int)blockIdx.x >= g.p[i].cta0) pi = i;
    const WgProblem &pr = g.p[pi];
    const int slice = blockIdx.x - pr.cta0;
    const int rows = (K + pr.nslices - 1) / pr.nslices;
    const int per = (rows + WG_KB - 1) / WG_KB * WG_KB;         // slices start on block boundaries
    const int kbeg = slice * per, kend = min(K, kbeg + per);
    if (kbeg >= kend) return;
    switch (pr.shape) {
        case 0: wgrad_body<2, 1>(pr, kbeg, kend, smem); break;
        case 1: wgrad_body<2, 2>(pr, kbeg, kend, smem); break;
        case 2: wgrad_body<2, 4>(pr, kbeg, kend, smem); break;
        case 3: wgrad_body<2, 5>(pr, kbeg, kend, smem); break;
        default: wgrad_body<8, 6>(pr, kbeg, kend, smem); break;
    }
}

}  // namespace splatco
