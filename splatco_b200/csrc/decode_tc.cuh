// Fused anchor-decode forward on 5th-gen tensor cores (tcgen05, TMEM accumulators, 3xTF32).
// Included by decode.cu after its constants (KO, FD, GD, XI, HD, ZD).
//
// One persistent CTA per SM, 128 visible anchors per tile (UMMA M = 128 = TMEM lanes), 512 threads: warp w
// works on anchor rows (= TMEM lanes) 32 (w & 3) .. +31 -- the lane quarter a warp may read with tcgen05.ld -- and on
// column group w >> 2 of every operand load and epilogue, so four warps share each row block and the global loads /
// TMEM read-outs / stores of a tile run four-wide (the chain is latency bound: with one warp per lane quarter a tile
// took ~24 us).  Per tile, three chained GEMM stages whose accumulators never leave the SM:
//   A  geo[128,64]  = [P | g] * [Wp' | Wc']          (BatchNorm folded into the weights by dec_fold_kernel)
//   B  H  [128,96]  = relu([feat | dir,dist | geo] * W1 + b1)
//   C  Z  [128,112] = H * W2(block-diagonal) + b2  -> tanh / sigmoid / mask bits in the epilogue
// A operands are split hi/lo by the threads into the canonical no-swizzle K-major layout (tc.cuh); the
// weight tiles are pre-split by dec_tc_pack_kernel and arrive by 1-D bulk TMA copies (cp.async.bulk)
// signalled on an mbarrier; the epilogue of stage A/B writes the next stage's A operand straight into
// shared memory, so geo and H only touch HBM once (saved for the backward).
//
// Shared memory map (bytes; chunk = 128 rows x 16 B = 2 KB; every operand exists as a hi and a lo tile):
//   [      0,  73728)  A hi region, 36 chunks:  0 dir/dist | 1..16 P (then geo) | 17 zero | 18..35 g (18..25 = feat)
//   [  73728, 147456)  A lo region, same structure
//   [ 147456, 227328)  weights of stage A (34816 B) then stage B (79872 B)
//   stage C:  H hi chunks 0..23 at 0, H lo at 73728, W2 (86016 B) at 122880
#pragma once
#include "tc.cuh"

namespace splatco {

constexpr int TC_ROWS = 128;
constexpr int TC_THREADS = 512;                             // 4 column groups x 4 TMEM lane quarters
constexpr int TC_NP = 16;                                   // P chunks in shared memory (3*rc real ones)
constexpr int TC_ACH = 36;                                  // chunks per A region
constexpr uint32_t TC_CHUNK = TC_ROWS * 16;                 // 2048
constexpr uint32_t TC_AREG = TC_ACH * TC_CHUNK;             // 73728
constexpr uint32_t TC_OFF_B = 2 * TC_AREG;                  // 147456
constexpr uint32_t TC_OFF_W2 = TC_AREG + 24 * TC_CHUNK;     // 122880
constexpr uint32_t TC_BA_HALF = 34 * 32 * 16;               // 17408: [Wp 16 chunks | Wc 18 chunks] x 32 rows
constexpr uint32_t TC_W1_HALF = 26 * 96 * 16;               // 39936
constexpr uint32_t TC_W2_HALF = 24 * 112 * 16;              // 43008
constexpr uint32_t TC_SMEM = TC_OFF_B + 2 * TC_W1_HALF;     // 227328
constexpr int TC_MAX_RC = 5;

// global tile layout: ceil(DP/4) plane chunks (zero padded) | 18 context chunks | 1 chunk (dir.xyz, dist)
__host__ __device__ inline int tc_tile_chunks(int DP) { return ((DP + 3) >> 2) + 19; }

// Pack the folded weights into hi/lo canonical B tiles (one CTA, after dec_fold_kernel).
__global__ void __launch_bounds__(256)
dec_tc_pack_kernel(int DP, const float *__restrict__ WpT, const float *__restrict__ WcT, const float *__restrict__ W1T,
                   const float *__restrict__ W2T, uint8_t *__restrict__ BA, uint8_t *__restrict__ W1B,
                   uint8_t *__restrict__ W2B) {
    const int tid = blockIdx.x * 256 + threadIdx.x, nthr = gridDim.x * 256;
    // stage A: rows n < 32; chunks 0..15 Wp (k = P channel), 16..33 Wc (k = g channel)
    for (int e = tid; e < 34 * 32; e += nthr) {
        const int c = e >> 5, n = e & 31;
        float w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (c < 16) { const int k = 4 * c + q; w[q] = k < DP ? WpT[k * 32 + n] : 0.f; }
            else { const int k = 4 * (c - 16) + q; w[q] = k < GD ? WcT[k * 32 + n] : 0.f; }
        }
        const float4 h = make_float4(tc::tf32_hi(w[0]), tc::tf32_hi(w[1]), tc::tf32_hi(w[2]), tc::tf32_hi(w[3]));
        *reinterpret_cast<float4 *>(BA + (size_t)e * 16) = h;
        *reinterpret_cast<float4 *>(BA + TC_BA_HALF + (size_t)e * 16) = make_float4(w[0] - h.x, w[1] - h.y, w[2] - h.z, w[3] - h.w);
    }
    // stage B: rows n < 96, K order = x100 (feat | dir,dist | geo), 26 chunks
    for (int e = tid; e < 26 * 96; e += nthr) {
        const int c = e / 96, n = e - c * 96;
        float w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { const int k = 4 * c + q; w[q] = k < XI ? W1T[k * HD + n] : 0.f; }
        const float4 h = make_float4(tc::tf32_hi(w[0]), tc::tf32_hi(w[1]), tc::tf32_hi(w[2]), tc::tf32_hi(w[3]));
        *reinterpret_cast<float4 *>(W1B + (size_t)e * 16) = h;
        *reinterpret_cast<float4 *>(W1B + TC_W1_HALF + (size_t)e * 16) = make_float4(w[0] - h.x, w[1] - h.y, w[2] - h.z, w[3] - h.w);
    }
    // stage C: rows n < 112, K = hidden index, 24 chunks
    for (int e = tid; e < 24 * 112; e += nthr) {
        const int c = e / 112, n = e - c * 112;
        float w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) w[q] = W2T[(4 * c + q) * ZD + n];
        const float4 h = make_float4(tc::tf32_hi(w[0]), tc::tf32_hi(w[1]), tc::tf32_hi(w[2]), tc::tf32_hi(w[3]));
        *reinterpret_cast<float4 *>(W2B + (size_t)e * 16) = h;
        *reinterpret_cast<float4 *>(W2B + TC_W2_HALF + (size_t)e * 16) = make_float4(w[0] - h.x, w[1] - h.y, w[2] - h.z, w[3] - h.w);
    }
}

__device__ __forceinline__ void tc_store_split(uint8_t *sm, int chunk, int row, const float *v4) {
    const float4 h = make_float4(tc::tf32_hi(v4[0]), tc::tf32_hi(v4[1]), tc::tf32_hi(v4[2]), tc::tf32_hi(v4[3]));
    const uint32_t off = (uint32_t)chunk * TC_CHUNK + (uint32_t)row * 16u;
    *reinterpret_cast<float4 *>(sm + off) = h;
    *reinterpret_cast<float4 *>(sm + TC_AREG + off) = make_float4(v4[0] - h.x, v4[1] - h.y, v4[2] - h.z, v4[3] - h.w);
}

__global__ void __launch_bounds__(TC_THREADS, 1)
dec_tc_fwd_kernel(int V, int DP, const float *__restrict__ XT, const uint8_t *__restrict__ BA,
                  const uint8_t *__restrict__ W1B, const uint8_t *__restrict__ W2B, const float *__restrict__ bgeo,
                  const float *__restrict__ b1e, const float *__restrict__ b2, float *__restrict__ XIN,
                  float *__restrict__ H, float *__restrict__ Z, float *__restrict__ neural_opacity,
                  uint8_t *__restrict__ mask_out, uint32_t *__restrict__ maskbits, uint32_t *__restrict__ block_sums) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t barL, barM;
    __shared__ uint32_t tmem_s;
    __shared__ float s_bias[64 + HD + ZD];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r = tid & (TC_ROWS - 1), grp = tid >> 7;       // row within the tile (= TMEM lane), column group 0..3
    if (warp == 0) tc::tmem_alloc<128>(&tmem_s);
    if (tid == 0) { tc::mbar_init(&barL, 1); tc::mbar_init(&barM, 1); tc::fence_barrier_init(); }
    for (int i = tid; i < 64 + HD + ZD; i += TC_THREADS) s_bias[i] = i < 64 ? bgeo[i] : (i < 64 + HD ? b1e[i - 64] : b2[i - 64 - HD]);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_s;
    const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const uint32_t a_hi = tc::smem_u32(sm), a_lo = a_hi + TC_AREG;
    const uint32_t b_hi = a_hi + TC_OFF_B;
    const uint32_t w2_hi = a_hi + TC_OFF_W2;
    constexpr uint32_t idesc32 = tc::make_idesc_tf32(128, 32), idesc96 = tc::make_idesc_tf32(128, HD),
                       idesc112 = tc::make_idesc_tf32(128, ZD);
    uint32_t phL = 0, phM = 0;
    const int npc = (DP + 3) >> 2, nch = tc_tile_chunks(DP);
    const int ntiles = (V + TC_ROWS - 1) / TC_ROWS;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int row = tile * TC_ROWS + r;
        const bool valid = row < V;
        // ---- operands of stage A ----------------------------------------------------------------------
        if (tid == 0) {
            tc::mbar_arrive_expect_tx(&barL, 2 * TC_BA_HALF);
            tc::bulk_g2s(sm + TC_OFF_B, BA, 2 * TC_BA_HALF, &barL);
        }
        const float4 *xt = reinterpret_cast<const float4 *>(XT) + (size_t)tile * nch * TC_ROWS;
        // column group g takes chunks g, g + 4, ...: all (at most 9) 16-byte loads of a thread are independent and in
        // flight together (one dependent load per chunk made this the kernel's critical path)
        {
            constexpr int XK = (TC_NP + 19 + 3) / 4;        // 9
            float4 x[XK];
#pragma unroll
            for (int k = 0; k < XK; ++k)
                if (grp + 4 * k < nch) x[k] = __ldg(xt + (grp + 4 * k) * TC_ROWS + r);
#pragma unroll
            for (int k = 0; k < XK; ++k) {
                const int c = grp + 4 * k;
                if (c < nch) {
                    const int sc = c < npc ? 1 + c : (c < npc + 18 ? 18 + (c - npc) : 0);
                    const float v4[4] = {x[k].x, x[k].y, x[k].z, x[k].w};
                    tc_store_split(sm, sc, r, v4);
                }
            }
        }
        for (int pc = npc + grp; pc < TC_NP + 1; pc += 4) { // zero the P padding chunks and chunk 17
            const uint32_t off = (uint32_t)(1 + pc) * TC_CHUNK + (uint32_t)r * 16u;
            *reinterpret_cast<float4 *>(sm + off) = zero4;
            *reinterpret_cast<float4 *>(sm + TC_AREG + off) = zero4;
        }
        tc::fence_proxy_async();
        tc::tc_fence_before();
        __syncthreads();
        tc::tc_fence_after();
        tc::mbar_wait(&barL, phL); phL ^= 1;
        if (tid == 0) {
            tc::issue_3xtf32(tmem, a_hi, a_lo, TC_ROWS, 1, b_hi, b_hi + TC_BA_HALF, 32, 0, TC_NP / 2, idesc32, false);
            tc::issue_3xtf32(tmem + 32, a_hi, a_lo, TC_ROWS, 18, b_hi, b_hi + TC_BA_HALF, 32, 16, 9, idesc32, false);
            tc::mma_commit(&barM);
        }
        tc::mbar_wait(&barM, phM); phM ^= 1;
        tc::tc_fence_after();
        if (tid == 0) {                                     // stage-A weights are dead: fetch W1
            tc::mbar_arrive_expect_tx(&barL, 2 * TC_W1_HALF);
            tc::bulk_g2s(sm + TC_OFF_B, W1B, 2 * TC_W1_HALF, &barL);
        }
        // ---- epilogue A: geo -> x100 columns 36..99 (global, for the backward) and stage-B operand chunks 1..16
#pragma unroll
        for (int n0 = grp * 16; n0 < grp * 16 + 16; n0 += 8) {
            float v[8];
            tc::tmem_ld8(tlane + n0, v);
            tc::tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] += s_bias[n0 + q];
            if (valid) {
                float4 *dst = reinterpret_cast<float4 *>(XIN + (size_t)row * XI + 36 + n0);
                dst[0] = make_float4(v[0], v[1], v[2], v[3]);
                dst[1] = make_float4(v[4], v[5], v[6], v[7]);
            }
            tc_store_split(sm, 1 + n0 / 4, r, v);
            tc_store_split(sm, 2 + n0 / 4, r, v + 4);
        }
        tc::fence_proxy_async();
        tc::tc_fence_before();
        __syncthreads();
        tc::tc_fence_after();
        tc::mbar_wait(&barL, phL); phL ^= 1;
        if (tid == 0) {
            tc::issue_3xtf32(tmem, a_hi, a_lo, TC_ROWS, 18, b_hi, b_hi + TC_W1_HALF, HD, 0, 4, idesc96, false);   // feat
            tc::issue_3xtf32(tmem, a_hi, a_lo, TC_ROWS, 0, b_hi, b_hi + TC_W1_HALF, HD, 8, 9, idesc96, true);     // dir,dist | geo | pad
            tc::mma_commit(&barM);
        }
        tc::mbar_wait(&barM, phM); phM ^= 1;
        tc::tc_fence_after();
        if (tid == 0) {                                     // every stage-B operand is dead: fetch W2
            tc::mbar_arrive_expect_tx(&barL, 2 * TC_W2_HALF);
            tc::bulk_g2s(sm + TC_OFF_W2, W2B, 2 * TC_W2_HALF, &barL);
        }
        // ---- epilogue B: H = relu(. + b1) -> global (for the backward) and stage-C operand chunks 0..23
#pragma unroll
        for (int n0 = grp * (HD / 4); n0 < (grp + 1) * (HD / 4); n0 += 8) {
            float v[8];
            tc::tmem_ld8(tlane + n0, v);
            tc::tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 8; ++q) v[q] = fmaxf(v[q] + s_bias[64 + n0 + q], 0.f);
            if (valid) {
                float4 *dst = reinterpret_cast<float4 *>(H + (size_t)row * HD + n0);
                dst[0] = make_float4(v[0], v[1], v[2], v[3]);
                dst[1] = make_float4(v[4], v[5], v[6], v[7]);
            }
            tc_store_split(sm, n0 / 4, r, v);
            tc_store_split(sm, n0 / 4 + 1, r, v + 4);
        }
        tc::fence_proxy_async();
        tc::tc_fence_before();
        __syncthreads();
        tc::tc_fence_after();
        tc::mbar_wait(&barL, phL); phL ^= 1;
        if (tid == 0) {
            tc::issue_3xtf32(tmem, a_hi, a_lo, TC_ROWS, 0, w2_hi, w2_hi + TC_W2_HALF, ZD, 0, HD / 8, idesc112, false);
            tc::mma_commit(&barM);
        }
        tc::mbar_wait(&barM, phM); phM ^= 1;
        tc::tc_fence_after();
        // ---- epilogue C: activations, mask bits, survivor counts -------------------------------------------
        uint32_t bits = 0;
        // 14 groups of 8 columns dealt 4 / 4 / 3 / 3; the opacity columns (j < KO) all belong to column group 0
        const int c_beg = grp < 2 ? grp * 32 : 64 + (grp - 2) * 24, c_end = c_beg + (grp < 2 ? 32 : 24);
        static_assert(ZD == 112 && KO <= 32 && HD % 32 == 0, "column-group split of the decode epilogues");
#pragma unroll 1
        for (int n0 = c_beg; n0 < c_end; n0 += 8) {
            float v[8];
            tc::tmem_ld8(tlane + n0, v);
            tc::tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int j = n0 + q;
                float z = v[q] + s_bias[64 + HD + j];
                if (j < KO) {
                    z = tanhf(z);
                    if (valid) {
                        neural_opacity[(size_t)row * KO + j] = z;
                        mask_out[(size_t)row * KO + j] = z > 0.f ? 1 : 0;
                    }
                    bits |= z > 0.f ? (1u << j) : 0u;
                } else if (j >= 8 * KO && j < 11 * KO) {
                    z = 1.f / (1.f + expf(-z));
                }
                v[q] = z;
            }
            if (valid) {
                float4 *dst = reinterpret_cast<float4 *>(Z + (size_t)row * ZD + n0);
                dst[0] = make_float4(v[0], v[1], v[2], v[3]);
                dst[1] = make_float4(v[4], v[5], v[6], v[7]);
            }
        }
        if (grp == 0) {                                     // warps 0..3 hold the mask bits of their 32 rows
            if (!valid) bits = 0;
            if (valid) maskbits[row] = bits;
            uint32_t cnt = __popc(bits);
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
            if (lane == 0 && cnt) atomicAdd(&block_sums[(tile * TC_ROWS + warp * 32) / 256], cnt);
        }
        tc::tc_fence_before();
        __syncthreads();                                    // TMEM and shared memory are free for the next tile
        tc::tc_fence_after();
    }
    if (warp == 0) tc::tmem_dealloc<128>(tmem);
}

}  // namespace splatco
