// Decode v2 backward MLP kernel (included at the end of decode.cu; see decode2.cuh for the algebra and the tile layouts).
//
// Persistent, one CTA per SM, 128 anchors per tile, 16 worker warps + 1 control warp (bulk copies, tcgen05.mma).
// Per tile:
//   P0   un-compaction: upstream gradients of the compacted Gaussians -> dZ rows (pre-activation), written straight
//        into the K-major operand tile; the direct anchor / offset / scaling gradients go to a small global tile (DGA).
//        One thread per (anchor row, kind of output: opacity+position | colour | scale | rotation), the ten offsets
//        unrolled so that every column index is a compile-time constant.  The tile's Gaussians are one contiguous
//        range of the compacted outputs: their upstream gradients are staged into shared memory with coalesced
//        4-byte cp.async copies (read straight from global, every (anchor, offset) touched its own sectors and the
//        sector traffic, not the arithmetic, set the pace: 15 us per tile in the phase trace)
//   b1   dH = (dZ W2^T) .* [H > 0]        three block products (K = 16 | 72 | 32), accumulator in TMEM
//   w2   gW2[j][i] += sum_v dZ[v][j] H[v][i]       reduction over the tile's anchors; a row of ones appended to H^T makes
//        column 96 of the result the column sums of dZ (the output-bias gradients) for free
//   b2   dU = dH Wc1^T                              (N = NB <= 144)  -> DUT (+ the direct gradients)
//   g    GT[n][uc]  += sum_v dH[v][n] u[v][uc]
// The two weight-gradient products reduce over anchors, i.e. over the ROWS of the activation tiles.  kind::tf32 reads
// no-swizzle tiles only K-major (tools/tc_mn_probe.py: with a majorness bit set the instruction produces nothing), so
// their operands are transposed copies written by the threads, 32 anchors (4 K steps) at a time: chunk = 4 anchors,
// chunk stride = (rows + 1) * 16 B, which makes the 4-byte transposed stores of a warp hit 32 different banks.
// Both products accumulate in TMEM over all tiles of the CTA and leave it once, as a per-CTA partial (dec2_reduce_kernel).
//
// Shared memory (bytes), regions alias across the phases of a tile:
//   P0/b1/w2: dZ rows hi [0,61440) lo [61440,122880) | W2R [122880,153600) | upstream gradient rows of the tile
//             [153600,225280) (P0 only), then the w2 quarter tiles from 122880
//   b2/g    : dH rows hi [0,49152) lo [49152,98304) | W1R [98304,208896), then the g quarter tiles from 98304
// TMEM: dH acc [0,96) | dU acc [96,240) | gW2 acc [240,352) | GT acc [352,496).
#pragma once
#include <type_traits>

namespace splatco {

constexpr uint32_t D2B_DZLO = D2_RCH * D2_CHUNK;                     // 61440
constexpr uint32_t D2B_W2R = 2 * D2B_DZLO;                           // 122880
constexpr uint32_t D2B_UST = D2B_W2R + 2 * D2_W2R_HALF;              // 153600: the tile's upstream gradient rows, staged for the un-compaction
constexpr int D2B_UMAX = D2_ROWS * KO;                               // Gaussians of a tile: [opacity n | xyz 3n | colour 3n | scale 3n | rot 4n]
constexpr int D2B_U_XYZ = D2B_UMAX, D2B_U_COL = 4 * D2B_UMAX, D2B_U_SCL = 7 * D2B_UMAX, D2B_U_ROT = 10 * D2B_UMAX;
constexpr uint32_t D2B_DHLO = 24 * D2_CHUNK;                         // 49152
constexpr uint32_t D2B_W1R = 2 * D2B_DHLO;                           // 98304
constexpr uint32_t D2B_SMEM = 225280;                                // upstream staging ends here; the two g quarter buffers at 222720
// chunk strides of the transposed quarter tiles ((rows + 1) * 16 B).  The 96-row tiles (H^T, dH^T) are read with M or N = 96
// .. 128: an M = 128 product reads rows 96..127 of a chunk from the next chunk's first rows -- finite data that only
// reaches accumulator rows nobody reads.
constexpr uint32_t D2B_LA = 129 * 16, D2B_LH = 97 * 16, D2B_LHW = 113 * 16, D2B_LU = 145 * 16;
// w2 quarter: dZ^T hi | lo | H^T hi | lo (single buffer)     g quarter: dH^T hi | lo | u^T hi | lo (two buffers)
constexpr uint32_t D2B_QW_A = D2B_W2R, D2B_QW_B = D2B_QW_A + 2 * 8 * D2B_LA;
constexpr uint32_t D2B_QG_BYTES = 2 * 8 * D2B_LH + 2 * 8 * D2B_LU;   // 61952
constexpr uint32_t D2B_QG_A = D2B_W1R, D2B_QG_B = D2B_QG_A + 2 * 8 * D2B_LH;
static_assert(D2B_QW_B + 2 * 8 * D2B_LHW <= D2B_SMEM && D2B_W1R + 2 * 24 * 144 * 16 <= D2B_SMEM, "shared memory map");
static_assert(D2B_QG_A + 2 * D2B_QG_BYTES + 32 * 16 <= D2B_SMEM && D2B_UST + 14 * D2B_UMAX * 4 <= D2B_SMEM, "shared memory map");

struct D2Bwd {
    int V, M, nch, nk, NB, ntiles, trace;
    const float4 *XT, *HT, *ZT;
    const uint32_t *maskbits, *offs;
    const float *d_xyz, *d_color, *d_opacity, *d_scaling, *d_rot, *d_nopac;
    const uint8_t *W2R, *W1R;
    float4 *DUT, *DGA;
    float *part;
};

__device__ __forceinline__ void d2_issue_lbo(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t lboA, uint32_t b_hi,
                                             uint32_t b_lo, uint32_t lboB, int ksteps, uint32_t idesc, bool accumulate_first) {
    tc::issue_3xtf32_lbo(d_tmem, a_hi, a_lo, lboA, 0, b_hi, b_lo, lboB, 0, ksteps, idesc, accumulate_first);
}

// transposed hi/lo store of one 16-byte cell (4 consecutive columns m0.. of anchor row rr) into a quarter tile
__device__ __forceinline__ void d2_put_t(uint8_t *sm, uint32_t base_hi, uint32_t base_lo, uint32_t lbo, int m0, int rr, const float4 &h,
                                         const float4 &l) {
    const uint32_t o = (uint32_t)((rr & 31) >> 2) * lbo + (uint32_t)m0 * 16u + (uint32_t)(rr & 3) * 4u;
    *reinterpret_cast<float *>(sm + base_hi + o) = h.x; *reinterpret_cast<float *>(sm + base_hi + o + 16) = h.y;
    *reinterpret_cast<float *>(sm + base_hi + o + 32) = h.z; *reinterpret_cast<float *>(sm + base_hi + o + 48) = h.w;
    *reinterpret_cast<float *>(sm + base_lo + o) = l.x; *reinterpret_cast<float *>(sm + base_lo + o + 16) = l.y;
    *reinterpret_cast<float *>(sm + base_lo + o + 32) = l.z; *reinterpret_cast<float *>(sm + base_lo + o + 48) = l.w;
}
__device__ __forceinline__ void d2_split4(const float4 &x, float4 &h, float4 &l) {
    h = make_float4(tc::tf32_hi(x.x), tc::tf32_hi(x.y), tc::tf32_hi(x.z), tc::tf32_hi(x.w));
    l = make_float4(x.x - h.x, x.y - h.y, x.z - h.z, x.w - h.w);
}

__global__ void __launch_bounds__(D2_THREADS, 1)
dec2_mlp_bwd_kernel(D2Bwd a) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t barWa, barWb, barB1, barB2, barQ, barG[2];
    __shared__ float s_ex[D2_ROWS][3];                       // un-compaction: scale-sum partials handed from column group 2 to 3
    __shared__ uint32_t tmem_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tc::tmem_alloc<512>(&tmem_s);
    if (tid == 0) {
        tc::mbar_init(&barWa, 1); tc::mbar_init(&barWb, 1); tc::mbar_init(&barB1, 1); tc::mbar_init(&barB2, 1); tc::mbar_init(&barQ, 1);
        tc::mbar_init(&barG[0], 1); tc::mbar_init(&barG[1], 1);
        tc::fence_barrier_init();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_s;
    const uint32_t sb = tc::smem_u32(sm);
    const int ntl = (a.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const uint32_t w1r_half = 24u * (uint32_t)a.NB * 16u;
    constexpr uint32_t T_DH = 0, T_DU = 96, T_GW2 = 240, T_GT = 352;
    uint32_t qph = 0, gph[2] = {0, 0};                       // phases of barQ / barG[] this thread has waited for

    if (warp == D2_WORKERS / 32) {
        // =========================== control warp ===========================
        constexpr uint32_t id32 = tc::make_idesc_tf32(128, 32), id112 = tc::make_idesc_tf32(128, D2_GW2_LD), id144 = tc::make_idesc_tf32(128, 144);
        const uint32_t idNB = tc::make_idesc_tf32(128, a.NB);
        for (int it = 0; it < ntl; ++it) {
            const uint32_t par = it & 1;
            if (it > 0) {                                    // the last two g products of the previous tile still read their tiles
                tc::mbar_wait(&barG[0], gph[0] & 1); ++gph[0];
                tc::mbar_wait(&barG[1], gph[1] & 1); ++gph[1];
            }
            if (lane == 0) {
                tc::mbar_arrive_expect_tx(&barWa, 2 * D2_W2R_HALF);
                tc::bulk_g2s(sm + D2B_W2R, a.W2R, 2 * D2_W2R_HALF, &barWa);
            }
            d2_bar_sync_all();                               // 1: dZ rows written
            tc::tc_fence_after();
            tc::mbar_wait(&barWa, par);
            if (lane == 0) {
                const uint32_t b = sb + D2B_W2R, bl = b + D2_W2R_HALF;
                tc::issue_3xtf32(tmem + T_DH, sb, sb + D2B_DZLO, D2_ROWS, 0, b, bl, 32, 0, 2, id32, false);
                tc::issue_3xtf32(tmem + T_DH + 32, sb, sb + D2B_DZLO, D2_ROWS, 4, b + 4 * 32 * 16, bl + 4 * 32 * 16, 32, 0, 9, id32, false);
                tc::issue_3xtf32(tmem + T_DH + 64, sb, sb + D2B_DZLO, D2_ROWS, 22, b + 22 * 32 * 16, bl + 22 * 32 * 16, 32, 0, 4, id32, false);
                tc::mma_commit(&barB1);
            }
            __syncwarp();
            tc::mbar_wait(&barB1, par);                      // W2R is dead: the w2 quarter tiles may land on it
            for (int q = 0; q < 4; ++q) {
                d2_bar_sync_all();                           // 2..5: w2 quarter operands written
                tc::tc_fence_after();
                if (lane == 0) {
                    d2_issue_lbo(tmem + T_GW2, sb + D2B_QW_A, sb + D2B_QW_A + 8 * D2B_LA, D2B_LA, sb + D2B_QW_B, sb + D2B_QW_B + 8 * D2B_LHW,
                                 D2B_LHW, 4, id112, it > 0 || q > 0);
                    tc::mma_commit(&barQ);
                }
                __syncwarp();
                tc::mbar_wait(&barQ, qph & 1); ++qph;
            }
            if (lane == 0) {
                tc::mbar_arrive_expect_tx(&barWb, 2 * w1r_half);
                tc::bulk_g2s(sm + D2B_W1R, a.W1R, 2 * w1r_half, &barWb);
            }
            d2_bar_sync_all();                               // 6: dH rows written
            tc::tc_fence_after();
            tc::mbar_wait(&barWb, par);
            if (lane == 0) {
                tc::issue_3xtf32(tmem + T_DU, sb, sb + D2B_DHLO, D2_ROWS, 0, sb + D2B_W1R, sb + D2B_W1R + w1r_half, a.NB, 0, HD / 8, idNB, false);
                tc::mma_commit(&barB2);
            }
            __syncwarp();
            tc::mbar_wait(&barB2, par);
            for (int q = 0; q < 4; ++q) {
                d2_bar_sync_all();                           // 7..10: g quarter operands written (buffer q & 1)
                tc::tc_fence_after();
                if (q >= 2) { tc::mbar_wait(&barG[q & 1], gph[q & 1] & 1); ++gph[q & 1]; }   // (keeps this thread's phase count in step)
                if (lane == 0) {
                    const uint32_t qa = sb + D2B_QG_A + (q & 1) * D2B_QG_BYTES, qb = sb + D2B_QG_B + (q & 1) * D2B_QG_BYTES;
                    d2_issue_lbo(tmem + T_GT, qa, qa + 8 * D2B_LH, D2B_LH, qb, qb + 8 * D2B_LU, D2B_LU, 4, id144, it > 0 || q > 0);
                    tc::mma_commit(&barG[q & 1]);
                }
                __syncwarp();
            }
        }
    } else {
        // =========================== workers ===========================
        const int r = tid & (D2_ROWS - 1), grp = tid >> 7;
        const uint32_t tlane = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        const int ng = a.NB / 8, g_beg = (ng * grp) / 4, g_end = (ng * (grp + 1)) / 4;      // dU column groups of this thread
        for (int it = 0; it < ntl; ++it) {
            const int tile = blockIdx.x + it * gridDim.x;
            const uint32_t par = it & 1;
            D2_TRACE(1, 16 * it + 0);
            // ---- P0: un-compaction (gaussian_renderer/__init__.py:96-111 backwards) --------------------------------------
            float scl_part[3] = {0.f, 0.f, 0.f};             // column group 3: its half of the scale sums, finished after sync 1
            {
                const int row0 = tile * D2_ROWS, v = row0 + r;
                const bool valid = v < a.V;
                const uint32_t bits = valid ? __ldg(a.maskbits + v) : 0u;
                const uint32_t off0 = valid ? __ldg(a.offs + v) : 0u;
                // the tile's Gaussians: rows [j0, j1) of the compacted outputs
                const uint32_t j0 = __ldg(a.offs + row0);
                const uint32_t j1 = row0 + D2_ROWS < a.V ? __ldg(a.offs + row0 + D2_ROWS) : (uint32_t)a.M;
                const int n = (int)(j1 - j0);
                const float4 *ztile = a.ZT + (size_t)tile * D2_ZCH * D2_ROWS + r;
                const float4 *gtile = a.XT + (size_t)tile * a.nch * D2_ROWS + r;
                float *ust = reinterpret_cast<float *>(sm + D2B_UST);
                D2_TRACE(1, 16 * it + 1);
                // the last two g products of the previous tile have completed (the control warp waited before its bulk
                // copy, these threads wait here): the staging area and the dZ rows may be overwritten
                if (it > 0) {
                    tc::mbar_wait(&barG[0], gph[0] & 1); ++gph[0];
                    tc::mbar_wait(&barG[1], gph[1] & 1); ++gph[1];
                }
                auto stage = [&](int dst, const float *src, int count) {
                    for (int e = tid; e < count; e += D2_WORKERS)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sb + D2B_UST + 4u * (uint32_t)(dst + e)), "l"(src + e) : "memory");
                };
                if (n > 0) {
                    stage(0, a.d_opacity + j0, n);
                    stage(D2B_U_XYZ, a.d_xyz + 3 * (size_t)j0, 3 * n);
                    stage(D2B_U_COL, a.d_color + 3 * (size_t)j0, 3 * n);
                    stage(D2B_U_SCL, a.d_scaling + 3 * (size_t)j0, 3 * n);
                    stage(D2B_U_ROT, a.d_rot + 4 * (size_t)j0, 4 * n);
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
                auto putcell = [&](int cell, float x0, float x1, float x2, float x3) {
                    float4 h, l;
                    d2_split4(make_float4(x0, x1, x2, x3), h, l);
                    const uint32_t o = (uint32_t)cell * D2_CHUNK + (uint32_t)r * 16u;
                    *reinterpret_cast<float4 *>(sm + o) = h;
                    *reinterpret_cast<float4 *>(sm + D2B_DZLO + o) = l;
                };
                auto putz = [&](int col, float x) {
                    const float h = tc::tf32_hi(x);
                    const uint32_t o = (uint32_t)(col >> 2) * D2_CHUNK + (uint32_t)r * 16u + (uint32_t)(col & 3) * 4u;
                    *reinterpret_cast<float *>(sm + o) = h;
                    *reinterpret_cast<float *>(sm + D2B_DZLO + o) = x - h;
                };
                // local index (in the staging area) of offset k's Gaussian; meaningful when bit k is set
                auto jl = [&](int k) -> int { return (int)(off0 - j0) + __popc(bits & ((1u << k) - 1u)); };
                auto staged = [&]() {
                    asm volatile("cp.async.wait_group 0;" ::: "memory");
                    asm volatile("bar.sync 1, %0;" ::"n"(D2_WORKERS) : "memory");
                };
                if (grp == 0) {
                    // opacity column and the position path: xyz = anchor + offset * scaling[:3]
                    float no[12], ga[36];
#pragma unroll
                    for (int c = 0; c < 3; ++c) { const float4 t = __ldg(ztile + c * D2_ROWS); no[4 * c] = t.x; no[4 * c + 1] = t.y; no[4 * c + 2] = t.z; no[4 * c + 3] = t.w; }
#pragma unroll
                    for (int c = 0; c < 9; ++c) { const float4 t = __ldg(gtile + (8 + c) * D2_ROWS); ga[4 * c] = t.x; ga[4 * c + 1] = t.y; ga[4 * c + 2] = t.z; ga[4 * c + 3] = t.w; }
                    float dnp[KO];
#pragma unroll
                    for (int k = 0; k < KO; ++k) dnp[k] = (valid && a.d_nopac) ? __ldg(a.d_nopac + (size_t)v * KO + k) : 0.f;
                    staged();
                    float *dga = reinterpret_cast<float *>(a.DGA + (size_t)tile * 10 * D2_ROWS + r);
                    float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, dzo[12];
#pragma unroll
                    for (int k = 0; k < KO; ++k) {
                        const bool m = (bits >> k) & 1u;
                        const int j = m ? jl(k) : 0;
                        const float dno = dnp[k] + (m ? ust[j] : 0.f);
                        dzo[k] = valid ? dno * (1.f - no[k] * no[k]) : 0.f;
#pragma unroll
                        for (int q = 0; q < 3; ++q) {
                            const float gq = m ? ust[D2B_U_XYZ + 3 * j + q] : 0.f;
                            dga[d2_tile_idx(3 + 3 * k + q)] = gq * ga[33 + q];            // ga: u columns 32..67; scaling at 65..67
                            acc[q] += gq;
                            acc[3 + q] = fmaf(gq, ga[3 + 3 * k + q], acc[3 + q]);         // offsets at u columns 35..64
                        }
                    }
                    dzo[10] = 0.f; dzo[11] = 0.f;
#pragma unroll
                    for (int q = 0; q < 3; ++q) { dga[d2_tile_idx(q)] = acc[q]; dga[d2_tile_idx(33 + q)] = acc[3 + q]; }
                    putcell(0, dzo[0], dzo[1], dzo[2], dzo[3]); putcell(1, dzo[4], dzo[5], dzo[6], dzo[7]);
                    putcell(2, dzo[8], dzo[9], 0.f, 0.f); putcell(3, 0.f, 0.f, 0.f, 0.f);
                } else if (grp == 1) {
                    float zc[32];                            // colour block: Z columns 96..127
#pragma unroll
                    for (int c = 0; c < 8; ++c) { const float4 t = __ldg(ztile + (24 + c) * D2_ROWS); zc[4 * c] = t.x; zc[4 * c + 1] = t.y; zc[4 * c + 2] = t.z; zc[4 * c + 3] = t.w; }
                    staged();
                    float dz[32];
#pragma unroll
                    for (int k = 0; k < KO; ++k) {
                        const bool m = (bits >> k) & 1u;
                        const int j = m ? jl(k) : 0;
#pragma unroll
                        for (int q = 0; q < 3; ++q) {
                            const float c = zc[3 * k + q];
                            dz[3 * k + q] = m ? ust[D2B_U_COL + 3 * j + q] * c * (1.f - c) : 0.f;
                        }
                    }
                    dz[30] = 0.f; dz[31] = 0.f;
#pragma unroll
                    for (int c = 0; c < 8; ++c) putcell(D2_RCOL / 4 + c, dz[4 * c], dz[4 * c + 1], dz[4 * c + 2], dz[4 * c + 3]);
                } else {
                    // covariance block: column group 2 takes offsets 0..4 (Z columns 16..50), group 3 offsets 5..9 (51..85).
                    // (generic lambda: the two halves are separate instantiations, so every index below is a constant)
                    auto cov = [&](auto KB, auto CELL0) {
                        constexpr int kb = decltype(KB)::value;
                        constexpr int cell0 = decltype(CELL0)::value;      // first Z cell this thread loads (columns 16.. / 48..)
                        float zc[40];
#pragma unroll
                        for (int c = 0; c < 10; ++c) { const float4 t = __ldg(ztile + (cell0 + c) * D2_ROWS); zc[4 * c] = t.x; zc[4 * c + 1] = t.y; zc[4 * c + 2] = t.z; zc[4 * c + 3] = t.w; }
                        const float4 sc = __ldg(gtile + 17 * D2_ROWS);      // u columns 68..71: scaling[3..5], 1
                        const float s3[3] = {sc.x, sc.y, sc.z};
                        staged();
                        float dz[35];
#pragma unroll
                        for (int kk = 0; kk < 5; ++kk) {
                            const int k = kb + kk;
                            const bool m = (bits >> k) & 1u;
                            const int j = m ? jl(k) : 0;
                            // Z column of (k, q) = 16 + 7 k + q; this thread's zc[] starts at column 4 * cell0
                            const int zb = D2_ZCOV + 7 * k - 4 * cell0;
#pragma unroll
                            for (int q = 0; q < 3; ++q) {
                                const float sg = 1.f / (1.f + expf(-zc[zb + q]));
                                const float gq = m ? ust[D2B_U_SCL + 3 * j + q] : 0.f;
                                dz[7 * kk + q] = gq * s3[q] * sg * (1.f - sg);
                                scl_part[q] = fmaf(gq, sg, scl_part[q]);
                            }
                            const float4 gr = m ? *reinterpret_cast<const float4 *>(ust + D2B_U_ROT + 4 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
                            const float q0 = zc[zb + 3], q1 = zc[zb + 4], q2 = zc[zb + 5], q3 = zc[zb + 6];
                            const float nrm = sqrtf(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
                            const float inv = 1.f / fmaxf(nrm, 1e-12f);
                            const float r0 = q0 * inv, r1 = q1 * inv, r2 = q2 * inv, r3 = q3 * inv;
                            const float dot = nrm > 1e-12f ? r0 * gr.x + r1 * gr.y + r2 * gr.z + r3 * gr.w : 0.f;
                            dz[7 * kk + 3] = (gr.x - r0 * dot) * inv; dz[7 * kk + 4] = (gr.y - r1 * dot) * inv;
                            dz[7 * kk + 5] = (gr.z - r2 * dot) * inv; dz[7 * kk + 6] = (gr.w - r3 * dot) * inv;
                        }
                        if constexpr (kb == 0) {             // dZ columns 16..50: cells 4..11, then 48, 49, 50
#pragma unroll
                            for (int c = 0; c < 8; ++c) putcell(4 + c, dz[4 * c], dz[4 * c + 1], dz[4 * c + 2], dz[4 * c + 3]);
                            putz(48, dz[32]); putz(49, dz[33]); putz(50, dz[34]);
                            s_ex[r][0] = scl_part[0]; s_ex[r][1] = scl_part[1]; s_ex[r][2] = scl_part[2];
                        } else {                             // dZ columns 51..85 (+ padding 86, 87): 51, then cells 13..21
                            putz(51, dz[0]);
#pragma unroll
                            for (int c = 0; c < 8; ++c) putcell(13 + c, dz[1 + 4 * c], dz[2 + 4 * c], dz[3 + 4 * c], dz[4 + 4 * c]);
                            putcell(21, dz[33], dz[34], 0.f, 0.f);
                        }
                    };
                    if (grp == 2) cov(std::integral_constant<int, 0>{}, std::integral_constant<int, 4>{});
                    else cov(std::integral_constant<int, 5>{}, std::integral_constant<int, 12>{});
                }
            }
            tc::fence_proxy_async();
            tc::tc_fence_before();
            D2_TRACE(1, 16 * it + 2);
            d2_bar_sync_all();                               // 1
            if (grp == 3)                                    // scale sums of all ten offsets -> DGA columns 36..38 (39: the constant's slot)
                a.DGA[((size_t)tile * 10 + 9) * D2_ROWS + r] = make_float4(scl_part[0] + s_ex[r][0], scl_part[1] + s_ex[r][1], scl_part[2] + s_ex[r][2], 0.f);
            // ---- gate bits of this thread's 24 hidden columns (for epilogue b1), loaded while b1 runs ---------------------
            uint32_t hbits = 0u;
            {
                const float4 *ht = a.HT + ((size_t)tile * 24 + 6 * grp) * D2_ROWS + r;
#pragma unroll
                for (int c = 0; c < 6; ++c) {
                    const float4 h = __ldg(ht + c * D2_ROWS);
                    hbits |= ((h.x > 0.f ? 1u : 0u) | (h.y > 0.f ? 2u : 0u) | (h.z > 0.f ? 4u : 0u) | (h.w > 0.f ? 8u : 0u)) << (4 * c);
                }
            }
            // ---- w2: gW2 += dZ^T H, 32 anchors at a time; the H cells of the next quarter are fetched while the current
            //      one multiplies ---------------------------------------------------------------------------------------------
            const float4 *ht_tile = a.HT + (size_t)tile * 24 * D2_ROWS;
            float4 hpre[2];
            auto load_h = [&](int q) {
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int e = tid + i * D2_WORKERS;
                    if (e < 24 * 32) hpre[i] = __ldg(ht_tile + (e >> 5) * D2_ROWS + 32 * q + (e & 31));
                }
            };
            load_h(0);
            tc::mbar_wait(&barB1, par);
            tc::tc_fence_after();
            D2_TRACE(1, 16 * it + 3);
            for (int q = 0; q < 4; ++q) {
                if (q > 0) { tc::mbar_wait(&barQ, qph & 1); ++qph; }
                D2_TRACE(1, 16 * it + 4 + q);
                for (int e = tid; e < D2_RCH * 32; e += D2_WORKERS) {
                    const int c = e >> 5, rr = 32 * q + (e & 31);
                    const uint32_t o = (uint32_t)c * D2_CHUNK + (uint32_t)rr * 16u;
                    d2_put_t(sm, D2B_QW_A, D2B_QW_A + 8 * D2B_LA, D2B_LA, 4 * c, rr, *reinterpret_cast<const float4 *>(sm + o),
                             *reinterpret_cast<const float4 *>(sm + D2B_DZLO + o));
                }
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int e = tid + i * D2_WORKERS;
                    if (e < 24 * 32) {
                        float4 h, l;
                        d2_split4(hpre[i], h, l);
                        d2_put_t(sm, D2B_QW_B, D2B_QW_B + 8 * D2B_LHW, D2B_LHW, 4 * (e >> 5), 32 * q + (e & 31), h, l);
                    }
                }
                if (tid < 16) {                              // row 96 of H^T: ones (hi) / zeros (lo) -> column 96 of gW2 = column sums of dZ
                    const float o1 = tid < 8 ? 1.f : 0.f;
                    *reinterpret_cast<float4 *>(sm + D2B_QW_B + (tid < 8 ? 0u : 8 * D2B_LHW) + (uint32_t)(tid & 7) * D2B_LHW + HD * 16) =
                        make_float4(o1, o1, o1, o1);
                }
                if (q < 3) load_h(q + 1);
                tc::fence_proxy_async();
                tc::tc_fence_before();
                d2_bar_sync_all();                           // 2..5
            }
            // ---- epilogue b1: dH = acc .* [H > 0] (registers) ---------------------------------------------------------------
            float dh[24];
#pragma unroll
            for (int n0 = 0; n0 < 24; n0 += 8) {
                float v[8];
                tc::tmem_ld8(tlane + T_DH + 24 * grp + n0, v);
                tc::tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 8; ++e) dh[n0 + e] = (hbits >> (n0 + e)) & 1u ? v[e] : 0.f;
            }
            D2_TRACE(1, 16 * it + 8);
            tc::mbar_wait(&barQ, qph & 1); ++qph;            // last w2 quarter done: the dZ rows and the quarter tiles are dead
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                float4 h, l;
                d2_split4(make_float4(dh[4 * c], dh[4 * c + 1], dh[4 * c + 2], dh[4 * c + 3]), h, l);
                const uint32_t o = (uint32_t)(6 * grp + c) * D2_CHUNK + (uint32_t)r * 16u;
                *reinterpret_cast<float4 *>(sm + o) = h;
                *reinterpret_cast<float4 *>(sm + D2B_DHLO + o) = l;
            }
            tc::fence_proxy_async();
            tc::tc_fence_before();
            D2_TRACE(1, 16 * it + 9);
            d2_bar_sync_all();                               // 6
            const float4 *xt_tile = a.XT + (size_t)tile * a.nch * D2_ROWS;
            float4 upre[3];
            auto load_u = [&](int q) {
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const int e = tid + i * D2_WORKERS;
                    if (e < a.nch * 32) upre[i] = __ldg(xt_tile + (e >> 5) * D2_ROWS + 32 * q + (e & 31));
                }
            };
            load_u(0);
            tc::mbar_wait(&barB2, par);
            tc::tc_fence_after();
            D2_TRACE(1, 16 * it + 10);
            // ---- epilogue b2: dU (+ direct gradients on the anchor / offset / scaling columns) -> DUT ------------------------
            {
                float4 *du = a.DUT + (size_t)tile * a.nch * D2_ROWS + r;
                const float4 *dg = a.DGA + (size_t)tile * 10 * D2_ROWS + r;      // written by this CTA in P0 (bar.sync in between)
                for (int gi = g_beg; gi < g_end; ++gi) {
                    float v[8];
                    tc::tmem_ld8(tlane + T_DU + 8 * gi, v);
                    tc::tmem_ld_wait();
                    // u columns 32..71 = [anchor | offsets | scaling | 1]: chunks 8..17 <-> DGA cells 0..9 (cell 9's last entry is 0)
                    if (2 * gi >= 8 && 2 * gi < 18) {
                        const float4 d = __ldcg(dg + (2 * gi - 8) * D2_ROWS);
                        v[0] += d.x; v[1] += d.y; v[2] += d.z; v[3] += d.w;
                    }
                    if (2 * gi + 1 >= 8 && 2 * gi + 1 < 18) {
                        const float4 d = __ldcg(dg + (2 * gi + 1 - 8) * D2_ROWS);
                        v[4] += d.x; v[5] += d.y; v[6] += d.z; v[7] += d.w;
                    }
                    if (2 * gi < a.nch) du[(2 * gi) * D2_ROWS] = make_float4(v[0], v[1], v[2], v[3]);
                    if (2 * gi + 1 < a.nch) du[(2 * gi + 1) * D2_ROWS] = make_float4(v[4], v[5], v[6], v[7]);
                }
            }
            // ---- g: GT += dH^T u, 32 anchors at a time, two tile buffers: quarter q + 1 is written while quarter q multiplies;
            //      the u cells are fetched one quarter ahead ---------------------------------------------------------------------
            for (int q = 0; q < 4; ++q) {
                const uint32_t qo = (q & 1) * D2B_QG_BYTES;
                D2_TRACE(1, 16 * it + 11 + q);
                if (q >= 2) { tc::mbar_wait(&barG[q & 1], gph[q & 1] & 1); ++gph[q & 1]; }
                for (int e = tid; e < 24 * 32; e += D2_WORKERS) {
                    const int c = e >> 5, rr = 32 * q + (e & 31);
                    const uint32_t o = (uint32_t)c * D2_CHUNK + (uint32_t)rr * 16u;
                    d2_put_t(sm, D2B_QG_A + qo, D2B_QG_A + qo + 8 * D2B_LH, D2B_LH, 4 * c, rr, *reinterpret_cast<const float4 *>(sm + o),
                             *reinterpret_cast<const float4 *>(sm + D2B_DHLO + o));
                }
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const int e = tid + i * D2_WORKERS;
                    if (e < a.nch * 32) {
                        float4 h, l;
                        d2_split4(upre[i], h, l);
                        d2_put_t(sm, D2B_QG_B + qo, D2B_QG_B + qo + 8 * D2B_LU, D2B_LU, 4 * (e >> 5), 32 * q + (e & 31), h, l);
                    }
                }
                if (q < 3) load_u(q + 1);
                tc::fence_proxy_async();
                tc::tc_fence_before();
                d2_bar_sync_all();                           // 7..10
            }
        }
        // ---- the CTA's weight-gradient accumulators leave TMEM once ---------------------------------------------------------------
        if (ntl > 0) {
            tc::mbar_wait(&barG[0], gph[0] & 1); ++gph[0];
            tc::mbar_wait(&barG[1], gph[1] & 1); ++gph[1];
            tc::tc_fence_after();
            float *pw = a.part + (size_t)blockIdx.x * D2_PART + (size_t)r * D2_GW2_LD;
            for (int gi = (14 * grp) / 4; gi < (14 * (grp + 1)) / 4; ++gi) {
                float v[8];
                tc::tmem_ld8(tlane + T_GW2 + 8 * gi, v);
                tc::tmem_ld_wait();
                *reinterpret_cast<float4 *>(pw + 8 * gi) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4 *>(pw + 8 * gi + 4) = make_float4(v[4], v[5], v[6], v[7]);
            }
            float *pg = a.part + (size_t)blockIdx.x * D2_PART + D2_ROWS * D2_GW2_LD + (size_t)r * 144;
            for (int gi = (18 * grp) / 4; gi < (18 * (grp + 1)) / 4; ++gi) {
                float v[8];
                tc::tmem_ld8(tlane + T_GT + 8 * gi, v);
                tc::tmem_ld_wait();
                *reinterpret_cast<float4 *>(pg + 8 * gi) = make_float4(v[0], v[1], v[2], v[3]);
                *reinterpret_cast<float4 *>(pg + 8 * gi + 4) = make_float4(v[4], v[5], v[6], v[7]);
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<512>(tmem);
}

}  // namespace splatco

namespace {

int v2_decode_bwd(const splatco_decode_desc *d, const void *fwd_ws, void *bwd_ws, int M, const float *d_xyz, const float *d_color,
                  const float *d_opacity, const float *d_scaling, const float *d_rot, const float *d_neural_opacity,
                  const splatco_decode_grads *g, void *stream) {
    if (d->V == 0) return 0;
    SPLATCO_REQUIRE(fwd_ws && bwd_ws && g, "decode_bwd: null pointer");
    SPLATCO_REQUIRE(M == 0 || (d_xyz && d_color && d_opacity && d_scaling && d_rot), "decode_bwd: null upstream gradient");
    SPLATCO_REQUIRE(g->anchor_feat && g->anchor && g->offset && g->scaling, "decode_bwd: null per-anchor gradient");
    SPLATCO_REQUIRE(!d->V_dev, "decode_bwd: the backward needs the exact V on the host (V_dev must be NULL)");
    SPLATCO_REQUIRE(d->V <= d2_layout_rows(d), "decode_bwd: V = %d exceeds V_layout = %d", d->V, d->V_layout);
    cudaStream_t st = (cudaStream_t)stream;
    const D2Dims dd = d2_dims(d2_layout_rows(d), d->rc, d->level);       // workspace layout (the forward's)
    const int V = d->V, ntiles = ceil_div(V, D2_ROWS);                     // exact row / tile counts
    F2View f = f2_view(const_cast<void *>(fwd_ws), dd);
    B2View b = b2_view(bwd_ws, dd);
    const DecPtrs p = make_ptrs(d);
    const DecWeights w = make_weights(d);
    DecWeightGrads gw;
    DecInputGrads gi;
    for (int l = 0; l < 3; ++l) {
        gw.bn_w[l] = g->bn_w[l]; gw.bn_b[l] = g->bn_b[l]; gw.lin_w[l] = g->lin_w[l]; gw.lin_b[l] = g->lin_b[l];
        gw.cbn_w[l] = g->cbn_w[l]; gw.cbn_b[l] = g->cbn_b[l]; gw.clin_w[l] = g->clin_w[l]; gw.clin_b[l] = g->clin_b[l];
        gw.w1[l] = g->w1[l]; gw.b1[l] = g->b1[l]; gw.w2[l] = g->w2[l]; gw.b2[l] = g->b2[l];
        for (int q = 0; q < 3; ++q) gi.plane[l][q] = g->plane[3 * l + q];
        gi.att[l] = g->att[l];
    }
    gw.app_vec = g->app_vec;
    gi.anchor_feat = g->anchor_feat; gi.anchor = g->anchor; gi.offset = g->offset; gi.scaling = g->scaling;

    static unsigned char attr_dev[64];
    const int attr_i = current_device() & 63;
    if (!attr_dev[attr_i]) {
        SPLATCO_CHECK_CUDA(cudaFuncSetAttribute(dec2_mlp_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)D2B_SMEM));
        attr_dev[attr_i] = 1;
    }
    D2Bwd a;
    a.V = V; a.M = M; a.nch = dd.nch; a.nk = dd.nk; a.NB = dd.NB; a.ntiles = ntiles;
    a.XT = f.XT; a.HT = f.HT; a.ZT = f.ZT; a.maskbits = f.maskbits; a.offs = f.offs;
    a.d_xyz = d_xyz; a.d_color = d_color; a.d_opacity = d_opacity; a.d_scaling = d_scaling; a.d_rot = d_rot; a.d_nopac = d_neural_opacity;
    a.W2R = f.W2R; a.W1R = f.W1R; a.DUT = b.DUT; a.DGA = b.DGA; a.part = b.part;
    a.trace = g_decode_profile == 2;
    const int ctas = min(ntiles, D2_MAX_CTAS);
    prof_record(2, st);
    dec2_mlp_bwd_kernel<<<ctas, D2_THREADS, D2B_SMEM, st>>>(a);
    SPLATCO_CHECK_LAUNCH();
    prof_record(3, st);
    dec2_reduce_kernel<<<ceil_div(D2_PART, 256), 256, 0, st>>>(ctas, b.part, b.red);
    SPLATCO_CHECK_LAUNCH();
    dec2_expand_kernel<<<48, 256, 0, st>>>(dd.DP, dd.LDX, b.red, f.WpT, f.WcT, f.bgeo, f.W1T, b.S1, b.S0, b.gW1T, b.gb1,
                                           b.gW2T, b.gb2);
    SPLATCO_CHECK_LAUNCH();
    dec_bwd_fold_kernel<<<dim3(BWD_FOLD_CTAS, BWD_FOLD_SECTIONS), 256, 0, st>>>(w, gw, V, dd.rc, dd.level, dd.DP, dd.LDX, f.mu, f.rstd, f.WpG, f.WcG, b.S1, b.S0,
                                                   b.gW1T, b.gb1, b.gW2T, b.gb2, b.m1, b.m2);
    SPLATCO_CHECK_LAUNCH();
    if (D2_DISPATCH(launch_inputs2, dd.level, dd.rc, d->plane_layout != 0, st, p, gi, dd, V, f.XT, b.DUT, f.mu, f.rstd, b.m1, b.m2)) return -2;
    return 0;
}

}  // namespace
