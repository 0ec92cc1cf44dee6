// tcgen05 / TMEM / mbarrier / bulk-copy primitives for sm_100a (inline PTX; no CUTLASS).
//
// Operand tiles live in shared memory in the canonical K-major, no-swizzle ("interleave") UMMA layout:
// a tile of R rows x K fp32 is stored as K/4 "chunks"; chunk c holds columns 4c..4c+3 of all R rows as
// R consecutive 16-byte cells, i.e.   byte_offset(r, k) = ((k >> 2) * R + r) * 16 + (k & 3) * 4.
// A core matrix (8 rows x 16 B) is therefore 128 contiguous bytes; core matrices adjacent along M/N are
// 128 B apart (SBO) and core matrices adjacent along K are R*16 B apart (LBO).  One kind::tf32
// instruction consumes K = 8 (two chunks).  Precision: fp32 operands are split x = hi + lo with hi
// exactly representable in TF32 (low 13 mantissa bits cleared) and three MMAs accumulate
// hi*hi + hi*lo + lo*hi in the fp32 TMEM accumulator ("3xTF32", error ~2^-21 per product).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace splatco {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ----------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

// ---- 1-D bulk copy global -> shared (TMA engine, no tensor map); size multiple of 16 B ---------------
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// generic-proxy smem writes -> visible to the async proxy (UMMA operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMEM --------------------------------------------------------------------------------------------
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem) {     // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(NCOLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {       // same warp that allocated
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS));
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 lanes x 8 consecutive columns: thread `lane` of warp w (w = warpid % 4) gets row 32w+lane
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- UMMA descriptors ----------------------------------------------------------------------------------
// shared-memory matrix descriptor, K-major, SWIZZLE_NONE, descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;
    return d;
}
// instruction descriptor for kind::tf32, fp32 accumulate, A and B K-major
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
    return (1u << 4)        // D format: F32
           | (2u << 7)      // A format: TF32
           | (2u << 10)     // B format: TF32
           | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// same with operand majorness bits: a_mn / b_mn = 1 selects an MN-major operand (bit 15 / 16), i.e. a tile whose
// 16-byte cells hold 4 consecutive M (or N) indices and whose 8-row core matrices run along K: the layout a K-major
// activation tile [rows = anchors][cols] has when the anchors are the REDUCTION dimension (weight gradients)
__host__ __device__ constexpr uint32_t make_idesc_tf32_mn(int M, int N, int a_mn, int b_mn) {
    return make_idesc_tf32(M, N) | ((uint32_t)(a_mn & 1) << 15) | ((uint32_t)(b_mn & 1) << 16);
}
// D[tmem] (+)= A[smem] * B[smem]^T   (single CTA); issue from ONE thread
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
// all previously issued MMAs of this thread arrive on the mbarrier when complete
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- operand helpers ---------------------------------------------------------------------------------------
// hi = x rounded to nearest TF32 (low 13 mantissa bits zero), so |x - hi| <= 2^-11 |x| and lo = x - hi is exact
__device__ __forceinline__ float tf32_hi(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// byte offset of the 16-byte cell (row r, chunk c) in a tile of R rows
__device__ __forceinline__ uint32_t cell_off(int r, int c, int R) { return (uint32_t)(c * R + r) * 16u; }

// Issue the 3xTF32 product of one operand pair over `ksteps` K=8 steps.
//   a_hi/a_lo: tiles whose chunk stride is lboA bytes, b_hi/b_lo: chunk stride lboB; a0, b0: first chunk ids.
// ONE thread issues every MMA of the kernel, so the issue loop itself is on the critical path (measured with the phase
// trace: building four descriptors from scratch per K step made 150 MMAs cost 6 us): the four descriptors are built once
// and stepped by adding the chunk-pair stride to their 14-bit address field (shared memory < 256 KB: no carry out).
__device__ __forceinline__ void issue_3xtf32_lbo(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, uint32_t lboA, int a0,
                                                 uint32_t b_hi, uint32_t b_lo, uint32_t lboB, int b0, int ksteps,
                                                 uint32_t idesc, bool accumulate_first) {
    const uint32_t ao = (uint32_t)a0 * lboA, bo = (uint32_t)b0 * lboB;
    uint64_t dah = make_desc(a_hi + ao, lboA, 128), dal = make_desc(a_lo + ao, lboA, 128);
    uint64_t dbh = make_desc(b_hi + bo, lboB, 128), dbl = make_desc(b_lo + bo, lboB, 128);
    const uint64_t sa = (uint64_t)((2u * lboA) >> 4), sbb = (uint64_t)((2u * lboB) >> 4);
#pragma unroll 4
    for (int s = 0; s < ksteps; ++s) {
        mma_tf32(d_tmem, dal, dbh, idesc, accumulate_first || s > 0);     // small terms first
        mma_tf32(d_tmem, dah, dbl, idesc, true);
        mma_tf32(d_tmem, dah, dbh, idesc, true);
        dah += sa; dal += sa; dbh += sbb; dbl += sbb;
    }
}
__device__ __forceinline__ void issue_3xtf32(uint32_t d_tmem, uint32_t a_hi, uint32_t a_lo, int RA, int a0,
                                             uint32_t b_hi, uint32_t b_lo, int RB, int b0, int ksteps,
                                             uint32_t idesc, bool accumulate_first) {
    issue_3xtf32_lbo(d_tmem, a_hi, a_lo, (uint32_t)RA * 16u, a0, b_hi, b_lo, (uint32_t)RB * 16u, b0, ksteps, idesc, accumulate_first);
}

}  // namespace tc
}  // namespace splatco
