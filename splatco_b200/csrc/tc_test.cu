// Self-test of the tcgen05 3xTF32 tile GEMM primitives (tc.cuh): C[M,N] = A[M,K] * B[N,K]^T, one
// 128-row tile per CTA, operands staged into the canonical no-swizzle K-major layout by the threads.
// Exercised by tests/test_tc_gpu.py against an fp64 reference; the fused decode kernels build on
// exactly these primitives.
#include "common.cuh"
#include "tc.cuh"

namespace splatco {

template <int NT>
__global__ void __launch_bounds__(128)
tc_gemm_test_kernel(int M, int N, int K, const float *__restrict__ A, const float *__restrict__ B, float *__restrict__ C,
                    int swap_lbo_sbo, int single_pass) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int Kc = (K + 7) / 8 * 2;                     // chunks (4 fp32 each), even
    uint8_t *a_hi = smem, *a_lo = a_hi + Kc * 128 * 16;
    uint8_t *b_hi = a_lo + Kc * 128 * 16, *b_lo = b_hi + Kc * NT * 16;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int m0 = blockIdx.x * 128;
    if (warp == 0) tc::tmem_alloc<128>(&tmem_base_s);
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
    // stage A (row = tid) and B (rows strided) with hi/lo split
    for (int c = 0; c < Kc; ++c) {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        const int gm = m0 + tid;
        float t[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { const int k = 4 * c + q; t[q] = (gm < M && k < K) ? A[(size_t)gm * K + k] : 0.f; }
        v = make_float4(tc::tf32_hi(t[0]), tc::tf32_hi(t[1]), tc::tf32_hi(t[2]), tc::tf32_hi(t[3]));
        *reinterpret_cast<float4 *>(a_hi + tc::cell_off(tid, c, 128)) = v;
        *reinterpret_cast<float4 *>(a_lo + tc::cell_off(tid, c, 128)) = make_float4(t[0] - v.x, t[1] - v.y, t[2] - v.z, t[3] - v.w);
        for (int n = tid; n < NT; n += 128) {
            float s[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) { const int k = 4 * c + q; s[q] = (n < N && k < K) ? B[(size_t)n * K + k] : 0.f; }
            const float4 h = make_float4(tc::tf32_hi(s[0]), tc::tf32_hi(s[1]), tc::tf32_hi(s[2]), tc::tf32_hi(s[3]));
            *reinterpret_cast<float4 *>(b_hi + tc::cell_off(n, c, NT)) = h;
            *reinterpret_cast<float4 *>(b_lo + tc::cell_off(n, c, NT)) = make_float4(s[0] - h.x, s[1] - h.y, s[2] - h.z, s[3] - h.w);
        }
    }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        constexpr uint32_t idesc = tc::make_idesc_tf32(128, NT);
        if (!swap_lbo_sbo && !single_pass) {
            tc::issue_3xtf32(tmem, tc::smem_u32(a_hi), tc::smem_u32(a_lo), 128, 0, tc::smem_u32(b_hi), tc::smem_u32(b_lo),
                             NT, 0, Kc / 2, idesc, false);
        } else {
            // diagnostic variants: swapped LBO/SBO convention and/or plain single-pass TF32
            for (int s = 0; s < Kc / 2; ++s) {
                const uint32_t lboA = 128 * 16, lboB = NT * 16;
                const uint32_t ao = 2 * s * lboA, bo = 2 * s * lboB;
                const uint32_t la = swap_lbo_sbo ? 128u : lboA, sa = swap_lbo_sbo ? lboA : 128u;
                const uint32_t lb = swap_lbo_sbo ? 128u : lboB, sb = swap_lbo_sbo ? lboB : 128u;
                tc::mma_tf32(tmem, tc::make_desc(tc::smem_u32(a_hi) + ao, la, sa), tc::make_desc(tc::smem_u32(b_hi) + bo, lb, sb),
                             idesc, s > 0);
                if (!single_pass) {
                    tc::mma_tf32(tmem, tc::make_desc(tc::smem_u32(a_lo) + ao, la, sa), tc::make_desc(tc::smem_u32(b_hi) + bo, lb, sb), idesc, true);
                    tc::mma_tf32(tmem, tc::make_desc(tc::smem_u32(a_hi) + ao, la, sa), tc::make_desc(tc::smem_u32(b_lo) + bo, lb, sb), idesc, true);
                }
            }
        }
        tc::mma_commit(&bar);
    }
    tc::mbar_wait(&bar, 0);
    tc::tc_fence_after();
    const int gm = m0 + tid;
    for (int n0 = 0; n0 < NT; n0 += 8) {
        float v[8];
        tc::tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + n0, v);
        tc::tmem_ld_wait();
        if (gm < M)
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (n0 + q < N) C[(size_t)gm * N + n0 + q] = v[q];
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<128>(tmem);
}

// Transposed-operand ("MN-major") self-test: C[m][n] = sum_k At[k][m] * Bt[k][n], K = 128 reduction rows, m < 128.
// An MN-major operand is staged exactly like the K-major activation tiles of the decode kernels (cell (row k, chunk
// m >> 2)) and read through an MN-major descriptor; a K-major operand is transposed by the threads into the canonical
// K-major layout (cell (row m, chunk k >> 2)).  variant bits: 0 A is MN-major, 1 B is MN-major, 2 swap the LBO / SBO
// roles of the MN-major descriptors, 3 single pass, [4,7) descriptor layout_type (0 = no swizzle), 7: LBO = 0.
template <int NT>
__global__ void __launch_bounds__(128)
tc_wgrad_test_kernel(int N, const float *__restrict__ At, const float *__restrict__ Bt, float *__restrict__ C, int variant) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    constexpr int NC = NT / 4;
    const int a_mn = variant & 1, b_mn = (variant >> 1) & 1, swap = (variant >> 2) & 1, single = (variant >> 3) & 1;
    const uint64_t ltype = (uint64_t)((variant >> 4) & 7) << 61;
    const int lbo0 = (variant >> 7) & 1;
    uint8_t *a_hi = smem, *a_lo = a_hi + 32 * 2048;
    uint8_t *b_hi = a_lo + 32 * 2048, *b_lo = b_hi + 32 * NT * 16;       // (K-major B: 32 chunks x NT rows; MN-major: NC chunks x 128 rows)
    const int tid = threadIdx.x, warp = tid >> 5;       // tid = reduction row k
    if (warp == 0) tc::tmem_alloc<128>(&tmem_base_s);
    if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
    for (int c = 0; c < 32; ++c) {
        const float4 x = *reinterpret_cast<const float4 *>(At + (size_t)tid * 128 + 4 * c);
        const float xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float h = tc::tf32_hi(xv[q]);
            // MN-major: cell (row k = tid, chunk c), sub q.   K-major: cell (row m = 4c+q, chunk tid >> 2), sub tid & 3
            const uint32_t off = a_mn ? tc::cell_off(tid, c, 128) + 4 * q : tc::cell_off(4 * c + q, tid >> 2, 128) + 4 * (tid & 3);
            *reinterpret_cast<float *>(a_hi + off) = h;
            *reinterpret_cast<float *>(a_lo + off) = xv[q] - h;
        }
    }
    for (int c = 0; c < NC; ++c) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float t = 4 * c + q < N ? Bt[(size_t)tid * N + 4 * c + q] : 0.f;
            const float h = tc::tf32_hi(t);
            const uint32_t off = b_mn ? tc::cell_off(tid, c, 128) + 4 * q : tc::cell_off(4 * c + q, tid >> 2, NT) + 4 * (tid & 3);
            *reinterpret_cast<float *>(b_hi + off) = h;
            *reinterpret_cast<float *>(b_lo + off) = t - h;
        }
    }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        const uint32_t idesc = tc::make_idesc_tf32_mn(128, NT, a_mn, b_mn);
        // MN-major: core matrix = 8 reduction rows x 16 B (128 B contiguous); next 4 M/N indices: +2048 B; next 8 k: +128 B
        uint32_t lbo = swap ? 2048u : 128u, sbo = swap ? 128u : 2048u;
        if (lbo0) { if (swap) sbo = 0; else lbo = 0; }
        for (int s = 0; s < 16; ++s) {
            const uint32_t oa = a_mn ? (uint32_t)s * 128u : (uint32_t)(2 * s) * 128 * 16;
            const uint32_t ob = b_mn ? (uint32_t)s * 128u : (uint32_t)(2 * s) * NT * 16;
            const uint32_t la = a_mn ? lbo : 128u * 16, sa = a_mn ? sbo : 128u;
            const uint32_t lb = b_mn ? lbo : (uint32_t)NT * 16, sb = b_mn ? sbo : 128u;
            const uint64_t ta = a_mn ? ltype : 0, tb = b_mn ? ltype : 0;
            const uint64_t dah = tc::make_desc(tc::smem_u32(a_hi) + oa, la, sa) | ta, dal = tc::make_desc(tc::smem_u32(a_lo) + oa, la, sa) | ta;
            const uint64_t dbh = tc::make_desc(tc::smem_u32(b_hi) + ob, lb, sb) | tb, dbl = tc::make_desc(tc::smem_u32(b_lo) + ob, lb, sb) | tb;
            if (!single) {
                tc::mma_tf32(tmem, dal, dbh, idesc, s > 0);
                tc::mma_tf32(tmem, dah, dbl, idesc, true);
            }
            tc::mma_tf32(tmem, dah, dbh, idesc, !single || s > 0);
        }
        tc::mma_commit(&bar);
    }
    tc::mbar_wait(&bar, 0);
    tc::tc_fence_after();
    for (int n0 = 0; n0 < NT; n0 += 8) {
        float v[8];
        tc::tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + n0, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 8; ++q)
            if (n0 + q < N) C[(size_t)tid * N + n0 + q] = v[q];
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc<128>(tmem);
}

}  // namespace splatco

using namespace splatco;

extern "C" int splatco_tc_wgrad_selftest(int N, const float *At, const float *Bt, float *C, int variant, void *stream) {
    SPLATCO_REQUIRE(N > 0 && N <= 96, "tc wgrad selftest: N <= 96 required");
    SPLATCO_REQUIRE(At && Bt && C, "tc wgrad selftest: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = (size_t)64 * 2048 + 2 * 32 * 96 * 16 + 1024;
    SPLATCO_CHECK_CUDA(cudaFuncSetAttribute(tc_wgrad_test_kernel<96>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_wgrad_test_kernel<96><<<1, 128, smem, st>>>(N, At, Bt, C, variant);
    SPLATCO_CHECK_LAUNCH();
    return 0;
}
extern "C" int splatco_tc_gemm_selftest(int M, int N, int K, const float *A, const float *B, float *C, int variant,
                                        void *stream) {
    SPLATCO_REQUIRE(M > 0 && N > 0 && K > 0 && N <= 112 && K <= 136, "tc selftest: N<=112, K<=136 required");
    SPLATCO_REQUIRE(A && B && C, "tc selftest: null pointer");
    const int Kc = (K + 7) / 8 * 2;
    const int NT = N <= 32 ? 32 : (N <= 96 ? 96 : 112);
    const size_t smem = (size_t)Kc * 16 * 2 * (128 + NT) + 1024;
    cudaStream_t st = (cudaStream_t)stream;
    const int swap = variant & 1, single = (variant >> 1) & 1;
    const int grid = ceil_div(M, 128);
    if (NT == 32) {
        SPLATCO_CHECK_CUDA(cudaFuncSetAttribute(tc_gemm_test_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tc_gemm_test_kernel<32><<<grid, 128, smem, st>>>(M, N, K, A, B, C, swap, single);
    } else if (NT == 96) {
        SPLATCO_CHECK_CUDA(cudaFuncSetAttribute(tc_gemm_test_kernel<96>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tc_gemm_test_kernel<96><<<grid, 128, smem, st>>>(M, N, K, A, B, C, swap, single);
    } else {
        SPLATCO_CHECK_CUDA(cudaFuncSetAttribute(tc_gemm_test_kernel<112>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tc_gemm_test_kernel<112><<<grid, 128, smem, st>>>(M, N, K, A, B, C, swap, single);
    }
    SPLATCO_CHECK_LAUNCH();
    return 0;
}
