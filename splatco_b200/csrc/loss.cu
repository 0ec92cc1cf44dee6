// Fused L1 + SSIM image loss, forward and backward.
// Reference being replaced (file:line in /root/reference):
//   utils/loss_utils.py:17-18    l1_loss  = mean |img - gt|
//   utils/loss_utils.py:33-63    ssim: 11x11 Gaussian window (sigma 1.5, outer product of the normalised 1-D window),
//                                five grouped conv2d with zero padding 5, C1 = 0.01^2, C2 = 0.03^2, mean of the map
//   train.py:192-196             loss = (1 - lambda) * Ll1 + lambda * (1 - ssim(image, gt))   (+ scaling term, not here)
// The reference spends 5 cuDNN grouped convolutions + ~15 elementwise launches per call and the same again in
// autograd's backward.  Here: one forward kernel (16x16-pixel tiles, 5-pixel halo in shared memory, separable
// window: 11 horizontal + 11 vertical taps instead of 121) that also stores the three partial-derivative maps
// dm/dmu1, dm/dE[x^2], dm/dE[xy], and one backward kernel that convolves those maps with the same window:
//   dL/dx = conv(dm/dmu1) + 2 x conv(dm/dE[x^2]) + y conv(dm/dE[xy])          (window symmetric)
// HBM-bound: forward reads 2 and writes 3 floats per pixel-channel, backward reads 5 and writes 1.
#include "common.cuh"

namespace splatco {

constexpr int LS_T = 16, LS_R = 5, LS_K = 11, LS_REG = LS_T + 2 * LS_R;      // 26
constexpr float LS_C1 = 0.01f * 0.01f, LS_C2 = 0.03f * 0.03f;

struct LossWindow { float g[LS_K]; };

// normalised 1-D Gaussian window, computed like utils/loss_utils.py:23-25 (fp32 exp, fp32 sum)
static LossWindow make_window() {
    LossWindow w;
    float s = 0.f;
    for (int x = 0; x < LS_K; ++x) { w.g[x] = expf(-(float)((x - LS_K / 2) * (x - LS_K / 2)) / (2.0f * 1.5f * 1.5f)); s += w.g[x]; }
    for (int x = 0; x < LS_K; ++x) w.g[x] /= s;
    return w;
}

// Tiles of 32 x 16 pixels, 256 threads.  The first version (16 x 16 tiles, one tap = one 4-byte shared-memory load) was bound
// by the shared-memory pipe (ncu: LSU 48 %, mio_throttle the top stall, 5.5 M bank conflicts).  Here the horizontal pass
// gives each thread FOUR adjacent output columns of a row: it fetches the 14 (16) inputs they share with 16-byte loads
// (a quarter warp reads 128 contiguous bytes: conflict-free) and the vertical pass gives each thread two vertically
// adjacent pixels (12 loads for 2 x 11 taps); the rows of the intermediate maps are padded to 33 floats.
constexpr int LT_W = 32, LT_H = 16, LT_RW = LT_W + 2 * LS_R + 2 /*44: 16-byte rows*/, LT_RH = LT_H + 2 * LS_R /*26*/;
constexpr int LT_ITEMS = LT_RH * (LT_W / 4);           // 208 (row, group of 4 columns) items of the horizontal pass

// sums[0] += sum |x - y|, sums[1] += sum ssim_map   (fp64 accumulators: the means are over ~1.6 M terms)
__global__ void __launch_bounds__(256)
l1_ssim_fwd_kernel(int H, int W, const float *__restrict__ img, const float *__restrict__ gt, LossWindow win,
                   float *__restrict__ d_mu1, float *__restrict__ d_exx, float *__restrict__ d_exy,
                   double *__restrict__ sums, double count, float lambda, float *__restrict__ out) {
    __shared__ __align__(16) float s_x[LT_RH][LT_RW], s_y[LT_RH][LT_RW];
    __shared__ float s_h[5][LT_RH][LT_W + 1];           // horizontally filtered x, y, xx, yy, xy
    __shared__ float s_red[2][8];
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * LT_W, y0 = blockIdx.y * LT_H;
    const size_t plane = (size_t)blockIdx.z * H * W;
    for (int r = tid; r < LT_RH * LT_RW; r += 256) {
        const int ry = r / LT_RW, rx = r - ry * LT_RW;
        const int gy = y0 + ry - LS_R, gx = x0 + rx - LS_R;
        float a = 0.f, b = 0.f;
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) { a = __ldg(img + plane + (size_t)gy * W + gx); b = __ldg(gt + plane + (size_t)gy * W + gx); }
        s_x[ry][rx] = a; s_y[ry][rx] = b;
    }
    __syncthreads();
    if (tid < LT_ITEMS) {
        const int ry = tid >> 3, cx0 = (tid & 7) * 4;
        const float4 *px = reinterpret_cast<const float4 *>(&s_x[ry][cx0]), *py = reinterpret_cast<const float4 *>(&s_y[ry][cx0]);
        const float4 a0 = px[0], a1 = px[1], a2 = px[2], a3 = px[3], b0 = py[0], b1 = py[1], b2 = py[2], b3 = py[3];
        const float xs[16] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, a2.x, a2.y, a2.z, a2.w, a3.x, a3.y, a3.z, a3.w};
        const float ys[16] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w, b2.x, b2.y, b2.z, b2.w, b3.x, b3.y, b3.z, b3.w};
        float acc[4][5];
#pragma unroll
        for (int o = 0; o < 4; ++o)
#pragma unroll
            for (int m = 0; m < 5; ++m) acc[o][m] = 0.f;
#pragma unroll
        for (int p = 0; p < LS_K + 3; ++p) {
            const float a = xs[p], b = ys[p], aa = a * a, bb = b * b, ab = a * b;
#pragma unroll
            for (int o = 0; o < 4; ++o) {
                const int k = p - o;
                if (k >= 0 && k < LS_K) {
                    const float g = win.g[k];
                    acc[o][0] = fmaf(g, a, acc[o][0]); acc[o][1] = fmaf(g, b, acc[o][1]);
                    acc[o][2] = fmaf(g, aa, acc[o][2]); acc[o][3] = fmaf(g, bb, acc[o][3]); acc[o][4] = fmaf(g, ab, acc[o][4]);
                }
            }
        }
#pragma unroll
        for (int o = 0; o < 4; ++o)
#pragma unroll
            for (int m = 0; m < 5; ++m) s_h[m][ry][cx0 + o] = acc[o][m];
    }
    __syncthreads();
    const int lx = tid & 31, ly0 = (tid >> 5) * 2;
    const int gx = x0 + lx;
    float v[2][5];
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int m = 0; m < 5; ++m) v[q][m] = 0.f;
#pragma unroll
    for (int k = 0; k < LS_K + 1; ++k) {
        float h[5];
#pragma unroll
        for (int m = 0; m < 5; ++m) h[m] = s_h[m][ly0 + k][lx];
        if (k < LS_K) {
            const float g = win.g[k];
#pragma unroll
            for (int m = 0; m < 5; ++m) v[0][m] = fmaf(g, h[m], v[0][m]);
        }
        if (k >= 1) {
            const float g = win.g[k - 1];
#pragma unroll
            for (int m = 0; m < 5; ++m) v[1][m] = fmaf(g, h[m], v[1][m]);
        }
    }
    float l1 = 0.f, msum = 0.f;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int gy = y0 + ly0 + q;
        if (gy < H && gx < W) {
            const float mu1 = v[q][0], mu2 = v[q][1], exx = v[q][2], eyy = v[q][3], exy = v[q][4];
            const float s1 = exx - mu1 * mu1, s2 = eyy - mu2 * mu2, s12 = exy - mu1 * mu2;
            const float A = 2.f * mu1 * mu2 + LS_C1, B = 2.f * s12 + LS_C2;
            const float Cc = mu1 * mu1 + mu2 * mu2 + LS_C1, D = s1 + s2 + LS_C2;
            const float inv = 1.f / (Cc * D);
            const float m = A * B * inv;
            // partial derivatives of m with E[x], E[x^2], E[xy] as the independent variables
            const size_t o = plane + (size_t)gy * W + gx;
            d_mu1[o] = 2.f * mu2 * (B - A) * inv - m * 2.f * mu1 * (D - Cc) * inv;
            d_exx[o] = -m / D;
            d_exy[o] = 2.f * A * inv;
            msum += m;
            l1 += fabsf(s_x[ly0 + q + LS_R][lx + LS_R] - s_y[ly0 + q + LS_R][lx + LS_R]);
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { l1 += __shfl_xor_sync(0xffffffffu, l1, d); msum += __shfl_xor_sync(0xffffffffu, msum, d); }
    if ((tid & 31) == 0) { s_red[0][tid >> 5] = l1; s_red[1][tid >> 5] = msum; }
    __syncthreads();
    if (tid < 2) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += (double)s_red[tid][w];
        atomicAdd(&sums[tid], s);
        __threadfence();
    }
    __syncthreads();
    // the CTA that takes the last ticket finishes: out[0] = loss, out[1] = l1 mean, out[2] = ssim mean
    if (tid == 0) {
        unsigned int *ticket = reinterpret_cast<unsigned int *>(sums + 2);
        const unsigned int total = gridDim.x * gridDim.y * gridDim.z;
        if (atomicAdd(ticket, 1u) == total - 1) {
            __threadfence();
            const double l1m = __ldcg(sums) / count, ss = __ldcg(sums + 1) / count;
            out[0] = (float)((1.0 - (double)lambda) * l1m + (double)lambda * (1.0 - ss));
            out[1] = (float)l1m;
            out[2] = (float)ss;
        }
    }
}

// dL/dimg = g_loss * [ (1 - lambda) sign(x - y) / count  -  lambda / count * (conv(d_mu1) + 2 x conv(d_exx) + y conv(d_exy)) ]
__global__ void __launch_bounds__(256)
l1_ssim_bwd_kernel(int H, int W, const float *__restrict__ img, const float *__restrict__ gt, LossWindow win,
                   const float *__restrict__ d_mu1, const float *__restrict__ d_exx, const float *__restrict__ d_exy,
                   const float *__restrict__ g_loss, float c_l1, float c_ssim, float *__restrict__ dimg) {
    __shared__ __align__(16) float s_m[3][LT_RH][LT_RW];
    __shared__ float s_h[3][LT_RH][LT_W + 1];
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * LT_W, y0 = blockIdx.y * LT_H;
    const size_t plane = (size_t)blockIdx.z * H * W;
    for (int r = tid; r < LT_RH * LT_RW; r += 256) {
        const int ry = r / LT_RW, rx = r - ry * LT_RW;
        const int gy = y0 + ry - LS_R, gx = x0 + rx - LS_R;
        float a = 0.f, b = 0.f, c = 0.f;
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
            const size_t o = plane + (size_t)gy * W + gx;
            a = __ldg(d_mu1 + o); b = __ldg(d_exx + o); c = __ldg(d_exy + o);
        }
        s_m[0][ry][rx] = a; s_m[1][ry][rx] = b; s_m[2][ry][rx] = c;
    }
    __syncthreads();
    if (tid < LT_ITEMS) {
        const int ry = tid >> 3, cx0 = (tid & 7) * 4;
#pragma unroll
        for (int m = 0; m < 3; ++m) {
            const float4 *pm = reinterpret_cast<const float4 *>(&s_m[m][ry][cx0]);
            const float4 a0 = pm[0], a1 = pm[1], a2 = pm[2], a3 = pm[3];
            const float xs[16] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, a2.x, a2.y, a2.z, a2.w, a3.x, a3.y, a3.z, a3.w};
            float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int p = 0; p < LS_K + 3; ++p)
#pragma unroll
                for (int o = 0; o < 4; ++o) {
                    const int k = p - o;
                    if (k >= 0 && k < LS_K) acc[o] = fmaf(win.g[k], xs[p], acc[o]);
                }
#pragma unroll
            for (int o = 0; o < 4; ++o) s_h[m][ry][cx0 + o] = acc[o];
        }
    }
    __syncthreads();
    const int lx = tid & 31, ly0 = (tid >> 5) * 2;
    const int gx = x0 + lx;
    float v[2][3];
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int m = 0; m < 3; ++m) v[q][m] = 0.f;
#pragma unroll
    for (int k = 0; k < LS_K + 1; ++k) {
        float h[3];
#pragma unroll
        for (int m = 0; m < 3; ++m) h[m] = s_h[m][ly0 + k][lx];
        if (k < LS_K) {
#pragma unroll
            for (int m = 0; m < 3; ++m) v[0][m] = fmaf(win.g[k], h[m], v[0][m]);
        }
        if (k >= 1) {
#pragma unroll
            for (int m = 0; m < 3; ++m) v[1][m] = fmaf(win.g[k - 1], h[m], v[1][m]);
        }
    }
    const float gl = __ldg(g_loss);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int gy = y0 + ly0 + q;
        if (gy >= H || gx >= W) continue;
        const size_t o = plane + (size_t)gy * W + gx;
        const float x = __ldg(img + o), y = __ldg(gt + o);
        const float dssim = v[q][0] + 2.f * x * v[q][1] + y * v[q][2];
        const float d = x - y;
        const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
        dimg[o] = gl * (c_l1 * sgn - c_ssim * dssim);
    }
}


// ---- cross-view consistency term of the mv batch (train.py:199-216,237-239) ------------------------------------
// For every pair i < j of the iteration's views the reference evaluates, after cropping the pair's four images to their
// common size (align_images, train.py:79-96),
//     s = ssim(real_i, real_j);   loss_ij = s * | l1_loss(real_i - real_j, gen_i - gen_j) |   if s > 0.6 else 0
// and adds 0.05 * sum_ij loss_ij to the summed loss: per pair 2 SSIMs (10 grouped convolutions), 3 elementwise passes
// and a host sync for the `if`.  Here one pass over the mv generated + mv real images forms all mv(mv-1)/2 sums
// (2 mv loads per pixel instead of 4 per pair), the gate is applied on the device, and the backward writes every
// view's dL/dgen in one pass.  s only depends on the ground-truth images; the caller supplies it (splatco_l1_ssim_fwd).
// A pixel belongs to pair (i, j) when it lies inside both images, i.e. inside the pair's crop.
constexpr int MVC_MAX = SPLATCO_MVC_MAX_VIEWS, MVC_PAIRS = MVC_MAX * (MVC_MAX - 1) / 2;

struct MvcViews {
    const float *gen[MVC_MAX], *real[MVC_MAX];
    float *dgen[MVC_MAX];
    int h[MVC_MAX], w[MVC_MAX];
};

template <int NV>
__global__ void __launch_bounds__(256)
mvc_fwd_kernel(int H, int W, int64_t total, MvcViews v, double *__restrict__ sums) {
    constexpr int NP = NV * (NV - 1) / 2;
    float acc[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) acc[p] = 0.f;
    const int64_t hw = (int64_t)H * W;
    for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (int64_t)gridDim.x * 256) {
        const int64_t c = e / hw, r = e - c * hw;
        const int y = (int)(r / W), x = (int)(r - (int64_t)y * W);
        float g[NV], t[NV];
        bool in[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            in[i] = y < v.h[i] && x < v.w[i];
            const int64_t o = (c * v.h[i] + y) * (int64_t)v.w[i] + x;
            g[i] = in[i] ? __ldg(v.gen[i] + o) : 0.f; t[i] = in[i] ? __ldg(v.real[i] + o) : 0.f;
        }
        int p = 0;
#pragma unroll
        for (int i = 0; i < NV; ++i)
#pragma unroll
            for (int j = i + 1; j < NV; ++j, ++p)
                if (in[i] && in[j]) acc[p] += fabsf((t[i] - t[j]) - (g[i] - g[j]));
    }
    __shared__ float s_red[8][NP];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        float a = acc[p];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) a += __shfl_xor_sync(0xffffffffu, a, d);
        if (lane == 0) s_red[warp][p] = a;
    }
    __syncthreads();
    if (threadIdx.x < NP) {
        double a = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) a += (double)s_red[w][threadIdx.x];
        atomicAdd(sums + threadIdx.x, a);
    }
}

struct MvcCounts { double n[MVC_PAIRS]; };

// w_p = s_p if s_p > gate else 0;  out[1+p] = w_p * mean_p,  out[0] = their sum;  weights[p] = w_p / count_p (backward)
__global__ void mvc_finish_kernel(int NP, const double *__restrict__ sums, const float *__restrict__ pair_ssim, float gate, MvcCounts cnt,
                                  float *__restrict__ weights, float *__restrict__ out) {
    float tot = 0.f;
    for (int p = 0; p < NP; ++p) {
        const float s = pair_ssim[p];
        const float w = s > gate ? s : 0.f;
        const float l = w * (float)(sums[p] / cnt.n[p]);
        weights[p] = (float)((double)w / cnt.n[p]); out[1 + p] = l; tot += l;
    }
    out[0] = tot;
}

template <int NV>
__global__ void __launch_bounds__(256)
mvc_bwd_kernel(int H, int W, int64_t total, MvcViews v, const float *__restrict__ weights, const float *__restrict__ g_loss) {
    constexpr int NP = NV * (NV - 1) / 2;
    const int64_t hw = (int64_t)H * W;
    const float gl = __ldg(g_loss);
    float wp[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) wp[p] = __ldg(weights + p);
    for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < total; e += (int64_t)gridDim.x * 256) {
        const int64_t c = e / hw, r = e - c * hw;
        const int y = (int)(r / W), x = (int)(r - (int64_t)y * W);
        float g[NV], t[NV], d[NV];
        int64_t o[NV];
        bool in[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            in[i] = y < v.h[i] && x < v.w[i];
            o[i] = (c * v.h[i] + y) * (int64_t)v.w[i] + x;
            g[i] = in[i] ? __ldg(v.gen[i] + o[i]) : 0.f; t[i] = in[i] ? __ldg(v.real[i] + o[i]) : 0.f; d[i] = 0.f;
        }
        int p = 0;
#pragma unroll
        for (int i = 0; i < NV; ++i)
#pragma unroll
            for (int j = i + 1; j < NV; ++j, ++p) {
                const float e_ij = (t[i] - t[j]) - (g[i] - g[j]);
                const float sg = (in[i] && in[j]) ? (e_ij > 0.f ? 1.f : (e_ij < 0.f ? -1.f : 0.f)) : 0.f;
                const float w = wp[p] * sg;
                d[i] -= w; d[j] += w;
            }
#pragma unroll
        for (int i = 0; i < NV; ++i)
            if (in[i]) v.dgen[i][o[i]] = gl * d[i];
    }
}

template <int NV> struct MvcLaunch {
    static void fwd(int grid, cudaStream_t st, int H, int W, int64_t total, const MvcViews &v, double *sums) {
        mvc_fwd_kernel<NV><<<grid, 256, 0, st>>>(H, W, total, v, sums);
    }
    static void bwd(int grid, cudaStream_t st, int H, int W, int64_t total, const MvcViews &v, const float *weights, const float *g) {
        mvc_bwd_kernel<NV><<<grid, 256, 0, st>>>(H, W, total, v, weights, g);
    }
};


// ---- scaling regulariser of the per-view loss: mean_i prod_j scaling[i, j]  (train.py:195) -----------------------
// torch's `scaling.prod(dim=1).mean()` is three launches forward, and its backward (prod_backward) counts the zeros of
// its input with `.item()` — a device-to-host sync at the START of every view's backward, which keeps the host from
// queueing the view's blend / decode backward until the previous view's has drained.  Here: one reduction forward
// (fp64 accumulator), one elementwise backward  d/ds[i,j] = g / M * prod_{k != j} s[i,k]  (exact with zeros, no sync).
__global__ void __launch_bounds__(256)
scaling_reg_fwd_kernel(int M, const float *__restrict__ s, double *__restrict__ sum, float *__restrict__ out) {
    float a = 0.f;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < M; i += gridDim.x * 256)
        a += s[3 * (size_t)i] * s[3 * (size_t)i + 1] * s[3 * (size_t)i + 2];
    __shared__ float red[8];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) a += __shfl_xor_sync(0xffffffffu, a, d);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += (double)red[w];
        atomicAdd(sum, t);
        __threadfence();
        // last ticket: the mean
        if (atomicAdd(reinterpret_cast<unsigned int *>(sum + 1), 1u) == gridDim.x - 1) {
            __threadfence();
            out[0] = (float)(__ldcg(sum) / (double)M);
        }
    }
}

__global__ void __launch_bounds__(256)
scaling_reg_bwd_kernel(int M, const float *__restrict__ s, const float *__restrict__ g_loss, float *__restrict__ ds) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= M) return;
    const float g = __ldg(g_loss) / (float)M;
    const float a = s[3 * (size_t)i], b = s[3 * (size_t)i + 1], c = s[3 * (size_t)i + 2];
    ds[3 * (size_t)i] = g * (b * c); ds[3 * (size_t)i + 1] = g * (a * c); ds[3 * (size_t)i + 2] = g * (a * b);
}

}  // namespace splatco

using namespace splatco;

static int loss_check(int C, int H, int W) {
    SPLATCO_REQUIRE(C >= 1 && C <= 65535 && H >= 1 && W >= 1 && (int64_t)C * H * W < 0x7fffffff, "l1_ssim: bad sizes C=%d H=%d W=%d", C, H, W);
    return 0;
}

extern "C" size_t splatco_loss_ws_bytes(int C, int H, int W) {
    return 3 * align_up((size_t)C * H * W * sizeof(float)) + align_up(3 * sizeof(double));
}

extern "C" int splatco_l1_ssim_fwd(int C, int H, int W, const float *img, const float *gt, float lambda_dssim, void *ws,
                                   float *out3, void *stream) {
    if (loss_check(C, H, W)) return -1;
    SPLATCO_REQUIRE(img && gt && ws && out3, "l1_ssim_fwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t map = align_up((size_t)C * H * W * sizeof(float));
    char *b = (char *)ws;
    double *sums = (double *)(b + 3 * map);
    SPLATCO_CHECK_CUDA(cudaMemsetAsync(sums, 0, 3 * sizeof(double), st));          // two sums + the finish ticket
    static const LossWindow win = make_window();
    const dim3 grid(ceil_div(W, LT_W), ceil_div(H, LT_H), C);
    l1_ssim_fwd_kernel<<<grid, 256, 0, st>>>(H, W, img, gt, win, (float *)b, (float *)(b + map), (float *)(b + 2 * map), sums,
                                                     (double)C * H * W, lambda_dssim, out3);
    SPLATCO_CHECK_LAUNCH();
    return 0;
}

extern "C" int splatco_l1_ssim_bwd(int C, int H, int W, const float *img, const float *gt, float lambda_dssim, const void *ws,
                                   const float *grad_loss, float *dL_dimg, void *stream) {
    if (loss_check(C, H, W)) return -1;
    SPLATCO_REQUIRE(img && gt && ws && grad_loss && dL_dimg, "l1_ssim_bwd: null pointer");
    const size_t map = align_up((size_t)C * H * W * sizeof(float));
    const char *b = (const char *)ws;
    static const LossWindow win = make_window();
    const float count = (float)((double)C * H * W);
    const dim3 grid(ceil_div(W, LT_W), ceil_div(H, LT_H), C);
    l1_ssim_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(H, W, img, gt, win, (const float *)b, (const float *)(b + map),
                                                                      (const float *)(b + 2 * map), grad_loss,
                                                                      (1.f - lambda_dssim) / count, lambda_dssim / count, dL_dimg);
    SPLATCO_CHECK_LAUNCH();
    return 0;
}

static int mvc_fill(int n_views, int C, const float *const *gen, const float *const *real, float *const *dgen,
                    const int *img_h, const int *img_w, MvcViews *v, int *Hmax, int *Wmax) {
    SPLATCO_REQUIRE(n_views >= 2 && n_views <= MVC_MAX, "mv_consistency: 2..%d views supported, got %d", MVC_MAX, n_views);
    SPLATCO_REQUIRE(C >= 1, "mv_consistency: bad channel count %d", C);
    SPLATCO_REQUIRE(gen && real && img_h && img_w, "mv_consistency: null pointer");
    memset(v, 0, sizeof(*v));
    *Hmax = *Wmax = 0;
    for (int i = 0; i < n_views; ++i) {
        SPLATCO_REQUIRE(gen[i] && real[i] && (!dgen || dgen[i]), "mv_consistency: null image pointer (view %d)", i);
        SPLATCO_REQUIRE(img_h[i] >= 1 && img_w[i] >= 1, "mv_consistency: view %d has size %dx%d", i, img_h[i], img_w[i]);
        v->gen[i] = gen[i]; v->real[i] = real[i]; v->dgen[i] = dgen ? dgen[i] : nullptr;
        v->h[i] = img_h[i]; v->w[i] = img_w[i];
        if (img_h[i] > *Hmax) *Hmax = img_h[i];
        if (img_w[i] > *Wmax) *Wmax = img_w[i];
    }
    return 0;
}

static int mvc_grid(int64_t total) {
    const int64_t want = (total + 256 * 4 - 1) / (256 * 4);      // ~4 pixels per thread
    const int64_t cap = 148 * 8;
    return (int)(want < 1 ? 1 : (want > cap ? cap : want));
}

extern "C" size_t splatco_mv_consistency_ws_bytes(int n_views) {
    const size_t np = (size_t)n_views * (n_views - 1) / 2;
    return align_up(np * sizeof(double)) + align_up(np * sizeof(float));
}

extern "C" int splatco_mv_consistency_fwd(int n_views, int C, const float *const *gen, const float *const *real, const int *img_h,
                                          const int *img_w, const float *pair_ssim, float ssim_gate, void *ws, float *out,
                                          void *stream) {
    MvcViews v;
    int H, W;
    if (mvc_fill(n_views, C, gen, real, nullptr, img_h, img_w, &v, &H, &W)) return -1;
    SPLATCO_REQUIRE(pair_ssim && ws && out, "mv_consistency_fwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int np = n_views * (n_views - 1) / 2;
    double *sums = (double *)ws;
    float *weights = (float *)((char *)ws + align_up(np * sizeof(double)));
    SPLATCO_CHECK_CUDA(cudaMemsetAsync(sums, 0, np * sizeof(double), st));
    MvcCounts cnt;
    for (int i = 0, p = 0; i < n_views; ++i)
        for (int j = i + 1; j < n_views; ++j, ++p)
            cnt.n[p] = (double)C * (img_h[i] < img_h[j] ? img_h[i] : img_h[j]) * (img_w[i] < img_w[j] ? img_w[i] : img_w[j]);
    const int64_t total = (int64_t)C * H * W;
    const int grid = mvc_grid(total);
    switch (n_views) {
        case 2: MvcLaunch<2>::fwd(grid, st, H, W, total, v, sums); break;
        case 3: MvcLaunch<3>::fwd(grid, st, H, W, total, v, sums); break;
        case 4: MvcLaunch<4>::fwd(grid, st, H, W, total, v, sums); break;
        case 5: MvcLaunch<5>::fwd(grid, st, H, W, total, v, sums); break;
        case 6: MvcLaunch<6>::fwd(grid, st, H, W, total, v, sums); break;
        case 7: MvcLaunch<7>::fwd(grid, st, H, W, total, v, sums); break;
        default: MvcLaunch<8>::fwd(grid, st, H, W, total, v, sums); break;
    }
    SPLATCO_CHECK_LAUNCH();
    mvc_finish_kernel<<<1, 1, 0, st>>>(np, sums, pair_ssim, ssim_gate, cnt, weights, out);
    SPLATCO_CHECK_LAUNCH();
    return 0;
}

extern "C" int splatco_mv_consistency_bwd(int n_views, int C, const float *const *gen, const float *const *real, float *const *dgen,
                                          const int *img_h, const int *img_w, const void *ws, const float *grad_loss, void *stream) {
    MvcViews v;
    int H, W;
    SPLATCO_REQUIRE(dgen, "mv_consistency_bwd: null pointer");
    if (mvc_fill(n_views, C, gen, real, dgen, img_h, img_w, &v, &H, &W)) return -1;
    SPLATCO_REQUIRE(ws && grad_loss, "mv_consistency_bwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int np = n_views * (n_views - 1) / 2;
    const float *weights = (const float *)((const char *)ws + align_up(np * sizeof(double)));
    const int64_t total = (int64_t)C * H * W;
    const int grid = mvc_grid(total);
    switch (n_views) {
        case 2: MvcLaunch<2>::bwd(grid, st, H, W, total, v, weights, grad_loss); break;
        case 3: MvcLaunch<3>::bwd(grid, st, H, W, total, v, weights, grad_loss); break;
        case 4: MvcLaunch<4>::bwd(grid, st, H, W, total, v, weights, grad_loss); break;
        case 5: MvcLaunch<5>::bwd(grid, st, H, W, total, v, weights, grad_loss); break;
        case 6: MvcLaunch<6>::bwd(grid, st, H, W, total, v, weights, grad_loss); break;
        case 7: MvcLaunch<7>::bwd(grid, st, H, W, total, v, weights, grad_loss); break;
        default: MvcLaunch<8>::bwd(grid, st, H, W, total, v, weights, grad_loss); break;
    }
    SPLATCO_CHECK_LAUNCH();
    return 0;
}

extern "C" int splatco_scaling_reg_fwd(int M, const float *scaling, void *ws, float *out, void *stream) {
    SPLATCO_REQUIRE(M >= 1, "scaling_reg_fwd: needs at least one row (mean of an empty tensor is NaN in the reference), M=%d", M);
    SPLATCO_REQUIRE(scaling && ws && out, "scaling_reg_fwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    SPLATCO_CHECK_CUDA(cudaMemsetAsync(ws, 0, 2 * sizeof(double), st));            // the sum + the finish ticket
    const int grid = ceil_div(M, 256 * 8) < 148 * 4 ? ceil_div(M, 256 * 8) : 148 * 4;
    scaling_reg_fwd_kernel<<<grid, 256, 0, st>>>(M, scaling, (double *)ws, out);
    SPLATCO_CHECK_LAUNCH();
    return 0;
}

extern "C" int splatco_scaling_reg_bwd(int M, const float *scaling, const float *grad_loss, float *d_scaling, void *stream) {
    SPLATCO_REQUIRE(M >= 1 && scaling && grad_loss && d_scaling, "scaling_reg_bwd: bad arguments");
    scaling_reg_bwd_kernel<<<ceil_div(M, 256), 256, 0, (cudaStream_t)stream>>>(M, scaling, grad_loss, d_scaling);
    SPLATCO_CHECK_LAUNCH();
    return 0;
}
