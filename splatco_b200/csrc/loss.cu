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

// sums[0] += sum |x - y|, sums[1] += sum ssim_map   (fp64 accumulators: the means are over ~1.6 M terms)
__global__ void __launch_bounds__(LS_T * LS_T)
l1_ssim_fwd_kernel(int H, int W, const float *__restrict__ img, const float *__restrict__ gt, LossWindow win,
                   float *__restrict__ d_mu1, float *__restrict__ d_exx, float *__restrict__ d_exy,
                   double *__restrict__ sums, double count, float lambda, float *__restrict__ out) {
    __shared__ float s_x[LS_REG][LS_REG + 1], s_y[LS_REG][LS_REG + 1];
    __shared__ float s_h[5][LS_REG][LS_T + 1];          // horizontally filtered x, y, xx, yy, xy
    __shared__ float s_red[2][LS_T * LS_T / 32];
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * LS_T, y0 = blockIdx.y * LS_T;
    const size_t plane = (size_t)blockIdx.z * H * W;
    for (int r = tid; r < LS_REG * LS_REG; r += LS_T * LS_T) {
        const int ry = r / LS_REG, rx = r - ry * LS_REG;
        const int gy = y0 + ry - LS_R, gx = x0 + rx - LS_R;
        float a = 0.f, b = 0.f;
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) { a = __ldg(img + plane + (size_t)gy * W + gx); b = __ldg(gt + plane + (size_t)gy * W + gx); }
        s_x[ry][rx] = a; s_y[ry][rx] = b;
    }
    __syncthreads();
    for (int r = tid; r < LS_REG * LS_T; r += LS_T * LS_T) {
        const int ry = r / LS_T, cx = r - ry * LS_T;
        float hx = 0.f, hy = 0.f, hxx = 0.f, hyy = 0.f, hxy = 0.f;
#pragma unroll
        for (int k = 0; k < LS_K; ++k) {
            const float a = s_x[ry][cx + k], b = s_y[ry][cx + k], g = win.g[k];
            hx = fmaf(g, a, hx); hy = fmaf(g, b, hy);
            hxx = fmaf(g, a * a, hxx); hyy = fmaf(g, b * b, hyy); hxy = fmaf(g, a * b, hxy);
        }
        s_h[0][ry][cx] = hx; s_h[1][ry][cx] = hy; s_h[2][ry][cx] = hxx; s_h[3][ry][cx] = hyy; s_h[4][ry][cx] = hxy;
    }
    __syncthreads();
    const int ly = tid / LS_T, lx = tid - ly * LS_T;
    const int gy = y0 + ly, gx = x0 + lx;
    float l1 = 0.f, m = 0.f;
    if (gy < H && gx < W) {
        float mu1 = 0.f, mu2 = 0.f, exx = 0.f, eyy = 0.f, exy = 0.f;
#pragma unroll
        for (int k = 0; k < LS_K; ++k) {
            const float g = win.g[k];
            mu1 = fmaf(g, s_h[0][ly + k][lx], mu1); mu2 = fmaf(g, s_h[1][ly + k][lx], mu2);
            exx = fmaf(g, s_h[2][ly + k][lx], exx); eyy = fmaf(g, s_h[3][ly + k][lx], eyy);
            exy = fmaf(g, s_h[4][ly + k][lx], exy);
        }
        const float s1 = exx - mu1 * mu1, s2 = eyy - mu2 * mu2, s12 = exy - mu1 * mu2;
        const float A = 2.f * mu1 * mu2 + LS_C1, B = 2.f * s12 + LS_C2;
        const float Cc = mu1 * mu1 + mu2 * mu2 + LS_C1, D = s1 + s2 + LS_C2;
        const float inv = 1.f / (Cc * D);
        m = A * B * inv;
        // partial derivatives of m with E[x], E[x^2], E[xy] as the independent variables
        const size_t o = plane + (size_t)gy * W + gx;
        d_mu1[o] = 2.f * mu2 * (B - A) * inv - m * 2.f * mu1 * (D - Cc) * inv;
        d_exx[o] = -m / D;
        d_exy[o] = 2.f * A * inv;
        l1 = fabsf(s_x[ly + LS_R][lx + LS_R] - s_y[ly + LS_R][lx + LS_R]);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) { l1 += __shfl_xor_sync(0xffffffffu, l1, d); m += __shfl_xor_sync(0xffffffffu, m, d); }
    if ((tid & 31) == 0) { s_red[0][tid >> 5] = l1; s_red[1][tid >> 5] = m; }
    __syncthreads();
    if (tid < 2) {
        double s = 0.0;
        for (int w = 0; w < LS_T * LS_T / 32; ++w) s += (double)s_red[tid][w];
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
            const double l1 = __ldcg(sums) / count, ss = __ldcg(sums + 1) / count;
            out[0] = (float)((1.0 - (double)lambda) * l1 + (double)lambda * (1.0 - ss));
            out[1] = (float)l1;
            out[2] = (float)ss;
        }
    }
}

// dL/dimg = g_loss * [ (1 - lambda) sign(x - y) / count  -  lambda / count * (conv(d_mu1) + 2 x conv(d_exx) + y conv(d_exy)) ]
__global__ void __launch_bounds__(LS_T * LS_T)
l1_ssim_bwd_kernel(int H, int W, const float *__restrict__ img, const float *__restrict__ gt, LossWindow win,
                   const float *__restrict__ d_mu1, const float *__restrict__ d_exx, const float *__restrict__ d_exy,
                   const float *__restrict__ g_loss, float c_l1, float c_ssim, float *__restrict__ dimg) {
    __shared__ float s_m[3][LS_REG][LS_REG + 1];
    __shared__ float s_h[3][LS_REG][LS_T + 1];
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * LS_T, y0 = blockIdx.y * LS_T;
    const size_t plane = (size_t)blockIdx.z * H * W;
    for (int r = tid; r < LS_REG * LS_REG; r += LS_T * LS_T) {
        const int ry = r / LS_REG, rx = r - ry * LS_REG;
        const int gy = y0 + ry - LS_R, gx = x0 + rx - LS_R;
        float a = 0.f, b = 0.f, c = 0.f;
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
            const size_t o = plane + (size_t)gy * W + gx;
            a = __ldg(d_mu1 + o); b = __ldg(d_exx + o); c = __ldg(d_exy + o);
        }
        s_m[0][ry][rx] = a; s_m[1][ry][rx] = b; s_m[2][ry][rx] = c;
    }
    __syncthreads();
    for (int r = tid; r < LS_REG * LS_T; r += LS_T * LS_T) {
        const int ry = r / LS_T, cx = r - ry * LS_T;
        float h0 = 0.f, h1 = 0.f, h2 = 0.f;
#pragma unroll
        for (int k = 0; k < LS_K; ++k) {
            const float g = win.g[k];
            h0 = fmaf(g, s_m[0][ry][cx + k], h0); h1 = fmaf(g, s_m[1][ry][cx + k], h1); h2 = fmaf(g, s_m[2][ry][cx + k], h2);
        }
        s_h[0][ry][cx] = h0; s_h[1][ry][cx] = h1; s_h[2][ry][cx] = h2;
    }
    __syncthreads();
    const int ly = tid / LS_T, lx = tid - ly * LS_T;
    const int gy = y0 + ly, gx = x0 + lx;
    if (gy >= H || gx >= W) return;
    float v0 = 0.f, v1 = 0.f, v2 = 0.f;
#pragma unroll
    for (int k = 0; k < LS_K; ++k) {
        const float g = win.g[k];
        v0 = fmaf(g, s_h[0][ly + k][lx], v0); v1 = fmaf(g, s_h[1][ly + k][lx], v1); v2 = fmaf(g, s_h[2][ly + k][lx], v2);
    }
    const size_t o = plane + (size_t)gy * W + gx;
    const float x = __ldg(img + o), y = __ldg(gt + o);
    const float dssim = v0 + 2.f * x * v1 + y * v2;
    const float d = x - y;
    const float sgn = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
    dimg[o] = __ldg(g_loss) * (c_l1 * sgn - c_ssim * dssim);
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
    const dim3 grid(ceil_div(W, LS_T), ceil_div(H, LS_T), C);
    l1_ssim_fwd_kernel<<<grid, LS_T * LS_T, 0, st>>>(H, W, img, gt, win, (float *)b, (float *)(b + map), (float *)(b + 2 * map), sums,
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
    const dim3 grid(ceil_div(W, LS_T), ceil_div(H, LS_T), C);
    l1_ssim_bwd_kernel<<<grid, LS_T * LS_T, 0, (cudaStream_t)stream>>>(H, W, img, gt, win, (const float *)b, (const float *)(b + map),
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
