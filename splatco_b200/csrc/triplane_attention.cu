// TriPlaneAttention over the level-0 feature planes, forward + backward, on sm_100a.
// Reference being replaced (file:line in /root/reference):
//   scene/grids.py:22-36   ChannelAttention: global avg / max pool -> 1x1 conv C -> C/5 -> ReLU -> 1x1 conv -> sigmoid(avg + max)
//   scene/grids.py:38-52   SpatialAttention: channel mean / max -> 7x7 conv (2 -> 1, zero padding 3, no bias) -> sigmoid
//   scene/grids.py:55-64   TriPlaneAttention: x = ca(x) * x;  x = sa(x) * x
//   scene/grids.py:166-169 applied to cat(xy, xz, yz planes) of the TA level on EVERY decode call
// The op is view-independent, so the host evaluates it once per iteration (decode.py caches it) and
// its backward runs once on the summed gradients of all views.
//
// Everything is HBM-bound elementwise / stencil work over C x E x E floats (C = 15, E = plane_size/4):
//   forward   ta_pool (read C/px)  ->  ta_channel (one CTA)  ->  ta_apply (read ~1.4 C/px, write C+3/px)
//   backward  ta_bwd_spatial (read ~2.8 C/px, write 2/px)  ->  ta_bwd_channel (one CTA)  ->  ta_bwd_apply
//             (read 2 C/px, read-modify-write C/px into the caller's accumulating plane gradients)
// Tiles are 32 x 32 pixels with a 3-pixel halo held in shared memory; the 7x7 stencil, its transpose
// and the 98 weight-gradient sums all run out of that tile.
#include "common.cuh"

namespace splatco {

constexpr int TA_MAXC = 32;        // channels of the concatenated planes (3 * rc)
constexpr int TA_MAXH = 8;         // hidden width of the channel-attention MLP (C / 5)
constexpr int TA_K = 7, TA_R = 3;  // spatial-attention kernel / radius
constexpr int TA_T = 32;           // tile edge
constexpr int TA_REG = TA_T + 2 * TA_R;          // 38
constexpr int TA_POOL_THREADS = 256;

struct TAPlanes {
    const float *p[3];
    // select without dynamic indexing (which would spill the parameter struct to local memory)
    __device__ __forceinline__ const float *at(int i) const { return i == 0 ? p[0] : (i == 1 ? p[1] : p[2]); }
};
struct TAPlanesOut {
    float *p[3];
    __device__ __forceinline__ float *at(int i) const { return i == 0 ? p[0] : (i == 1 ? p[1] : p[2]); }
};

__device__ __forceinline__ float sigmoidf_acc(float x) { return 1.f / (1.f + expf(-x)); }

// small per-model state kept in the forward workspace for the backward
struct TASmall {
    float ca[TA_MAXC], avg[TA_MAXC], mx[TA_MAXC];
    int argmax[TA_MAXC];                  // pixel index of each channel's global maximum
    float ha[TA_MAXH], hm[TA_MAXH];       // post-ReLU hidden activations of the avg / max branch
    float d_ca[TA_MAXC];                  // backward accumulators (zeroed by the backward entry point)
    float d_wsa[2 * TA_K * TA_K];
    float d_avg[TA_MAXC], d_mx[TA_MAXC];
};

// ---- forward 1: per-channel sum / max / argmax partials ------------------------------------------------
__global__ void __launch_bounds__(TA_POOL_THREADS)
ta_pool_kernel(TAPlanes x, int C, int rc, int npix, float *__restrict__ part_sum, float *__restrict__ part_max,
               int *__restrict__ part_arg) {
    __shared__ float s_sum[TA_POOL_THREADS / 32], s_max[TA_POOL_THREADS / 32];
    __shared__ int s_arg[TA_POOL_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int c = 0; c < C; ++c) {
        const float *src = x.at(c / rc) + (size_t)(c % rc) * npix;
        float s = 0.f, m = -INFINITY;
        int a = 0;
        for (int i = blockIdx.x * TA_POOL_THREADS + threadIdx.x; i < npix; i += gridDim.x * TA_POOL_THREADS) {
            const float v = __ldg(src + i);
            s += v;
            if (v > m) { m = v; a = i; }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, d);
            const float om = __shfl_xor_sync(0xffffffffu, m, d);
            const int oa = __shfl_xor_sync(0xffffffffu, a, d);
            if (om > m || (om == m && oa < a)) { m = om; a = oa; }
        }
        if (lane == 0) { s_sum[warp] = s; s_max[warp] = m; s_arg[warp] = a; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < TA_POOL_THREADS / 32; ++w) {
                s += s_sum[w];
                if (s_max[w] > m || (s_max[w] == m && s_arg[w] < a)) { m = s_max[w]; a = s_arg[w]; }
            }
            part_sum[(size_t)c * gridDim.x + blockIdx.x] = s;
            part_max[(size_t)c * gridDim.x + blockIdx.x] = m;
            part_arg[(size_t)c * gridDim.x + blockIdx.x] = a;
        }
        __syncthreads();
    }
}

// ---- forward 2: pooled vectors -> channel attention (one CTA, one warp per channel) ---------------------
__global__ void __launch_bounds__(1024)
ta_channel_kernel(int C, int hid, int npix, int nblk, const float *__restrict__ part_sum, const float *__restrict__ part_max,
                  const int *__restrict__ part_arg, const float *__restrict__ w1 /*[hid][C]*/,
                  const float *__restrict__ w2 /*[C][hid]*/, TASmall *__restrict__ sm) {
    __shared__ float s_avg[TA_MAXC], s_mx[TA_MAXC], s_ha[TA_MAXH], s_hm[TA_MAXH];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp < C) {
        double s = 0.0;
        float m = -INFINITY;
        int a = 0;
        for (int b = lane; b < nblk; b += 32) {
            s += (double)part_sum[(size_t)warp * nblk + b];
            const float om = part_max[(size_t)warp * nblk + b];
            const int oa = part_arg[(size_t)warp * nblk + b];
            if (om > m || (om == m && oa < a)) { m = om; a = oa; }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, d);
            const float om = __shfl_xor_sync(0xffffffffu, m, d);
            const int oa = __shfl_xor_sync(0xffffffffu, a, d);
            if (om > m || (om == m && oa < a)) { m = om; a = oa; }
        }
        if (lane == 0) {
            s_avg[warp] = (float)(s / (double)npix); s_mx[warp] = m;
            sm->avg[warp] = s_avg[warp]; sm->mx[warp] = m; sm->argmax[warp] = a;
        }
    }
    __syncthreads();
    if (threadIdx.x < hid) {
        float a = 0.f, m = 0.f;
        for (int c = 0; c < C; ++c) { a = fmaf(w1[threadIdx.x * C + c], s_avg[c], a); m = fmaf(w1[threadIdx.x * C + c], s_mx[c], m); }
        s_ha[threadIdx.x] = fmaxf(a, 0.f); s_hm[threadIdx.x] = fmaxf(m, 0.f);
        sm->ha[threadIdx.x] = s_ha[threadIdx.x]; sm->hm[threadIdx.x] = s_hm[threadIdx.x];
    }
    __syncthreads();
    if (threadIdx.x < C) {
        float za = 0.f, zm = 0.f;
        for (int j = 0; j < hid; ++j) { za = fmaf(w2[threadIdx.x * hid + j], s_ha[j], za); zm = fmaf(w2[threadIdx.x * hid + j], s_hm[j], zm); }
        sm->ca[threadIdx.x] = sigmoidf_acc(za + zm);
    }
}

// ---- forward 3: spatial attention + output ------------------------------------------------------------
// region pixel r = (ry, rx) in [0, 38)^2 <-> image pixel (y0 + ry - 3, x0 + rx - 3); outside the image the
// conv input is zero (zero padding).
__global__ void __launch_bounds__(TA_T * TA_T)
ta_apply_kernel(TAPlanes x, TAPlanesOut out, int C, int rc, int E, const float *__restrict__ w_sa,
                const TASmall *__restrict__ sm, float *__restrict__ sa_out, float *__restrict__ smean_out,
                float *__restrict__ smax_out, uint8_t *__restrict__ arg_out) {
    __shared__ float s_mean[TA_REG * TA_REG], s_max[TA_REG * TA_REG];
    __shared__ float s_w[2 * TA_K * TA_K], s_ca[TA_MAXC];
    const int tid = threadIdx.x;
    if (tid < 2 * TA_K * TA_K) s_w[tid] = w_sa[tid];
    if (tid < C) s_ca[tid] = sm->ca[tid];
    __syncthreads();
    const int x0 = blockIdx.x * TA_T, y0 = blockIdx.y * TA_T;
    const size_t npix = (size_t)E * E;
    const float invC = 1.f / (float)C;
    for (int r = tid; r < TA_REG * TA_REG; r += TA_T * TA_T) {
        const int ry = r / TA_REG, rx = r - ry * TA_REG;
        const int gy = y0 + ry - TA_R, gx = x0 + rx - TA_R;
        float mean = 0.f, mx = 0.f;
        if (gy >= 0 && gy < E && gx >= 0 && gx < E) {
            const size_t pix = (size_t)gy * E + gx;
            float s = 0.f, m = -INFINITY;
            for (int c = 0; c < C; ++c) {
                const float v = s_ca[c] * __ldg(x.at(c / rc) + (size_t)(c % rc) * npix + pix);
                s += v; m = fmaxf(m, v);
            }
            mean = s * invC; mx = m;
        }
        s_mean[r] = mean; s_max[r] = mx;
    }
    __syncthreads();
    const int ly = tid / TA_T, lx = tid - ly * TA_T;
    const int gy = y0 + ly, gx = x0 + lx;
    if (gy >= E || gx >= E) return;
    float pre = 0.f;
#pragma unroll
    for (int dy = 0; dy < TA_K; ++dy)
#pragma unroll
        for (int dx = 0; dx < TA_K; ++dx) {
            const int r = (ly + dy) * TA_REG + lx + dx;
            pre = fmaf(s_w[dy * TA_K + dx], s_mean[r], pre);
            pre = fmaf(s_w[TA_K * TA_K + dy * TA_K + dx], s_max[r], pre);
        }
    const float sa = sigmoidf_acc(pre);
    const size_t pix = (size_t)gy * E + gx;
    float m = -INFINITY;
    int arg = 0;
    for (int c = 0; c < C; ++c) {
        const float v = s_ca[c] * __ldg(x.at(c / rc) + (size_t)(c % rc) * npix + pix);
        if (v > m) { m = v; arg = c; }
        out.at(c / rc)[(size_t)(c % rc) * npix + pix] = sa * v;
    }
    sa_out[pix] = sa;
    smean_out[pix] = s_mean[(ly + TA_R) * TA_REG + lx + TA_R];
    smax_out[pix] = m;
    arg_out[pix] = (uint8_t)arg;
}

// ---- backward 1: through the spatial attention ---------------------------------------------------------
// d_pre over the tile + halo, then per interior pixel the transposed stencil d_s = W^T * d_pre, the 98
// conv-weight gradient sums and the C channel-attention gradient sums.
__global__ void __launch_bounds__(TA_T * TA_T)
ta_bwd_spatial_kernel(TAPlanes x, TAPlanes g, int C, int rc, int E, const float *__restrict__ w_sa,
                      TASmall *__restrict__ sm, const float *__restrict__ sa_in, const float *__restrict__ smean_in,
                      const float *__restrict__ smax_in, const uint8_t *__restrict__ arg_in,
                      float *__restrict__ ds_mean, float *__restrict__ ds_max) {
    __shared__ float s_dpre[TA_REG * TA_REG];
    __shared__ float s_sm[TA_T * (TA_T + 1)], s_sx[TA_T * (TA_T + 1)];
    __shared__ float s_w[2 * TA_K * TA_K], s_ca[TA_MAXC], s_red[TA_MAXC];
    const int tid = threadIdx.x;
    if (tid < 2 * TA_K * TA_K) s_w[tid] = w_sa[tid];
    if (tid < C) { s_ca[tid] = sm->ca[tid]; s_red[tid] = 0.f; }
    __syncthreads();
    const int x0 = blockIdx.x * TA_T, y0 = blockIdx.y * TA_T;
    const size_t npix = (size_t)E * E;
    for (int r = tid; r < TA_REG * TA_REG; r += TA_T * TA_T) {
        const int ry = r / TA_REG, rx = r - ry * TA_REG;
        const int gy = y0 + ry - TA_R, gx = x0 + rx - TA_R;
        float dpre = 0.f;
        if (gy >= 0 && gy < E && gx >= 0 && gx < E) {
            const size_t pix = (size_t)gy * E + gx;
            float dsa = 0.f;
            for (int c = 0; c < C; ++c) {
                const size_t o = (size_t)(c % rc) * npix + pix;
                dsa = fmaf(__ldg(g.at(c / rc) + o), s_ca[c] * __ldg(x.at(c / rc) + o), dsa);
            }
            const float sa = __ldg(sa_in + pix);
            dpre = dsa * sa * (1.f - sa);
        }
        s_dpre[r] = dpre;
    }
    const int ly = tid / TA_T, lx = tid - ly * TA_T;
    const int gy = y0 + ly, gx = x0 + lx;
    const bool inside = gy < E && gx < E;
    const size_t pix = (size_t)gy * E + gx;
    s_sm[ly * (TA_T + 1) + lx] = inside ? __ldg(smean_in + pix) : 0.f;
    s_sx[ly * (TA_T + 1) + lx] = inside ? __ldg(smax_in + pix) : 0.f;
    __syncthreads();
    // transposed stencil: d_s_k[p] = sum_{dy,dx} w[k][dy][dx] * d_pre[p - (dy-3, dx-3)]
    float dsm = 0.f, dsx = 0.f;
#pragma unroll
    for (int dy = 0; dy < TA_K; ++dy)
#pragma unroll
        for (int dx = 0; dx < TA_K; ++dx) {
            const float d = s_dpre[(ly + 2 * TA_R - dy) * TA_REG + lx + 2 * TA_R - dx];
            dsm = fmaf(s_w[dy * TA_K + dx], d, dsm);
            dsx = fmaf(s_w[TA_K * TA_K + dy * TA_K + dx], d, dsx);
        }
    if (inside) { ds_mean[pix] = dsm; ds_max[pix] = dsx; }
    {
        // channel-attention gradient: d_ca[c] += d_x1[c] * x[c],  d_x1[c] = g[c] sa + d_s_mean / C + d_s_max [c == arg]
        // (pixels outside the image contribute zeros; the warp shuffles stay convergent)
        const float sa = inside ? __ldg(sa_in + pix) : 0.f;
        const int arg = inside ? (int)arg_in[pix] : -1;
        const float base = dsm / (float)C;
        for (int c = 0; c < C; ++c) {
            float v = 0.f;
            if (inside) {
                const size_t o = (size_t)(c % rc) * npix + pix;
                float dx1 = fmaf(__ldg(g.at(c / rc) + o), sa, base);
                if (c == arg) dx1 += dsx;
                v = dx1 * __ldg(x.at(c / rc) + o);
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
            if ((tid & 31) == 0 && v != 0.f) atomicAdd(&s_red[c], v);
        }
    }
    // conv-weight gradient: dW[k][dy][dx] = sum_q s_k[q] * d_pre[q - (dy-3, dx-3)], q over this tile's pixels.
    // 8 threads per tap, each covering 4 rows of the tile, then a 3-step shuffle reduction.
    {
        const int tap = tid >> 3, part = tid & 7;
        float acc = 0.f;
        if (tap < 2 * TA_K * TA_K) {
            const int k = tap / (TA_K * TA_K), t2 = tap - k * TA_K * TA_K;
            const int dy = t2 / TA_K, dx = t2 - dy * TA_K;
            const float *sk = k ? s_sx : s_sm;
            for (int row = part * 4; row < part * 4 + 4; ++row) {
                const float *dp = s_dpre + (row + 2 * TA_R - dy) * TA_REG + 2 * TA_R - dx;
                const float *sp = sk + row * (TA_T + 1);
#pragma unroll 8
                for (int col = 0; col < TA_T; ++col) acc = fmaf(sp[col], dp[col], acc);
            }
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 4);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        if (part == 0 && tap < 2 * TA_K * TA_K && acc != 0.f) atomicAdd(&sm->d_wsa[tap], acc);
    }
    __syncthreads();
    if (tid < C && s_red[tid] != 0.f) atomicAdd(&sm->d_ca[tid], s_red[tid]);
}

// ---- backward 2: through the channel-attention MLP (one CTA) --------------------------------------------
__global__ void __launch_bounds__(256)
ta_bwd_channel_kernel(int C, int hid, const float *__restrict__ w1, const float *__restrict__ w2, TASmall *__restrict__ sm,
                      float *__restrict__ g_w1, float *__restrict__ g_w2, float *__restrict__ g_wsa) {
    __shared__ float s_dz[TA_MAXC], s_dha[TA_MAXH], s_dhm[TA_MAXH];
    const int tid = threadIdx.x;
    if (tid < C) { const float ca = sm->ca[tid]; s_dz[tid] = sm->d_ca[tid] * ca * (1.f - ca); }
    __syncthreads();
    if (tid < hid) {
        float dh = 0.f;
        for (int c = 0; c < C; ++c) dh = fmaf(w2[c * hid + tid], s_dz[c], dh);
        s_dha[tid] = sm->ha[tid] > 0.f ? dh : 0.f;
        s_dhm[tid] = sm->hm[tid] > 0.f ? dh : 0.f;
    }
    __syncthreads();
    for (int e = tid; e < C * hid; e += 256) {            // w2 [C][hid]
        const int c = e / hid, j = e - c * hid;
        if (g_w2) g_w2[e] += s_dz[c] * (sm->ha[j] + sm->hm[j]);
    }
    for (int e = tid; e < hid * C; e += 256) {            // w1 [hid][C]
        const int j = e / C, c = e - j * C;
        if (g_w1) g_w1[e] += s_dha[j] * sm->avg[c] + s_dhm[j] * sm->mx[c];
    }
    if (tid < C) {
        float da = 0.f, dm = 0.f;
        for (int j = 0; j < hid; ++j) { da = fmaf(w1[j * C + tid], s_dha[j], da); dm = fmaf(w1[j * C + tid], s_dhm[j], dm); }
        sm->d_avg[tid] = da; sm->d_mx[tid] = dm;
    }
    if (g_wsa && tid < 2 * TA_K * TA_K) g_wsa[tid] += sm->d_wsa[tid];
}

// ---- backward 3: plane gradients (accumulated into the caller's buffers) ---------------------------------
__global__ void __launch_bounds__(256)
ta_bwd_apply_kernel(TAPlanes g, TAPlanesOut gx, int C, int rc, int E, const TASmall *__restrict__ sm,
                    const float *__restrict__ sa_in, const uint8_t *__restrict__ arg_in,
                    const float *__restrict__ ds_mean, const float *__restrict__ ds_max) {
    __shared__ float s_ca[TA_MAXC], s_davg[TA_MAXC], s_dmx[TA_MAXC];
    __shared__ int s_arg[TA_MAXC];
    const size_t npix = (size_t)E * E;
    if (threadIdx.x < C) {
        s_ca[threadIdx.x] = sm->ca[threadIdx.x];
        s_davg[threadIdx.x] = sm->d_avg[threadIdx.x] / (float)npix;
        s_dmx[threadIdx.x] = sm->d_mx[threadIdx.x];
        s_arg[threadIdx.x] = sm->argmax[threadIdx.x];
    }
    __syncthreads();
    for (size_t pix = (size_t)blockIdx.x * 256 + threadIdx.x; pix < npix; pix += (size_t)gridDim.x * 256) {
        const float sa = __ldg(sa_in + pix);
        const int arg = arg_in[pix];
        const float base = __ldg(ds_mean + pix) / (float)C, dsx = __ldg(ds_max + pix);
        for (int c = 0; c < C; ++c) {
            const size_t o = (size_t)(c % rc) * npix + pix;
            float dx1 = fmaf(__ldg(g.at(c / rc) + o), sa, base);
            if (c == arg) dx1 += dsx;
            float v = fmaf(dx1, s_ca[c], s_davg[c]);
            if ((int)pix == s_arg[c]) v += s_dmx[c];
            gx.at(c / rc)[o] += v;
        }
    }
}

struct TAFwdWs { TASmall *sm; float *sa, *smean, *smax; uint8_t *arg; float *part_sum, *part_max; int *part_arg; };
struct TABwdWs { float *ds_mean, *ds_max; };

static int ta_pool_blocks(int E) { return min(148 * 4, ceil_div(E * E, TA_POOL_THREADS)); }

static size_t ta_fwd_offsets(int C, int E, size_t off[9]) {
    const size_t npix = (size_t)E * E, nb = (size_t)ta_pool_blocks(E);
    size_t o = 0;
    off[0] = o; o += align_up(sizeof(TASmall));
    off[1] = o; o += align_up(npix * 4);
    off[2] = o; o += align_up(npix * 4);
    off[3] = o; o += align_up(npix * 4);
    off[4] = o; o += align_up(npix);
    off[5] = o; o += align_up(nb * C * 4);
    off[6] = o; o += align_up(nb * C * 4);
    off[7] = o; o += align_up(nb * C * 4);
    off[8] = o;
    return o;
}
static TAFwdWs ta_fwd_view(void *ws, int C, int E) {
    size_t off[9];
    ta_fwd_offsets(C, E, off);
    char *b = (char *)ws;
    TAFwdWs v;
    v.sm = (TASmall *)(b + off[0]); v.sa = (float *)(b + off[1]); v.smean = (float *)(b + off[2]); v.smax = (float *)(b + off[3]);
    v.arg = (uint8_t *)(b + off[4]); v.part_sum = (float *)(b + off[5]); v.part_max = (float *)(b + off[6]);
    v.part_arg = (int *)(b + off[7]);
    return v;
}

}  // namespace splatco

using namespace splatco;

static int ta_check(int rc, int E, int hidden, int ksize) {
    SPLATCO_REQUIRE(rc >= 1 && 3 * rc <= TA_MAXC, "triplane attention: %d channels per plane unsupported", rc);
    SPLATCO_REQUIRE(hidden >= 1 && hidden <= TA_MAXH, "triplane attention: hidden width %d unsupported", hidden);
    SPLATCO_REQUIRE(ksize == TA_K, "triplane attention: spatial kernel size %d unsupported (the reference constructs 7)", ksize);
    SPLATCO_REQUIRE(E >= 1 && (int64_t)E * E < 0x7fffffff, "triplane attention: bad plane edge %d", E);
    return 0;
}

extern "C" size_t splatco_ta_fwd_ws_bytes(int rc, int E) {
    size_t off[9];
    return ta_fwd_offsets(3 * rc, E, off);
}
extern "C" size_t splatco_ta_bwd_ws_bytes(int rc, int E) { return 2 * align_up((size_t)E * E * 4); }

extern "C" int splatco_ta_fwd(int rc, int E, int hidden, int ksize, const float *xy, const float *xz, const float *yz,
                              const float *w_ca1, const float *w_ca2, const float *w_sa, void *ws, float *out_xy,
                              float *out_xz, float *out_yz, void *stream) {
    if (ta_check(rc, E, hidden, ksize)) return -1;
    SPLATCO_REQUIRE(xy && xz && yz && w_ca1 && w_ca2 && w_sa && ws && out_xy && out_xz && out_yz, "ta_fwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int C = 3 * rc, npix = E * E;
    TAFwdWs f = ta_fwd_view(ws, C, E);
    TAPlanes x = {{xy, xz, yz}};
    TAPlanesOut o = {{out_xy, out_xz, out_yz}};
    const int nblk = ta_pool_blocks(E);
    ta_pool_kernel<<<nblk, TA_POOL_THREADS, 0, st>>>(x, C, rc, npix, f.part_sum, f.part_max, f.part_arg);
    SPLATCO_CHECK_LAUNCH();
    ta_channel_kernel<<<1, 1024, 0, st>>>(C, hidden, npix, nblk, f.part_sum, f.part_max, f.part_arg, w_ca1, w_ca2, f.sm);
    SPLATCO_CHECK_LAUNCH();
    const dim3 grid(ceil_div(E, TA_T), ceil_div(E, TA_T));
    ta_apply_kernel<<<grid, TA_T * TA_T, 0, st>>>(x, o, C, rc, E, w_sa, f.sm, f.sa, f.smean, f.smax, f.arg);
    SPLATCO_CHECK_LAUNCH();
    return 0;
}

extern "C" int splatco_ta_bwd(int rc, int E, int hidden, int ksize, const float *xy, const float *xz, const float *yz,
                              const float *w_ca1, const float *w_ca2, const float *w_sa, void *fwd_ws, void *bwd_ws,
                              const float *g_out_xy, const float *g_out_xz, const float *g_out_yz, float *g_xy,
                              float *g_xz, float *g_yz, float *g_w_ca1, float *g_w_ca2, float *g_w_sa, void *stream) {
    if (ta_check(rc, E, hidden, ksize)) return -1;
    SPLATCO_REQUIRE(xy && xz && yz && w_ca1 && w_ca2 && w_sa && fwd_ws && bwd_ws && g_out_xy && g_out_xz && g_out_yz &&
                    g_xy && g_xz && g_yz, "ta_bwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int C = 3 * rc;
    TAFwdWs f = ta_fwd_view(fwd_ws, C, E);
    TABwdWs b;
    b.ds_mean = (float *)bwd_ws;
    b.ds_max = (float *)((char *)bwd_ws + align_up((size_t)E * E * 4));
    TAPlanes x = {{xy, xz, yz}};
    TAPlanes g = {{g_out_xy, g_out_xz, g_out_yz}};
    TAPlanesOut gx = {{g_xy, g_xz, g_yz}};
    SPLATCO_CHECK_CUDA(cudaMemsetAsync(&f.sm->d_ca[0], 0, sizeof(float) * (TA_MAXC + 2 * TA_K * TA_K + 2 * TA_MAXC), st));
    const dim3 grid(ceil_div(E, TA_T), ceil_div(E, TA_T));
    ta_bwd_spatial_kernel<<<grid, TA_T * TA_T, 0, st>>>(x, g, C, rc, E, w_sa, f.sm, f.sa, f.smean, f.smax, f.arg, b.ds_mean, b.ds_max);
    SPLATCO_CHECK_LAUNCH();
    ta_bwd_channel_kernel<<<1, 256, 0, st>>>(C, hidden, w_ca1, w_ca2, f.sm, g_w_ca1, g_w_ca2, g_w_sa);
    SPLATCO_CHECK_LAUNCH();
    ta_bwd_apply_kernel<<<min(148 * 8, ceil_div(E * E, 256)), 256, 0, st>>>(g, gx, C, rc, E, f.sm, f.sa, f.arg, b.ds_mean, b.ds_max);
    SPLATCO_CHECK_LAUNCH();
    return 0;
}
