/*
 * raster_oracle.c — CPU restatement of the tile-based differentiable Gaussian rasterizer that
 * SplatCo calls through `diff_gaussian_rasterization` (reference call sites:
 * gaussian_renderer/__init__.py:145-171 forward, :208-242 visible_filter).
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (splatco_b200/) never imports, links or executes anything in oracle/.
 *
 * PARITY UNPINNED for this file: the reference's rasterizer sources live in submodules.zip, which
 * is absent from /root/reference (.MISSING_LARGE_BLOBS:1), the reference ships no tests or golden
 * vectors, and the package is not installed.  The algorithm below restates the
 * published semantics of the Inria diff-gaussian-rasterization lineage with the Scaffold-GS
 * `visible_filter` addition (un-pinned dependency, environment.yml:26), as recorded in SURVEY.md
 * Appendix A.1-A.5.  It is cross-checked against an independent fp64 PyTorch autograd composite
 * (oracle/composite_torch.py) in tests/test_oracle_raster.py.
 *
 * Rounding contract for the integer-deciding chain (radii, rects, tiles_touched, keys): every
 * fp32 operation is an individually rounded IEEE mul/add/sub/div/sqrt in the exact order written
 * here (compile with -ffp-contract=off, no -ffast-math); the CUDA kernels use __fmul_rn /
 * __fadd_rn / __fsub_rn / __fdiv_rn / __fsqrt_rn in the same order, so integers agree bit for bit.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

/* ---- tiny pthread parallel-for (the image has no libgomp) ------------------------------------ */
typedef void (*range_fn)(int lo, int hi, void *ctx);
typedef struct { range_fn fn; void *ctx; int n, chunk; volatile int *next; } pf_job_t;
static int g_threads = 0;
void oracle_set_threads(int n) { g_threads = n; }
int oracle_get_threads(void) {
    if (g_threads > 0) return g_threads;
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}
static void *pf_worker(void *arg) {
    pf_job_t *j = (pf_job_t *)arg;
    for (;;) {
        int lo = __atomic_fetch_add((int *)j->next, j->chunk, __ATOMIC_RELAXED);
        if (lo >= j->n) break;
        int hi = lo + j->chunk; if (hi > j->n) hi = j->n;
        j->fn(lo, hi, j->ctx);
    }
    return NULL;
}
static void parallel_for(int n, int chunk, range_fn fn, void *ctx) {
    int nt = oracle_get_threads();
    if (nt > 256) nt = 256;
    volatile int next = 0;
    pf_job_t job = { fn, ctx, n, chunk < 1 ? 1 : chunk, &next };
    if (nt <= 1 || n <= chunk) { pf_worker(&job); return; }
    pthread_t th[256];
    int started = 0;
    for (int t = 0; t < nt - 1; ++t) if (pthread_create(&th[started], NULL, pf_worker, &job) == 0) ++started;
    pf_worker(&job);
    for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
}
static void atomic_add_f64(double *addr, double v) {
    uint64_t *p = (uint64_t *)addr, old = __atomic_load_n(p, __ATOMIC_RELAXED), nw;
    double d;
    do { memcpy(&d, &old, 8); d += v; memcpy(&nw, &d, 8); }
    while (!__atomic_compare_exchange_n(p, &old, nw, 1, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
}

#define BLOCK_X 16
#define BLOCK_Y 16

/* CUDA's cvt.rzi.s32.f32: truncate, saturate, NaN -> 0 (C's cast is UB out of range). */
static int f2i_rz_sat(float f) {
    if (f != f) return 0;
    if (f >= 2147483648.0f) return 2147483647;
    if (f <= -2147483648.0f) return (-2147483647 - 1);
    return (int)f;
}
static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

typedef struct {
    int radius;          /* 0 => culled / invisible */
    float depth;         /* view-space z */
    float px, py;        /* pixel-space mean */
    float cov3d[6];
    float conic[3];
    int rect[4];         /* min.x min.y max.x max.y (tile units, max exclusive) */
    uint32_t tiles;
} gs_proj_t;

/* Appendix A.2.  `need_rect` is 1 for the render preprocess, 1 as well for visible_filter (the
 * filter kernel is "the same up to the area test, writing only radii"). */
static void project_one(const float *p, const float *sc, const float *q, float mod,
                        const float *view, const float *proj, float tanfovx, float tanfovy,
                        float focal_x, float focal_y, int H, int W, int gx, int gy, gs_proj_t *o)
{
    memset(o, 0, sizeof(*o));
    const float x = p[0], y = p[1], z = p[2];
    /* transformPoint4x3 with the row-vector (transposed) tensors read column-major */
    const float vx = ((view[0] * x + view[4] * y) + view[8] * z) + view[12];
    const float vy = ((view[1] * x + view[5] * y) + view[9] * z) + view[13];
    const float vz = ((view[2] * x + view[6] * y) + view[10] * z) + view[14];
    if (vz <= 0.2f) return;
    const float hx = ((proj[0] * x + proj[4] * y) + proj[8] * z) + proj[12];
    const float hy = ((proj[1] * x + proj[5] * y) + proj[9] * z) + proj[13];
    const float hw = ((proj[3] * x + proj[7] * y) + proj[11] * z) + proj[15];
    const float pw = 1.0f / (hw + 0.0000001f);
    const float ndc_x = hx * pw, ndc_y = hy * pw;

    /* cov3D = R S^2 R^T, quaternion (r,x,y,z) used as given (no renormalisation) */
    const float s0 = mod * sc[0], s1 = mod * sc[1], s2 = mod * sc[2];
    const float qr = q[0], qx = q[1], qy = q[2], qz = q[3];
    float R[3][3];
    R[0][0] = 1.0f - 2.0f * (qy * qy + qz * qz);
    R[0][1] = 2.0f * (qx * qy - qr * qz);
    R[0][2] = 2.0f * (qx * qz + qr * qy);
    R[1][0] = 2.0f * (qx * qy + qr * qz);
    R[1][1] = 1.0f - 2.0f * (qx * qx + qz * qz);
    R[1][2] = 2.0f * (qy * qz - qr * qx);
    R[2][0] = 2.0f * (qx * qz - qr * qy);
    R[2][1] = 2.0f * (qy * qz + qr * qx);
    R[2][2] = 1.0f - 2.0f * (qx * qx + qy * qy);
    float M[3][3];
    for (int i = 0; i < 3; ++i) { M[i][0] = R[i][0] * s0; M[i][1] = R[i][1] * s1; M[i][2] = R[i][2] * s2; }
    float S[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = i; j < 3; ++j) {
            S[i][j] = (M[i][0] * M[j][0] + M[i][1] * M[j][1]) + M[i][2] * M[j][2];
            S[j][i] = S[i][j];
        }
    o->cov3d[0] = S[0][0]; o->cov3d[1] = S[0][1]; o->cov3d[2] = S[0][2];
    o->cov3d[3] = S[1][1]; o->cov3d[4] = S[1][2]; o->cov3d[5] = S[2][2];

    /* EWA cov2D = (J Rot) Sigma (J Rot)^T, +0.3 dilation */
    const float limx = 1.3f * tanfovx, limy = 1.3f * tanfovy;
    const float txtz = vx / vz, tytz = vy / vz;
    const float tx = fminf(limx, fmaxf(-limx, txtz)) * vz;
    const float ty = fminf(limy, fmaxf(-limy, tytz)) * vz;
    const float J00 = focal_x / vz;
    const float J02 = -(focal_x * tx) / (vz * vz);
    const float J11 = focal_y / vz;
    const float J12 = -(focal_y * ty) / (vz * vz);
    float A0[3], A1[3];
    for (int k = 0; k < 3; ++k) {
        const float r0 = view[4 * k + 0], r1 = view[4 * k + 1], r2 = view[4 * k + 2];
        A0[k] = J00 * r0 + J02 * r2;
        A1[k] = J11 * r1 + J12 * r2;
    }
    float B0[3], B1[3];
    for (int l = 0; l < 3; ++l) {
        B0[l] = (A0[0] * S[0][l] + A0[1] * S[1][l]) + A0[2] * S[2][l];
        B1[l] = (A1[0] * S[0][l] + A1[1] * S[1][l]) + A1[2] * S[2][l];
    }
    const float a = ((B0[0] * A0[0] + B0[1] * A0[1]) + B0[2] * A0[2]) + 0.3f;
    const float b = (B0[0] * A1[0] + B0[1] * A1[1]) + B0[2] * A1[2];
    const float c = ((B1[0] * A1[0] + B1[1] * A1[1]) + B1[2] * A1[2]) + 0.3f;
    const float det = a * c - b * b;
    if (det == 0.0f) return;
    const float det_inv = 1.0f / det;
    o->conic[0] = c * det_inv; o->conic[1] = -b * det_inv; o->conic[2] = a * det_inv;
    const float mid = 0.5f * (a + c);
    const float sq = sqrtf(fmaxf(0.1f, mid * mid - det));
    const float l1 = mid + sq, l2 = mid - sq;
    const float rad_f = ceilf(3.0f * sqrtf(fmaxf(l1, l2)));
    const int radius = f2i_rz_sat(rad_f);
    const float px = ((ndc_x + 1.0f) * (float)W - 1.0f) * 0.5f;
    const float py = ((ndc_y + 1.0f) * (float)H - 1.0f) * 0.5f;
    const float rf = (float)radius;
    int r0x = imin(gx, imax(0, f2i_rz_sat((px - rf) / (float)BLOCK_X)));
    int r0y = imin(gy, imax(0, f2i_rz_sat((py - rf) / (float)BLOCK_Y)));
    int r1x = imin(gx, imax(0, f2i_rz_sat((px + rf + (float)(BLOCK_X - 1)) / (float)BLOCK_X)));
    int r1y = imin(gy, imax(0, f2i_rz_sat((py + rf + (float)(BLOCK_Y - 1)) / (float)BLOCK_Y)));
    const int area = (r1x - r0x) * (r1y - r0y);
    if (area == 0) return;
    o->radius = radius; o->depth = vz; o->px = px; o->py = py;
    o->rect[0] = r0x; o->rect[1] = r0y; o->rect[2] = r1x; o->rect[3] = r1y;
    o->tiles = (uint32_t)area;
}

typedef struct {
    const float *means3D, *scales, *rots, *opacities, *view, *proj;
    int scale_stride, H, W, filter_only; float scale_mod, tanfovx, tanfovy;
    int32_t *radii, *rect; float *xy, *depths, *cov3d, *conic_opacity; uint32_t *tiles_touched;
} pre_ctx_t;

static void pre_range(int lo, int hi, void *vc)
{
    pre_ctx_t *c = (pre_ctx_t *)vc;
    const int H = c->H, W = c->W;
    const float fx = (float)W / (2.0f * c->tanfovx), fy = (float)H / (2.0f * c->tanfovy);
    const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
    for (int i = lo; i < hi; ++i) {
        gs_proj_t o;
        project_one(c->means3D + 3 * (size_t)i, c->scales + (size_t)c->scale_stride * i, c->rots + 4 * (size_t)i,
                    c->scale_mod, c->view, c->proj, c->tanfovx, c->tanfovy, fx, fy, H, W, gx, gy, &o);
        c->radii[i] = o.radius;
        if (c->filter_only) continue;
        c->tiles_touched[i] = o.tiles;
        c->depths[i] = o.radius > 0 ? o.depth : 0.0f;
        c->xy[2 * (size_t)i] = o.radius > 0 ? o.px : 0.0f; c->xy[2 * (size_t)i + 1] = o.radius > 0 ? o.py : 0.0f;
        for (int k = 0; k < 6; ++k) c->cov3d[6 * (size_t)i + k] = o.cov3d[k];
        for (int k = 0; k < 3; ++k) c->conic_opacity[4 * (size_t)i + k] = o.radius > 0 ? o.conic[k] : 0.0f;
        c->conic_opacity[4 * (size_t)i + 3] = o.radius > 0 ? c->opacities[i] : 0.0f;
        for (int k = 0; k < 4; ++k) c->rect[4 * (size_t)i + k] = o.radius > 0 ? o.rect[k] : 0;
    }
}

/* visible_filter (Scaffold-GS addition; reference call gaussian_renderer/__init__.py:239-242).
 * `scales` may be a strided slice ([:, :3] of an [N,6] tensor) => scale_stride in floats. */
int oracle_visible_filter(int N, const float *means3D, const float *scales, int scale_stride,
                          const float *rots, float scale_mod, const float *view, const float *proj,
                          float tanfovx, float tanfovy, int H, int W, int32_t *radii)
{
    pre_ctx_t c; memset(&c, 0, sizeof(c));
    c.means3D = means3D; c.scales = scales; c.scale_stride = scale_stride; c.rots = rots;
    c.scale_mod = scale_mod; c.view = view; c.proj = proj; c.tanfovx = tanfovx; c.tanfovy = tanfovy;
    c.H = H; c.W = W; c.radii = radii; c.filter_only = 1;
    parallel_for(N, 4096, pre_range, &c);
    return 0;
}

/* preprocess forward (Appendix A.2).  Outputs (all length P unless noted):
 * radii i32, xy f32[2P], depths f32, cov3d f32[6P], conic_opacity f32[4P], rect i32[4P],
 * tiles_touched u32.  Culled Gaussians get radii = tiles = 0 and zeros elsewhere. */
int oracle_preprocess(int P, const float *means3D, const float *scales, int scale_stride,
                      const float *rots, const float *opacities, float scale_mod,
                      const float *view, const float *proj, float tanfovx, float tanfovy,
                      int H, int W, int32_t *radii, float *xy, float *depths, float *cov3d,
                      float *conic_opacity, int32_t *rect, uint32_t *tiles_touched)
{
    pre_ctx_t c; memset(&c, 0, sizeof(c));
    c.means3D = means3D; c.scales = scales; c.scale_stride = scale_stride; c.rots = rots;
    c.opacities = opacities; c.scale_mod = scale_mod; c.view = view; c.proj = proj;
    c.tanfovx = tanfovx; c.tanfovy = tanfovy; c.H = H; c.W = W; c.radii = radii; c.rect = rect;
    c.xy = xy; c.depths = depths; c.cov3d = cov3d; c.conic_opacity = conic_opacity;
    c.tiles_touched = tiles_touched; c.filter_only = 0;
    parallel_for(P, 4096, pre_range, &c);
    return 0;
}

/* Appendix A.3: total number of (tile, Gaussian) instances. */
int64_t oracle_num_rendered(int P, const uint32_t *tiles_touched)
{
    int64_t r = 0;
    for (int i = 0; i < P; ++i) r += tiles_touched[i];
    return r;
}

static uint32_t f32_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

/* Appendix A.3: emit keys (tile<<32 | depth bits) in (y outer, x inner) order per Gaussian,
 * stable-sort them, and derive per-tile [start,end).  keys_unsorted may be NULL.
 * A stable LSD byte radix sort over all 64 bits orders identically to the reference's stable sort
 * restricted to the low 32+msb(T) bits (the remaining high bits are zero). */
int oracle_binning(int P, const int32_t *radii, const int32_t *rect, const float *depths,
                   const uint32_t *tiles_touched, int grid_x, int grid_y, int64_t R,
                   uint64_t *keys_unsorted, uint64_t *keys_sorted, uint32_t *vals_sorted,
                   int32_t *ranges /* 2*T */)
{
    const int T = grid_x * grid_y;
    uint64_t *ka = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(R > 0 ? R : 1));
    uint64_t *kb = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(R > 0 ? R : 1));
    uint32_t *va = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(R > 0 ? R : 1));
    uint32_t *vb = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(R > 0 ? R : 1));
    if (!ka || !kb || !va || !vb) { free(ka); free(kb); free(va); free(vb); return -1; }
    int64_t off = 0;
    for (int i = 0; i < P; ++i) {
        if (radii[i] <= 0) continue;
        const int32_t *rc = rect + 4 * (size_t)i;
        for (int y = rc[1]; y < rc[3]; ++y)
            for (int x = rc[0]; x < rc[2]; ++x) {
                uint64_t key = (uint64_t)(uint32_t)(y * grid_x + x);
                key = (key << 32) | f32_bits(depths[i]);
                ka[off] = key; va[off] = (uint32_t)i; ++off;
            }
        (void)tiles_touched;
    }
    if (off != R) { free(ka); free(kb); free(va); free(vb); return -2; }
    if (keys_unsorted) memcpy(keys_unsorted, ka, sizeof(uint64_t) * (size_t)R);
    for (int pass = 0; pass < 8; ++pass) {
        size_t cnt[257]; memset(cnt, 0, sizeof(cnt));
        const int sh = pass * 8;
        for (int64_t i = 0; i < R; ++i) cnt[((ka[i] >> sh) & 255) + 1]++;
        for (int d = 0; d < 256; ++d) cnt[d + 1] += cnt[d];
        for (int64_t i = 0; i < R; ++i) {
            size_t dst = cnt[(ka[i] >> sh) & 255]++;
            kb[dst] = ka[i]; vb[dst] = va[i];
        }
        uint64_t *tk = ka; ka = kb; kb = tk;
        uint32_t *tv = va; va = vb; vb = tv;
    }
    memcpy(keys_sorted, ka, sizeof(uint64_t) * (size_t)R);
    memcpy(vals_sorted, va, sizeof(uint32_t) * (size_t)R);
    memset(ranges, 0, sizeof(int32_t) * 2 * (size_t)T);
    for (int64_t i = 0; i < R; ++i) {
        const uint32_t t = (uint32_t)(ka[i] >> 32);
        if (i == 0 || (uint32_t)(ka[i - 1] >> 32) != t) ranges[2 * t] = (int32_t)i;
        if (i == R - 1 || (uint32_t)(ka[i + 1] >> 32) != t) ranges[2 * t + 1] = (int32_t)(i + 1);
    }
    free(ka); free(kb); free(va); free(vb);
    return 0;
}

/* Appendix A.4: per-pixel front-to-back blend.  `fragile` (may be NULL) flags pixels where some
 * evaluated (pixel, Gaussian) pair sits within rel. 1e-4 of a discontinuous threshold
 * (alpha = 1/255, T = 1e-4, power = 0): there a 1-ulp difference in exp() legitimately changes the
 * pixel by up to ~0.4 %, so parity tests compare those pixels with the looser stated bound. */
typedef struct {
    int P, H, W; const int32_t *ranges; const uint32_t *point_list;
    const float *xy, *conic_opacity, *colors, *bg, *final_T_in, *dL_dpix; const int32_t *n_contrib_in;
    float *out_color, *final_T; int32_t *n_contrib; uint8_t *fragile; double *acc;
} blend_ctx_t;

static void blend_fwd_range(int lo, int hi, void *vc)
{
    blend_ctx_t *c = (blend_ctx_t *)vc;
    const int H = c->H, W = c->W; const int32_t *ranges = c->ranges; const uint32_t *point_list = c->point_list;
    const float *xy = c->xy, *conic_opacity = c->conic_opacity, *colors = c->colors, *bg = c->bg;
    float *out_color = c->out_color, *final_T = c->final_T; int32_t *n_contrib = c->n_contrib; uint8_t *fragile = c->fragile;
    const int gx = (W + BLOCK_X - 1) / BLOCK_X;
    for (int py = lo; py < hi; ++py) {
        for (int px = 0; px < W; ++px) {
            const int tile = (py / BLOCK_Y) * gx + (px / BLOCK_X);
            const int32_t beg = ranges[2 * tile], end = ranges[2 * tile + 1];
            const float pxf = (float)px, pyf = (float)py;
            float T = 1.0f, C0 = 0.f, C1 = 0.f, C2 = 0.f;
            int contributor = 0, last = 0; uint8_t frag = 0;
            for (int32_t k = beg; k < end; ++k) {
                ++contributor;
                const uint32_t g = point_list[k];
                const float dx = xy[2 * g] - pxf, dy = xy[2 * g + 1] - pyf;
                const float *co = conic_opacity + 4 * (size_t)g;
                const float power = -0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                if (fabsf(power) < 1e-6f) frag = 1;
                if (power > 0.0f) continue;
                const float ea = co[3] * expf(power);
                const float alpha = fminf(0.99f, ea);
                if (fabsf(alpha * 255.0f - 1.0f) < 1e-4f) frag = 1;
                if (alpha < 1.0f / 255.0f) continue;
                const float test_T = T * (1.0f - alpha);
                if (fabsf(test_T * 1e4f - 1.0f) < 1e-4f) frag = 1;
                if (test_T < 0.0001f) break;
                const float w = alpha * T;
                C0 += colors[3 * (size_t)g] * w; C1 += colors[3 * (size_t)g + 1] * w; C2 += colors[3 * (size_t)g + 2] * w;
                T = test_T; last = contributor;
            }
            const size_t pid = (size_t)py * W + px, HW = (size_t)H * W;
            final_T[pid] = T; n_contrib[pid] = last;
            out_color[pid] = C0 + T * bg[0];
            out_color[HW + pid] = C1 + T * bg[1];
            out_color[2 * HW + pid] = C2 + T * bg[2];
            if (fragile) fragile[pid] = frag;
        }
    }
}

int oracle_blend_fwd(int H, int W, const int32_t *ranges, const uint32_t *point_list,
                     const float *xy, const float *conic_opacity, const float *colors,
                     const float *bg, float *out_color /* 3*H*W CHW */, float *final_T,
                     int32_t *n_contrib, uint8_t *fragile)
{
    blend_ctx_t c; memset(&c, 0, sizeof(c));
    c.H = H; c.W = W; c.ranges = ranges; c.point_list = point_list; c.xy = xy; c.conic_opacity = conic_opacity;
    c.colors = colors; c.bg = bg; c.out_color = out_color; c.final_T = final_T; c.n_contrib = n_contrib; c.fragile = fragile;
    parallel_for(H, 4, blend_fwd_range, &c);
    return 0;
}

/* Appendix A.5 (first half): back-to-front replay.  Accumulates in fp64 (the oracle is the
 * "truth" the fp32-atomic GPU sums are compared against, 1e-3 relative).
 * Outputs per Gaussian: dL_dmean2D[3P] (x,y in NDC-scaled units: includes 0.5*W / 0.5*H, z=0),
 * dL_dconic[3P] (conic.x, conic.y, conic.z == upstream's .x .y .w), dL_dopacity[P], dL_dcolor[3P]. */
static void blend_bwd_range(int lo, int hi, void *vc)
{
    blend_ctx_t *c = (blend_ctx_t *)vc;
    const int H = c->H, W = c->W; const int32_t *ranges = c->ranges; const uint32_t *point_list = c->point_list;
    const float *xy = c->xy, *conic_opacity = c->conic_opacity, *colors = c->colors, *bg = c->bg;
    const float *final_T = c->final_T_in, *dL_dpix = c->dL_dpix; const int32_t *n_contrib = c->n_contrib_in;
    double *acc = c->acc;
    const int gx = (W + BLOCK_X - 1) / BLOCK_X;
    const size_t HW = (size_t)H * W;
    const float ddelx_dx = 0.5f * (float)W, ddely_dy = 0.5f * (float)H;
    for (int py = lo; py < hi; ++py) {
        for (int px = 0; px < W; ++px) {
            const int tile = (py / BLOCK_Y) * gx + (px / BLOCK_X);
            const int32_t beg = ranges[2 * tile];
            const size_t pid = (size_t)py * W + px;
            const float pxf = (float)px, pyf = (float)py;
            const float T_final = final_T[pid];
            float T = T_final;
            const int last = n_contrib[pid];
            const float dp0 = dL_dpix[pid], dp1 = dL_dpix[HW + pid], dp2 = dL_dpix[2 * HW + pid];
            const float bg_dot = bg[0] * dp0 + bg[1] * dp1 + bg[2] * dp2;
            float ar0 = 0.f, ar1 = 0.f, ar2 = 0.f, lc0 = 0.f, lc1 = 0.f, lc2 = 0.f, last_alpha = 0.f;
            for (int32_t k = beg + last - 1; k >= beg; --k) {
                const uint32_t g = point_list[k];
                const float dx = xy[2 * g] - pxf, dy = xy[2 * g + 1] - pyf;
                const float *co = conic_opacity + 4 * (size_t)g;
                const float power = -0.5f * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                if (power > 0.0f) continue;
                const float G = expf(power);
                const float alpha = fminf(0.99f, co[3] * G);
                if (alpha < 1.0f / 255.0f) continue;
                T = T / (1.0f - alpha);
                const float dch = alpha * T;
                const float c0 = colors[3 * (size_t)g], c1 = colors[3 * (size_t)g + 1], c2 = colors[3 * (size_t)g + 2];
                ar0 = last_alpha * lc0 + (1.f - last_alpha) * ar0; lc0 = c0;
                ar1 = last_alpha * lc1 + (1.f - last_alpha) * ar1; lc1 = c1;
                ar2 = last_alpha * lc2 + (1.f - last_alpha) * ar2; lc2 = c2;
                float dL_dalpha = (c0 - ar0) * dp0 + (c1 - ar1) * dp1 + (c2 - ar2) * dp2;
                dL_dalpha *= T;
                last_alpha = alpha;
                dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot;
                const float dL_dG = co[3] * dL_dalpha;
                const float gdx = G * dx, gdy = G * dy;
                const float dG_ddelx = -gdx * co[0] - gdy * co[1];
                const float dG_ddely = -gdy * co[2] - gdx * co[1];
                double *a = acc + 9 * (size_t)g;
                const double v[9] = {
                    (double)(dL_dG * dG_ddelx * ddelx_dx), (double)(dL_dG * dG_ddely * ddely_dy),
                    (double)(-0.5f * gdx * dx * dL_dG), (double)(-0.5f * gdx * dy * dL_dG),
                    (double)(-0.5f * gdy * dy * dL_dG), (double)(G * dL_dalpha),
                    (double)(dch * dp0), (double)(dch * dp1), (double)(dch * dp2) };
                for (int q = 0; q < 9; ++q) atomic_add_f64(a + q, v[q]);
            }
        }
    }
}

int oracle_blend_bwd(int P, int H, int W, const int32_t *ranges, const uint32_t *point_list,
                     const float *xy, const float *conic_opacity, const float *colors,
                     const float *bg, const float *final_T, const int32_t *n_contrib,
                     const float *dL_dpix /* 3*H*W */, float *dL_dmean2D, float *dL_dconic,
                     float *dL_dopacity, float *dL_dcolor)
{
    double *acc = (double *)calloc((size_t)P * 9 + 1, sizeof(double));
    if (!acc) return -1;
    blend_ctx_t c; memset(&c, 0, sizeof(c));
    c.P = P; c.H = H; c.W = W; c.ranges = ranges; c.point_list = point_list; c.xy = xy; c.conic_opacity = conic_opacity;
    c.colors = colors; c.bg = bg; c.final_T_in = final_T; c.n_contrib_in = n_contrib; c.dL_dpix = dL_dpix; c.acc = acc;
    parallel_for(H, 4, blend_bwd_range, &c);
    for (int i = 0; i < P; ++i) {
        const double *a = acc + 9 * (size_t)i;
        dL_dmean2D[3 * (size_t)i] = (float)a[0]; dL_dmean2D[3 * (size_t)i + 1] = (float)a[1]; dL_dmean2D[3 * (size_t)i + 2] = 0.f;
        dL_dconic[3 * (size_t)i] = (float)a[2]; dL_dconic[3 * (size_t)i + 1] = (float)a[3]; dL_dconic[3 * (size_t)i + 2] = (float)a[4];
        dL_dopacity[i] = (float)a[5];
        dL_dcolor[3 * (size_t)i] = (float)a[6]; dL_dcolor[3 * (size_t)i + 1] = (float)a[7]; dL_dcolor[3 * (size_t)i + 2] = (float)a[8];
    }
    free(acc);
    return 0;
}

/* Appendix A.5 (second half): per-Gaussian chain conic -> cov2D -> (mean3D, cov3D) -> (scale, quat)
 * and mean2D -> mean3D through the projection.  Derivation (own notation; validated against fp64
 * autograd in tests): cov2D = A Sigma A^T + 0.3 I with A = J Rot (2x3), conic = cov2D^-1.
 * Keeps two reference quirks: the conic gradient uses 1/(det^2 + 1e-7), and the x/y gradient
 * through the clamped t.x/t.z, t.y/t.z is zeroed where the clamp was active. */
typedef struct {
    const float *means3D, *scales, *rots, *view, *proj, *dL_dmean2D, *dL_dconic; const int32_t *radii;
    int scale_stride, H, W; float scale_mod, tanfovx, tanfovy; float *dL_dmeans3D, *dL_dscales, *dL_drots;
} pbwd_ctx_t;

static void pbwd_range(int lo, int hi, void *vc)
{
    pbwd_ctx_t *c = (pbwd_ctx_t *)vc;
    const float *means3D = c->means3D, *scales = c->scales, *rots = c->rots, *view = c->view, *proj = c->proj;
    const float *dL_dmean2D = c->dL_dmean2D, *dL_dconic = c->dL_dconic; const int32_t *radii = c->radii;
    const int scale_stride = c->scale_stride, H = c->H, W = c->W;
    const float scale_mod = c->scale_mod, tanfovx = c->tanfovx, tanfovy = c->tanfovy;
    float *dL_dmeans3D = c->dL_dmeans3D, *dL_dscales = c->dL_dscales, *dL_drots = c->dL_drots;
    const float fx = (float)W / (2.0f * tanfovx), fy = (float)H / (2.0f * tanfovy);
    for (int i = lo; i < hi; ++i) {
        float *gm = dL_dmeans3D + 3 * (size_t)i, *gs = dL_dscales + 3 * (size_t)i, *gq = dL_drots + 4 * (size_t)i;
        gm[0] = gm[1] = gm[2] = 0.f; gs[0] = gs[1] = gs[2] = 0.f; gq[0] = gq[1] = gq[2] = gq[3] = 0.f;
        if (!(radii[i] > 0)) continue;
        const float *p = means3D + 3 * (size_t)i, *sc = scales + (size_t)scale_stride * i, *q = rots + 4 * (size_t)i;
        const double x = p[0], y = p[1], z = p[2];
        double Rot[3][3], tr[3];
        for (int r = 0; r < 3; ++r) { for (int k = 0; k < 3; ++k) Rot[r][k] = view[4 * k + r]; tr[r] = view[12 + r]; }
        const double vx = Rot[0][0] * x + Rot[0][1] * y + Rot[0][2] * z + tr[0];
        const double vy = Rot[1][0] * x + Rot[1][1] * y + Rot[1][2] * z + tr[1];
        const double vz = Rot[2][0] * x + Rot[2][1] * y + Rot[2][2] * z + tr[2];
        /* forward recompute (fp64 inside the oracle's backward; inputs are the fp32 tensors) */
        const double s[3] = { (double)scale_mod * sc[0], (double)scale_mod * sc[1], (double)scale_mod * sc[2] };
        const double qr = q[0], qx = q[1], qy = q[2], qz = q[3];
        double R[3][3] = {
            { 1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - qr * qz), 2 * (qx * qz + qr * qy) },
            { 2 * (qx * qy + qr * qz), 1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - qr * qx) },
            { 2 * (qx * qz - qr * qy), 2 * (qy * qz + qr * qx), 1 - 2 * (qx * qx + qy * qy) } };
        double M[3][3], S[3][3];
        for (int a = 0; a < 3; ++a) for (int k = 0; k < 3; ++k) M[a][k] = R[a][k] * s[k];
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) S[a][b] = M[a][0] * M[b][0] + M[a][1] * M[b][1] + M[a][2] * M[b][2];
        const double limx = 1.3 * (double)tanfovx, limy = 1.3 * (double)tanfovy;
        const double txtz = vx / vz, tytz = vy / vz;
        const int clx = (txtz < -limx || txtz > limx), cly = (tytz < -limy || tytz > limy);
        const double tx = fmin(limx, fmax(-limx, txtz)) * vz, ty = fmin(limy, fmax(-limy, tytz)) * vz;
        const double J00 = fx / vz, J02 = -(fx * tx) / (vz * vz), J11 = fy / vz, J12 = -(fy * ty) / (vz * vz);
        double A[2][3];
        for (int k = 0; k < 3; ++k) { A[0][k] = J00 * Rot[0][k] + J02 * Rot[2][k]; A[1][k] = J11 * Rot[1][k] + J12 * Rot[2][k]; }
        double B[2][3];
        for (int r = 0; r < 2; ++r) for (int l = 0; l < 3; ++l) B[r][l] = A[r][0] * S[0][l] + A[r][1] * S[1][l] + A[r][2] * S[2][l];
        const double a = B[0][0] * A[0][0] + B[0][1] * A[0][1] + B[0][2] * A[0][2] + 0.3;
        const double b = B[0][0] * A[1][0] + B[0][1] * A[1][1] + B[0][2] * A[1][2];
        const double c = B[1][0] * A[1][0] + B[1][1] * A[1][1] + B[1][2] * A[1][2] + 0.3;
        const double denom = a * c - b * b;
        const double d2inv = 1.0 / (denom * denom + 0.0000001);
        const double gc0 = dL_dconic[3 * (size_t)i], gc1 = dL_dconic[3 * (size_t)i + 1], gc2 = dL_dconic[3 * (size_t)i + 2];
        /* conic = (c, -b, a)/denom */
        const double dL_da = d2inv * (-c * c * gc0 + 2 * b * c * gc1 + (denom - a * c) * gc2);
        const double dL_dc = d2inv * (-a * a * gc2 + 2 * a * b * gc1 + (denom - a * c) * gc0);
        const double dL_db = d2inv * 2 * (b * c * gc0 - (denom + 2 * b * b) * gc1 + a * b * gc2);
        /* G2 = dL/dcov2D as a symmetric 2x2 with the off-diagonal split in halves */
        const double G2[2][2] = { { dL_da, 0.5 * dL_db }, { 0.5 * dL_db, dL_dc } };
        /* dL/dSigma = A^T G2 A (symmetric 3x3);  dL/dA = 2 G2 A Sigma = 2 G2 B */
        double dS[3][3], dA[2][3];
        for (int k = 0; k < 3; ++k) for (int l = 0; l < 3; ++l) {
            double v = 0; for (int r = 0; r < 2; ++r) for (int t = 0; t < 2; ++t) v += A[r][k] * G2[r][t] * A[t][l];
            dS[k][l] = v;
        }
        for (int r = 0; r < 2; ++r) for (int l = 0; l < 3; ++l) dA[r][l] = 2 * (G2[r][0] * B[0][l] + G2[r][1] * B[1][l]);
        /* A0k = J00 Rot0k + J02 Rot2k ; A1k = J11 Rot1k + J12 Rot2k */
        double dJ00 = 0, dJ02 = 0, dJ11 = 0, dJ12 = 0;
        for (int k = 0; k < 3; ++k) { dJ00 += dA[0][k] * Rot[0][k]; dJ02 += dA[0][k] * Rot[2][k]; dJ11 += dA[1][k] * Rot[1][k]; dJ12 += dA[1][k] * Rot[2][k]; }
        const double iz = 1.0 / vz, iz2 = iz * iz, iz3 = iz2 * iz;
        const double dtx = clx ? 0.0 : -fx * iz2 * dJ02;
        const double dty = cly ? 0.0 : -fy * iz2 * dJ12;
        const double dtz = -fx * iz2 * dJ00 - fy * iz2 * dJ11 + 2 * fx * tx * iz3 * dJ02 + 2 * fy * ty * iz3 * dJ12;
        /* view-space -> world: dL/dp = Rot^T dL/dt */
        double gmx = Rot[0][0] * dtx + Rot[1][0] * dty + Rot[2][0] * dtz;
        double gmy = Rot[0][1] * dtx + Rot[1][1] * dty + Rot[2][1] * dtz;
        double gmz = Rot[0][2] * dtx + Rot[1][2] * dty + Rot[2][2] * dtz;
        /* mean2D -> mean3D through the perspective divide (pixel = ((ndc+1)S-1)/2; the 0.5*S factor is
         * already folded into dL_dmean2D by the blend backward) */
        const double hx = proj[0] * x + proj[4] * y + proj[8] * z + proj[12];
        const double hy = proj[1] * x + proj[5] * y + proj[9] * z + proj[13];
        const double hw = proj[3] * x + proj[7] * y + proj[11] * z + proj[15];
        const double mw = 1.0 / (hw + 0.0000001);
        const double mul1 = hx * mw * mw, mul2 = hy * mw * mw;
        const double g2x = dL_dmean2D[3 * (size_t)i], g2y = dL_dmean2D[3 * (size_t)i + 1];
        gmx += (proj[0] * mw - proj[3] * mul1) * g2x + (proj[1] * mw - proj[3] * mul2) * g2y;
        gmy += (proj[4] * mw - proj[7] * mul1) * g2x + (proj[5] * mw - proj[7] * mul2) * g2y;
        gmz += (proj[8] * mw - proj[11] * mul1) * g2x + (proj[9] * mw - proj[11] * mul2) * g2y;
        gm[0] = (float)gmx; gm[1] = (float)gmy; gm[2] = (float)gmz;
        /* Sigma = M M^T, M = R diag(s):  dL/dM = 2 dS M  (dS symmetric) */
        double dM[3][3];
        for (int a2 = 0; a2 < 3; ++a2) for (int k = 0; k < 3; ++k) dM[a2][k] = 2 * (dS[a2][0] * M[0][k] + dS[a2][1] * M[1][k] + dS[a2][2] * M[2][k]);
        /* M_ak = R_ak s_k */
        double dR[3][3];
        for (int k = 0; k < 3; ++k) {
            double v = 0; for (int a2 = 0; a2 < 3; ++a2) { v += dM[a2][k] * R[a2][k]; dR[a2][k] = dM[a2][k] * s[k]; }
            gs[k] = (float)(v * (double)scale_mod);
        }
        /* R(q) derivatives */
        const double dqr = 2 * (-qz * dR[0][1] + qy * dR[0][2] + qz * dR[1][0] - qx * dR[1][2] - qy * dR[2][0] + qx * dR[2][1]);
        const double dqx = 2 * (qy * dR[0][1] + qz * dR[0][2] + qy * dR[1][0] - 2 * qx * dR[1][1] - qr * dR[1][2] + qz * dR[2][0] + qr * dR[2][1] - 2 * qx * dR[2][2]);
        const double dqy = 2 * (-2 * qy * dR[0][0] + qx * dR[0][1] + qr * dR[0][2] + qx * dR[1][0] + qz * dR[1][2] - qr * dR[2][0] + qz * dR[2][1] - 2 * qy * dR[2][2]);
        const double dqz = 2 * (-2 * qz * dR[0][0] - qr * dR[0][1] + qx * dR[0][2] + qr * dR[1][0] - 2 * qz * dR[1][1] + qy * dR[1][2] + qx * dR[2][0] + qy * dR[2][1]);
        gq[0] = (float)dqr; gq[1] = (float)dqx; gq[2] = (float)dqy; gq[3] = (float)dqz;
    }
}

int oracle_preprocess_bwd(int P, const float *means3D, const float *scales, int scale_stride,
                          const float *rots, float scale_mod, const float *view, const float *proj,
                          float tanfovx, float tanfovy, int H, int W, const int32_t *radii,
                          const float *dL_dmean2D, const float *dL_dconic,
                          float *dL_dmeans3D, float *dL_dscales, float *dL_drots)
{
    pbwd_ctx_t c; memset(&c, 0, sizeof(c));
    c.means3D = means3D; c.scales = scales; c.rots = rots; c.view = view; c.proj = proj;
    c.dL_dmean2D = dL_dmean2D; c.dL_dconic = dL_dconic; c.radii = radii; c.scale_stride = scale_stride;
    c.H = H; c.W = W; c.scale_mod = scale_mod; c.tanfovx = tanfovx; c.tanfovy = tanfovy;
    c.dL_dmeans3D = dL_dmeans3D; c.dL_dscales = dL_dscales; c.dL_drots = dL_drots;
    parallel_for(P, 4096, pbwd_range, &c);
    return 0;
}

int oracle_abi_version(void) { return 1; }
