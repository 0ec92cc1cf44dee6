// C-ABI plumbing: error state, workspace layouts.  See include/splatco_b200.h.
#include <stdarg.h>

#include "common.cuh"

namespace splatco {

static thread_local char g_err[512] = "";

unsigned long long launch_count();

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static unsigned long long g_launches = 0;
void count_launch(int n) { __atomic_fetch_add(&g_launches, (unsigned long long)n, __ATOMIC_RELAXED); }
unsigned long long launch_count() { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

size_t geom_offsets(int P, size_t off[7]) {
    const size_t p = (size_t)(P > 0 ? P : 0);
    const size_t nb = (p + PRE_THREADS - 1) / PRE_THREADS;
    size_t o = 0;
    off[0] = o; o += align_up(p * 3 * sizeof(float4));
    off[1] = o; o += align_up(p * sizeof(float));
    off[2] = o; o += align_up(p * sizeof(uint32_t));
    off[3] = o; o += align_up(nb * sizeof(uint32_t));
    off[4] = o; o += align_up(nb * sizeof(uint32_t));
    off[5] = o; o += align_up(sizeof(uint32_t));
    off[6] = o;
    return o;
}

size_t bin_offsets(int64_t R, size_t off[7]) {
    const size_t r = (size_t)(R > 0 ? R : 0);
    const size_t nsb = (r + SORT_TILE - 1) / SORT_TILE;
    size_t o = 0;
    off[0] = o; o += align_up(r * sizeof(uint64_t));
    off[1] = o; o += align_up(r * sizeof(uint64_t));
    off[2] = o; o += align_up(r * sizeof(uint32_t));
    off[3] = o; o += align_up(r * sizeof(uint32_t));
    off[4] = o; o += align_up(256 * nsb * sizeof(uint32_t));
    off[5] = o; o += align_up(256 * sizeof(uint32_t));
    off[6] = o;
    return o;
}

size_t img_offsets(int H, int W, size_t off[9]) {
    const size_t T = (size_t)ceil_div(W, TILE) * ceil_div(H, TILE);
    const size_t hw = (size_t)H * W;
    size_t o = 0;
    off[0] = o; o += align_up(T * sizeof(int2));
    off[1] = o; o += align_up(hw * sizeof(float));
    off[2] = o; o += align_up(hw * sizeof(int32_t));
    off[3] = o; o += align_up(T * sizeof(uint32_t));      // tile_count
    off[4] = o; o += align_up(T * sizeof(uint32_t));      // cursor
    off[5] = o; o += align_up(3 * T * sizeof(uint32_t));  // per-size-class tile lists of the per-tile sort
    off[6] = o; o += align_up(8 * sizeof(uint32_t));      // list lengths / queue heads
    off[7] = o; o += align_up(T * sizeof(uint32_t));      // blend launch order (longest tiles first)
    off[8] = o;
    return o;
}

GeomWs geom_view(void *base, int P) {
    size_t off[7];
    geom_offsets(P, off);
    char *b = (char *)base;
    GeomWs g;
    g.rec = (float4 *)(b + off[0]);
    g.depths = (float *)(b + off[1]);
    g.tiles = (uint32_t *)(b + off[2]);
    g.block_sums = (uint32_t *)(b + off[3]);
    g.block_offsets = (uint32_t *)(b + off[4]);
    g.total = (uint32_t *)(b + off[5]);
    return g;
}

BinWs bin_view(void *base, int64_t R) {
    size_t off[7];
    bin_offsets(R, off);
    char *b = (char *)base;
    BinWs w;
    w.keys[0] = (uint64_t *)(b + off[0]);
    w.keys[1] = (uint64_t *)(b + off[1]);
    w.vals[0] = (uint32_t *)(b + off[2]);
    w.vals[1] = (uint32_t *)(b + off[3]);
    w.hist = (uint32_t *)(b + off[4]);
    w.bin_totals = (uint32_t *)(b + off[5]);
    return w;
}

ImgWs img_view(void *base, int H, int W) {
    size_t off[9];
    img_offsets(H, W, off);
    char *b = (char *)base;
    ImgWs w;
    w.ranges = (int2 *)(b + off[0]);
    w.final_T = (float *)(b + off[1]);
    w.n_contrib = (int32_t *)(b + off[2]);
    w.tile_count = (uint32_t *)(b + off[3]);
    w.cursor = (uint32_t *)(b + off[4]);
    w.lists = (uint32_t *)(b + off[5]);
    w.work = (uint32_t *)(b + off[6]);
    w.order = (uint32_t *)(b + off[7]);
    return w;
}

}  // namespace splatco

using namespace splatco;

extern "C" int splatco_abi_version(void) { return SPLATCO_ABI_VERSION; }
extern "C" const char *splatco_last_error(void) { return g_err; }
extern "C" uint64_t splatco_launch_count(void) { return (uint64_t)launch_count(); }

extern "C" size_t splatco_geom_bytes(int P) { size_t off[7]; return geom_offsets(P, off); }
extern "C" size_t splatco_binning_bytes(int64_t R) { size_t off[7]; return bin_offsets(R, off); }
extern "C" size_t splatco_image_bytes(int H, int W) { size_t off[9]; return img_offsets(H, W, off); }

static int copy_layout(const size_t *src, int n, size_t *dst, int max_chunks) {
    int k = n < max_chunks ? n : max_chunks;
    for (int i = 0; i < k; ++i) dst[i] = src[i];
    return k;
}
extern "C" int splatco_geom_layout(int P, size_t *offsets, int max_chunks) {
    size_t off[7]; geom_offsets(P, off); return copy_layout(off, 6, offsets, max_chunks);
}
extern "C" int splatco_binning_layout(int64_t R, size_t *offsets, int max_chunks) {
    size_t off[7]; bin_offsets(R, off); return copy_layout(off, 6, offsets, max_chunks);
}
extern "C" int splatco_image_layout(int H, int W, size_t *offsets, int max_chunks) {
    size_t off[9]; img_offsets(H, W, off); return copy_layout(off, 8, offsets, max_chunks);
}
