/*
 * k_vis.cu -- K4a, order-independent form: coverage + depth resolution of the tiles whose records all belong to
 * the batch's "unordered class" (RC_UNORDERED: deferrable colour work, depth test and depth write on, no stencil
 * test, one common depth function out of GL_LESS / GL_LEQUAL / GL_GREATER / GL_GEQUAL).
 *
 * Replaces the scan loop, edge test and depth test/write of rasterize_triangle_smooth (src/raster.c:532-548,
 * 582-587, 707-710) for those tiles.  The reference draws fragments in submission order; for this class the final
 * depth and the surviving fragment of a pixel are a pure function of the SET of fragments:
 *
 *     GL_LESS    : smallest depth, ties -> earliest submission (the later equal depth fails `<`), the value already
 *                  in the depth plane wins ties;
 *     GL_LEQUAL  : smallest depth, ties -> latest submission, a fragment beats an equal value in the plane;
 *     GL_GREATER / GL_GEQUAL : the same with the largest depth.
 *
 * So each pixel holds one 64-bit key  (order-preserving image of the depth) << 32 | tie-break word  in shared
 * memory, every fragment does an atomicMin on it in whatever order the hardware gets to it, and the result is
 * bit-identical to the in-order loop.  What that buys: no per-tile sort, no pixel ownership, no window limit, and
 * freedom to map work to lanes by triangle size --
 *
 *   small triangles (clamped box <= 256 pixels, the C4 case: ~14 covered pixels each): a warp fetches 32 record
 *     heads, one per lane, then works on four triangles at a time, 8 lanes each; a lane owns one column of the box
 *     and walks its rows with the column part of the three edge functions hoisted;
 *   large triangles: queued, then rasterised by all 8 warps together over 8x4 pixel blocks.
 *
 * The arithmetic is the reference's, operation for operation (edge_function raster.c:299-302, barycentrics 541-543,
 * depth 546-548); the only liberty is algebraically exact: for clockwise triangles all three edge values and
 * 1/area are negated together (IEEE negation is exact), which turns the two inclusive tests of raster.c:539-540
 * into one.
 *
 * The kernel leaves the depth plane and, per pixel, the record index of the surviving fragment (or VIS_NONE) in the
 * visibility plane for k_shade (K4b).  The depth / stencil part of the batch's leading glClear is fused here.
 *
 * Algorithmic bytes per tile: 64 B head per referenced record + 4 B list entry, 4 B/pixel depth in (unless
 * cleared), 4 B/pixel depth out, 4 B/pixel visibility out.
 */
#include "dev_common.cuh"

namespace mtgl_dev_impl {

void note_launch();

constexpr uint32_t VIS_NONE = 0xFFFFFFFFu;
constexpr int VIS_PITCH = 66;           /* 64-bit words per tile row: rows start 4 banks apart, 16-byte aligned */
constexpr int VIS_LARGE_CAP = 192;      /* capacity of the large-triangle queue (a full queue makes the finding warp do the triangle alone) */
constexpr int VIS_EXACT_EXTENT = 2047;  /* vertex extent up to which all edge values inside a tile are exactly represented integers */
constexpr int VIS_COORD_LIMIT = 1 << 22;
constexpr int VIS_SMALL_AREA = 1024;    /* clamped box area up to which one warp turns a triangle into spans (phase 1) ... */
constexpr int VIS_SMALL_AREA_BUSY = 4096;   /* ... any box, when the tile's list is long enough to keep all warps busy with whole triangles */
constexpr uint32_t VIS_BUSY_LIST = 64;

/* per warp: the prepared small triangles of the current chunk (slot = compacted position in the chunk) and the
 * row spans of the current group of 32 rows.  Edge k is e = A*x + B*y + C in tile-relative pixel coordinates. */
struct VisWarp {
    float4 T0[32];          /* A0 A1 A2 1/area                                   (pixel stage) */
    float4 T1[32];          /* z0 z1 z2 id                                       (pixel stage) */
    float4 T2[32];          /* B0 B1 B2 X0 | Y0 << 6 | (width - 1) << 12 | first-row-in-run << 18   (span stage) */
    float C[3][32];         /*                                                   (span stage) */
    float RN[3][32];        /* -1 / A_k, 0 when A_k == 0                         (span stage) */
    float4 S[32];           /* spans: row terms t_k = B_k*y + C_k, x_left | y << 6 | triangle slot << 12 | first-pixel-in-run << 17 */
};

template <int THREADS>
struct VisSmem {
    unsigned long long key[TILE_H * VIS_PITCH];
    uint32_t large_rec[VIS_LARGE_CAP];
    VisWarp w[THREADS / 32];
    uint32_t next_chunk;
    uint32_t large_n;
};
/* four 8-warp CTAs (or two 16-warp CTAs) per SM: n * (sizeof + 1 KB reserved) must fit the SM's 228 KB */
static_assert(4 * (sizeof(VisSmem<256>) + 1024) <= 228 * 1024, "k_vis: four tiles per SM");
static_assert(2 * (sizeof(VisSmem<512>) + 1024) <= 228 * 1024, "k_vis<512>: two tiles per SM");

/* monotone map float -> uint32 (total order of the reals, -0 < +0) and back */
__device__ __forceinline__ uint32_t ord_bits(float f)
{
    const uint32_t u = __float_as_uint(f);
    return u ^ ((uint32_t)((int32_t)u >> 31) | 0x80000000u);
}
__device__ __forceinline__ float ord_float(uint32_t k)
{
    return __uint_as_float((k & 0x80000000u) ? (k ^ 0x80000000u) : ~k);
}

struct VisMode {
    uint32_t inv;           /* 0xFFFFFFFF for GL_GREATER / GL_GEQUAL: the key holds the complement, so min = largest depth */
    bool first_wins;        /* GL_LESS / GL_GREATER: equal depths keep the earliest fragment and the plane's value */
    bool all_range01;       /* every unordered state has depth range [0,1] */
};

__device__ __forceinline__ unsigned long long make_key(const VisMode &m, float depth, uint32_t id)
{
    const uint32_t hi = ord_bits(depth) ^ m.inv;
    const uint32_t lo = m.first_wins ? id + 1u : 0xFFFFFFFEu - id;
    return ((unsigned long long)hi << 32) | lo;
}

__device__ __forceinline__ void key_min(unsigned long long *addr, unsigned long long key)
{
    /* keys only ever decrease, so a fragment that loses against a (possibly stale) read loses for good */
    if (key < *reinterpret_cast<volatile unsigned long long *>(addr)) atomicMin(addr, key);
}

/* window-space depth of raster.c:546-548 */
__device__ __forceinline__ float depth_of(float z, bool range01, double dnear, double dfar)
{
    const float d = (z + 1.0f) * 0.5f;
    if (range01) return d;
    return (float)((double)d * (dfar - dnear) + dnear);
}

struct VisHead {                /* the 64 leading bytes of a record: rows 0-3 */
    int4 row0, row1;            /* x0 y0 x1 y1 | x2 y2 area inv_area */
    uint4 row2;                 /* bbox_min bbox_max state_flags id */
    float z0, z1, z2;
};

__device__ __forceinline__ void load_vis_head(VisHead &h, const TriRecord *rec)
{
    h.row0 = __ldg(reinterpret_cast<const int4 *>(rec) + 0);
    h.row1 = __ldg(reinterpret_cast<const int4 *>(rec) + 1);
    h.row2 = __ldg(reinterpret_cast<const uint4 *>(rec) + 2);
    const float4 row3 = __ldg(reinterpret_cast<const float4 *>(rec) + 3);
    h.z0 = row3.x; h.z1 = row3.y; h.z2 = row3.z;
}

/* the three edge functions of a triangle prepared for evaluation at (px, py):
 * e_k = (px - ax[k]) * dy[k] - (py - ay[k]) * dx[k], with all of them and inv_area negated for negative areas */
struct EdgeSet {
    float ax[3], ay[3], dx[3], dy[3];
    float inv_area;
};

__device__ __forceinline__ void prepare_edges(EdgeSet &E, int x0, int y0, int x1, int y1, int x2, int y2, float area, float inv_area)
{
    const float fx0 = (float)x0, fy0 = (float)y0, fx1 = (float)x1, fy1 = (float)y1, fx2 = (float)x2, fy2 = (float)y2;
    /* raster.c:536-538: w0 = edge(v1, v2, p), w1 = edge(v2, v0, p), w2 = edge(v0, v1, p) */
    E.ax[0] = fx1; E.ay[0] = fy1; E.dx[0] = fx2 - fx1; E.dy[0] = fy2 - fy1;
    E.ax[1] = fx2; E.ay[1] = fy2; E.dx[1] = fx0 - fx2; E.dy[1] = fy0 - fy2;
    E.ax[2] = fx0; E.ay[2] = fy0; E.dx[2] = fx1 - fx0; E.dy[2] = fy1 - fy0;
    E.inv_area = inv_area;
    if (!(area > 0)) {
#pragma unroll
        for (int k = 0; k < 3; k++) { E.dx[k] = -E.dx[k]; E.dy[k] = -E.dy[k]; }
        E.inv_area = -inv_area;
    }
}

__device__ __forceinline__ bool coord_small(int32_t v)      /* |v| < VIS_COORD_LIMIT without overflowing on INT32_MIN */
{
    return (uint32_t)v + (uint32_t)VIS_COORD_LIMIT < 2u * (uint32_t)VIS_COORD_LIMIT;
}


/* one triangle over its clamped box in 8x4 pixel blocks, the blocks first, first + stride, ... by the calling warp:
 * the reference's expressions evaluated directly (any size of triangle, any coordinates) */
__device__ __forceinline__ void raster_blocks(unsigned long long *keys, const VisMode &mode, const BatchDev &b, const VisHead &h,
                                              int px0, int py0, int first, int stride)
{
    const uint32_t lane = threadIdx.x & 31;
    const int X0 = max((int)(h.row2.x & 0xFFFFu) - px0, 0), Y0 = max((int)(h.row2.x >> 16) - py0, 0);
    const int X1 = min((int)(h.row2.y & 0xFFFFu) - px0, TILE_W - 1), Y1 = min((int)(h.row2.y >> 16) - py0, TILE_H - 1);
    bool range01 = true;
    double dnear = 0.0, dfar = 1.0;
    if (!mode.all_range01) {
        const RasterCfg *cfg = b.cfgs + (h.row2.z & STATE_INDEX_MASK);
        range01 = (cfg->flags & RC_DEPTH_RANGE_01) != 0u;
        dnear = cfg->depth_near; dfar = cfg->depth_far;
    }
    EdgeSet E;
    prepare_edges(E, h.row0.x, h.row0.y, h.row0.z, h.row0.w, h.row1.x, h.row1.y, __int_as_float(h.row1.z), __int_as_float(h.row1.w));
    const int nbx = (X1 - X0 + 8) >> 3, nby = (Y1 - Y0 + 4) >> 2;
    const int nblk = nbx * nby;
    for (int blk = first; blk < nblk; blk += stride) {
        const int byi = blk / nbx, bxi = blk - byi * nbx;
        const int x = X0 + bxi * 8 + (int)(lane & 7), y = Y0 + byi * 4 + (int)(lane >> 3);
        if (x > X1 || y > Y1) continue;
        const float fx = (float)(px0 + x), fy = (float)(py0 + y);
        const float e0 = (fx - E.ax[0]) * E.dy[0] - (fy - E.ay[0]) * E.dx[0];
        const float e1 = (fx - E.ax[1]) * E.dy[1] - (fy - E.ay[1]) * E.dx[1];
        const float e2 = (fx - E.ax[2]) * E.dy[2] - (fy - E.ay[2]) * E.dx[2];
        if (fminf(fminf(e0, e1), e2) >= 0.0f) {         /* inclusive on all three edges (raster.c:539-540) */
            const float b0 = e0 * E.inv_area, b1 = e1 * E.inv_area, b2 = e2 * E.inv_area;
            const float z = b0 * h.z0 + b1 * h.z1 + b2 * h.z2;
            key_min(&keys[y * VIS_PITCH + x], make_key(mode, depth_of(z, range01, dnear, dfar), h.row2.w));
        }
    }
}

/* THREADS = 256: four tiles per SM, the throughput shape for grids of several waves.  THREADS = 512: two tiles per SM
 * with sixteen warps on each tile's list -- for grids of at most about one wave (a band of a multi-GPU frame), where
 * the kernel's duration is the time of its heaviest tile. */
template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS)
k_vis(BatchDev b, FrameTargets fb, ClearOp clr, uint32_t planes, uint32_t depth_func, uint32_t all_range01)
{
    constexpr uint32_t NWARPS = THREADS / 32;
    extern __shared__ __align__(16) unsigned char vis_smem_raw[];
    VisSmem<THREADS> &sm = *reinterpret_cast<VisSmem<THREADS> *>(vis_smem_raw);
    if (!lists_fit(b)) return;

    const uint32_t tile = b.tile_order ? b.tile_order[blockIdx.x] : blockIdx.x;
    const int tx = (int)(tile % (uint32_t)fb.tiles_x), ty = (int)(tile / (uint32_t)fb.tiles_x) + fb.tile_y0;
    const int px0 = tx << TILE_LOG, py0t = ty << TILE_LOG;
    const int py0 = max(py0t, fb.band_y0);          /* shared-memory row 0 is framebuffer row py0 */
    const int vw = min(TILE_W, fb.width - px0);
    const int vh = min(py0t + TILE_H, fb.band_y1) - py0;
    if (vw <= 0 || vh <= 0) return;

    const uint32_t L = b.tile_count ? b.tile_count[tile] : 0u;
    if (L && b.tile_flags[tile] != 0u) return;      /* an ordered kernel owns this tile */
    const bool clr_here = clr.mask && clr.x0 < px0 + vw && clr.x1 > px0 && clr.y0 < py0 + vh && clr.y1 > py0;
    if (L == 0 && !clr_here) return;

    VisMode mode;
    mode.inv = (depth_func == 4u || depth_func == 6u) ? 0xFFFFFFFFu : 0u;
    mode.first_wins = (depth_func == 1u || depth_func == 4u);
    mode.all_range01 = all_range01 != 0u;
    const uint32_t none_lo = mode.first_wins ? 0u : 0xFFFFFFFFu;

    /* ---- fused glClear (gl_api.c:409-457): stencil goes straight to HBM, depth becomes the initial keys ---- */
    const int cx0 = max(clr.x0 - px0, 0), cy0 = max(clr.y0 - py0, 0);
    const int cx1 = min(clr.x1 - px0, vw), cy1 = min(clr.y1 - py0, vh);
    const bool clr_depth = clr_here && (clr.mask & G_DEPTH_BUFFER_BIT) && (planes & 2u);
    if (clr_here && (clr.mask & G_STENCIL_BUFFER_BIT) && (planes & 4u)) {
        for (int i = threadIdx.x; i < vh * TILE_W; i += THREADS) {
            const int y = i >> 6, x = i & 63;
            if (x >= cx0 && x < cx1 && y >= cy0 && y < cy1) fb.stencil[(size_t)(py0 + y) * fb.width + px0 + x] = (uint8_t)clr.stencil;
        }
    }
    if (L == 0) {
        if (clr_depth)
            for (int i = threadIdx.x; i < vh * TILE_W; i += THREADS) {
                const int y = i >> 6, x = i & 63;
                if (x >= cx0 && x < cx1 && y >= cy0 && y < cy1) fb.depth[(size_t)(py0 + y) * fb.width + px0 + x] = clr.depth;
            }
        return;
    }
    const bool clr_full = clr_depth && cx0 == 0 && cy0 == 0 && cx1 == vw && cy1 == vh;
    for (int i = threadIdx.x; i < vh * TILE_W; i += THREADS) {
        const int y = i >> 6, x = i & 63;
        float d = 0.0f;
        if (x < vw) {
            if (clr_full || (clr_depth && x >= cx0 && x < cx1 && y >= cy0 && y < cy1)) d = clr.depth;
            else d = fb.depth[(size_t)(py0 + y) * fb.width + px0 + x];
        }
        sm.key[y * VIS_PITCH + x] = ((unsigned long long)(ord_bits(d) ^ mode.inv) << 32) | none_lo;
    }
    if (threadIdx.x == 0) { sm.next_chunk = 0; sm.large_n = 0; }
    __syncthreads();

    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t lt_mask = (1u << lane) - 1u;
    const uint32_t *list = b.tile_list + b.tile_offset[tile];

    /* ---- phase 1: warps take up to 32 list entries at a time.  "Small" triangles have a clamped box of at most 256
     * pixels and a vertex extent below 2^11: then every edge value inside the tile is an integer below 2^24, exactly
     * represented in float, and e = A*x + B*y + C with tile-relative x, y gives the same bits as the reference's
     * expression (raster.c:299-302).  Because the values are exact, the covered pixels of a box row are exactly the
     * integer solutions of three linear inequalities: the warp first turns (triangle, row) pairs -- 32 at a time, one per
     * lane -- into spans [x_left, x_right], then lays the spans end to end into one run of pixels and walks it 32 pixels
     * per step, one COVERED pixel per lane (a box walk tests ~3 pixels per covered one on C4's slivers).  Everything
     * else is queued for phase 2. ---- */
    const float px0f = (float)px0, py0f = (float)py0;
    /* guided self-scheduling: the list is dealt in chunks of 32 entries, then 16, then 8 towards its end, so that the
     * warps of the tile finish phase 1 within a fraction of a chunk of each other (the barrier before phase 2 was 16 %
     * of the kernel's stall samples with uniform chunks).  Ticket k -> [c, c_end) is a pure function of L. */
    const uint32_t tail8 = min(L, 64u);
    const uint32_t head32 = ((L - tail8) > 128u) ? ((L - tail8 - 128u) & ~31u) : 0u;
    const uint32_t mid16 = L - tail8 - head32;
    const uint32_t n1 = head32 >> 5, n2 = (mid16 + 15u) >> 4;
    for (;;) {
        uint32_t k = 0;
        if (lane == 0) k = atomicAdd(&sm.next_chunk, 1u);
        k = __shfl_sync(0xFFFFFFFFu, k, 0);
        uint32_t c, c_end;
        if (k < n1) { c = k << 5; c_end = c + 32u; }
        else if (k - n1 < n2) { c = head32 + ((k - n1) << 4); c_end = min(c + 16u, head32 + mid16); }
        else { c = head32 + mid16 + ((k - n1 - n2) << 3); c_end = min(c + 8u, L); }
        if (c >= L) break;
        const uint32_t e = (c + lane < c_end) ? c + lane : L;
        uint32_t nrows = 0;                     /* box rows of this lane's triangle; 0 = no small triangle here */
        bool alone = false;                     /* a large triangle that did not fit the queue */
        VisHead h;
        int X0 = 0, Y0 = 0, bw = 1;
        if (e < L) {
            const uint32_t r = list[e];
            load_vis_head(h, b.records + r);
            /* the record's box is already clamped to viewport, scissor, framebuffer and band: clamp to the tile */
            X0 = max((int)(h.row2.x & 0xFFFFu) - px0, 0); Y0 = max((int)(h.row2.x >> 16) - py0, 0);
            const int X1 = min((int)(h.row2.y & 0xFFFFu) - px0, TILE_W - 1), Y1 = min((int)(h.row2.y >> 16) - py0, TILE_H - 1);
            bw = X1 - X0 + 1;
            nrows = (uint32_t)(Y1 - Y0 + 1);
            const bool small = (uint32_t)bw * nrows <= (uint32_t)(L >= VIS_BUSY_LIST ? VIS_SMALL_AREA_BUSY : VIS_SMALL_AREA) && mode.all_range01 &&
                               coord_small(h.row0.x) && coord_small(h.row0.y) && coord_small(h.row0.z) && coord_small(h.row0.w) &&
                               coord_small(h.row1.x) && coord_small(h.row1.y) &&
                               max(max(h.row0.x, h.row0.z), h.row1.x) - min(min(h.row0.x, h.row0.z), h.row1.x) <= VIS_EXACT_EXTENT &&
                               max(max(h.row0.y, h.row0.w), h.row1.y) - min(min(h.row0.y, h.row0.w), h.row1.y) <= VIS_EXACT_EXTENT;
            if (!small) {
                const uint32_t at = atomicAdd(&sm.large_n, 1u);
                if (at < (uint32_t)VIS_LARGE_CAP) sm.large_rec[at] = r; else alone = true;
                nrows = 0;
            }
        }
        /* (rare) queue full: the warp rasterises those triangles by itself, one after the other */
        for (uint32_t am = __ballot_sync(0xFFFFFFFFu, alone); am; am &= am - 1) {
            const int src = __ffs(am) - 1;
            VisHead g;
            g.row0.x = __shfl_sync(0xFFFFFFFFu, h.row0.x, src); g.row0.y = __shfl_sync(0xFFFFFFFFu, h.row0.y, src);
            g.row0.z = __shfl_sync(0xFFFFFFFFu, h.row0.z, src); g.row0.w = __shfl_sync(0xFFFFFFFFu, h.row0.w, src);
            g.row1.x = __shfl_sync(0xFFFFFFFFu, h.row1.x, src); g.row1.y = __shfl_sync(0xFFFFFFFFu, h.row1.y, src);
            g.row1.z = __shfl_sync(0xFFFFFFFFu, h.row1.z, src); g.row1.w = __shfl_sync(0xFFFFFFFFu, h.row1.w, src);
            g.row2.x = __shfl_sync(0xFFFFFFFFu, h.row2.x, src); g.row2.y = __shfl_sync(0xFFFFFFFFu, h.row2.y, src);
            g.row2.z = __shfl_sync(0xFFFFFFFFu, h.row2.z, src); g.row2.w = __shfl_sync(0xFFFFFFFFu, h.row2.w, src);
            g.z0 = __shfl_sync(0xFFFFFFFFu, h.z0, src); g.z1 = __shfl_sync(0xFFFFFFFFu, h.z1, src); g.z2 = __shfl_sync(0xFFFFFFFFu, h.z2, src);
            raster_blocks(sm.key, mode, b, g, px0, py0, 0, 1);
            __syncwarp();
        }
        const uint32_t smask = __ballot_sync(0xFFFFFFFFu, nrows != 0u);
        if (!smask) continue;
        /* first row of each small triangle in the run of rows (exclusive scan of the row counts) */
        uint32_t incl = nrows;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= (uint32_t)o) incl += up;
        }
        const uint32_t total_rows = __shfl_sync(0xFFFFFFFFu, incl, 31);
        const uint32_t row_start = incl - nrows;
        VisWarp &vw_ = sm.w[warp];
        if (nrows) {                            /* prepared triangle -> this warp's staging slots, in compacted order */
            EdgeSet E;
            prepare_edges(E, h.row0.x, h.row0.y, h.row0.z, h.row0.w, h.row1.x, h.row1.y, __int_as_float(h.row1.z), __int_as_float(h.row1.w));
            float A[3], B[3], C[3], RN[3];
#pragma unroll
            for (int k = 0; k < 3; k++) {       /* all operands and results are integers below 2^24: exact */
                A[k] = E.dy[k]; B[k] = -E.dx[k];
                C[k] = (px0f - E.ax[k]) * E.dy[k] - (py0f - E.ay[k]) * E.dx[k];
                RN[k] = (A[k] != 0.0f) ? -1.0f / A[k] : 0.0f;
            }
            const uint32_t ci = (uint32_t)__popc(smask & lt_mask);
            vw_.T0[ci] = make_float4(A[0], A[1], A[2], E.inv_area);
            vw_.T1[ci] = make_float4(h.z0, h.z1, h.z2, __uint_as_float(h.row2.w));
            vw_.T2[ci] = make_float4(B[0], B[1], B[2],
                                     __uint_as_float((uint32_t)X0 | ((uint32_t)Y0 << 6) | ((uint32_t)(bw - 1) << 12) | (row_start << 18)));
#pragma unroll
            for (int k = 0; k < 3; k++) { vw_.C[k][ci] = C[k]; vw_.RN[k][ci] = RN[k]; }
        }
        __syncwarp();
        uint32_t before_t = 0;                  /* small triangles that start before the current group of rows */
        for (uint32_t rbase = 0; rbase < total_rows; rbase += 32) {
            /* ---- span stage: lane = one (triangle, row) pair ---- */
            const int rel = (int)row_start - (int)rbase;
            const uint32_t tmarks = __reduce_or_sync(0xFFFFFFFFu, (nrows != 0u && rel >= 0 && rel < 32) ? (1u << rel) : 0u);
            const uint32_t rho = rbase + lane;
            uint32_t width = 0, sbits = 0;
            float t0 = 0.0f, t1 = 0.0f, t2 = 0.0f;
            if (rho < total_rows) {
                const uint32_t ci = before_t + (uint32_t)__popc(tmarks & (0xFFFFFFFFu >> (31u - lane))) - 1u;   /* the triangle whose box holds row rho */
                const float4 g2 = vw_.T2[ci];
                const uint32_t bs = __float_as_uint(g2.w);
                const int y = (int)((bs >> 6) & 63u) + (int)(rho - (bs >> 18));
                const float fy = (float)y;
                t0 = __fmaf_rn(g2.x, fy, vw_.C[0][ci]);
                t1 = __fmaf_rn(g2.y, fy, vw_.C[1][ci]);
                t2 = __fmaf_rn(g2.z, fy, vw_.C[2][ci]);
                /* A*x + t >= 0  <=>  x >= -t/A (A > 0)  or  x <= -t/A (A < 0)  or  t >= 0 (A == 0).  q = t * (-1/A) is within
                 * 2^-13 of the quotient wherever it matters (|q| < 2^10); a non-integral quotient of integers with |A| < 2^11 is
                 * at least 2^-11 away from an integer, so rounding q -+ 2^-12 up / down is exact.  The pixel stage re-tests
                 * every pixel anyway. */
                float lo = (float)(int)(bs & 63u), hi = lo + (float)(int)((bs >> 12) & 63u);
                bool ok = true;
                const float tk[3] = { t0, t1, t2 };
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    const float rn = vw_.RN[k][ci];
                    const float q = fminf(fmaxf(tk[k] * rn, -128.0f), 128.0f);
                    if (rn < 0.0f) lo = fmaxf(lo, ceilf(q - 0.000244140625f));
                    else if (rn > 0.0f) hi = fminf(hi, floorf(q + 0.000244140625f));
                    else ok = ok && (tk[k] >= 0.0f);
                }
                if (ok && hi >= lo) {
                    width = (uint32_t)(int)(hi - lo) + 1u;
                    sbits = (uint32_t)(int)lo | ((uint32_t)y << 6) | (ci << 12);
                }
            }
            before_t += (uint32_t)__popc(tmarks);
            /* first pixel of each non-empty span in the run of pixels */
            uint32_t pincl = width;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t up = __shfl_up_sync(0xFFFFFFFFu, pincl, o);
                if (lane >= (uint32_t)o) pincl += up;
            }
            const uint32_t total_px = __shfl_sync(0xFFFFFFFFu, pincl, 31);
            const uint32_t pstart = pincl - width;
            const uint32_t nonempty = __ballot_sync(0xFFFFFFFFu, width != 0u);
            if (width) vw_.S[__popc(nonempty & lt_mask)] = make_float4(t0, t1, t2, __uint_as_float(sbits | (pstart << 17)));
            __syncwarp();
            /* ---- pixel stage: lane = one pixel of the run ---- */
            uint32_t before_s = 0;
            for (uint32_t pbase = 0; pbase < total_px; pbase += 32) {
                const int prel = (int)pstart - (int)pbase;
                const uint32_t smarks = __reduce_or_sync(0xFFFFFFFFu, (width != 0u && prel >= 0 && prel < 32) ? (1u << prel) : 0u);
                const uint32_t p = pbase + lane;
                if (p < total_px) {
                    const uint32_t si = before_s + (uint32_t)__popc(smarks & (0xFFFFFFFFu >> (31u - lane))) - 1u;   /* the span that holds pixel p */
                    const float4 sp = vw_.S[si];
                    const uint32_t sb = __float_as_uint(sp.w);
                    const uint32_t ci = (sb >> 12) & 31u;
                    const int x = (int)(sb & 63u) + (int)(p - (sb >> 17)), y = (int)((sb >> 6) & 63u);
                    const float fx = (float)x;
                    const float4 g0 = vw_.T0[ci];
                    const float e0 = __fmaf_rn(g0.x, fx, sp.x);
                    const float e1 = __fmaf_rn(g0.y, fx, sp.y);
                    const float e2 = __fmaf_rn(g0.z, fx, sp.z);
                    if (fminf(fminf(e0, e1), e2) >= 0.0f) {         /* inclusive on all three edges (raster.c:539-540) */
                        const float4 g1 = vw_.T1[ci];
                        const float b0 = e0 * g0.w, b1 = e1 * g0.w, b2 = e2 * g0.w;
                        const float z = b0 * g1.x + b1 * g1.y + b2 * g1.z;
                        key_min(&sm.key[y * VIS_PITCH + x], make_key(mode, depth_of(z, true, 0.0, 1.0), __float_as_uint(g1.w)));
                    }
                }
                before_s += (uint32_t)__popc(smarks);
            }
            __syncwarp();       /* the span slots are rewritten by the next group */
        }
        __syncwarp();           /* the triangle slots are rewritten by the next chunk */
    }
    __syncthreads();

    /* ---- phase 2: the queued triangles, all warps on each, 8x4 pixel blocks dealt round-robin (the first warp rotates
     * with the queue position so that one-block triangles do not all land on warp 0) ---- */
    const uint32_t nl = min(sm.large_n, (uint32_t)VIS_LARGE_CAP);
    for (uint32_t q = 0; q < nl; q++) {
        VisHead h;
        load_vis_head(h, b.records + sm.large_rec[q]);
        raster_blocks(sm.key, mode, b, h, px0, py0, (int)((warp + NWARPS - (q & (NWARPS - 1u))) & (NWARPS - 1u)), (int)NWARPS);
    }
    __syncthreads();

    /* ---- write-back: depth plane + visibility plane (record index of the surviving fragment) ---- */
    const uint32_t slot_mask = (1u << GROUP_SHIFT) - 1u;
    const bool vec = (vw == TILE_W) && ((fb.width & 3) == 0);
    if (vec) {
        for (int i = threadIdx.x; i < vh * 16; i += THREADS) {
            const int y = i >> 4, q4 = (i & 15) * 4;
            float dv[4];
            uint32_t rv[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const unsigned long long key = sm.key[y * VIS_PITCH + q4 + k];
                const uint32_t lo = (uint32_t)key;
                dv[k] = ord_float((uint32_t)(key >> 32) ^ mode.inv);
                if (lo == none_lo) rv[k] = VIS_NONE;
                else {
                    const uint32_t id = mode.first_wins ? lo - 1u : 0xFFFFFFFEu - lo;
                    rv[k] = __ldg(&b.group_base[id >> GROUP_SHIFT]) + (id & slot_mask);
                }
            }
            const size_t p = (size_t)(py0 + y) * fb.width + px0 + q4;
            if (planes & 2u) *reinterpret_cast<float4 *>(fb.depth + p) = make_float4(dv[0], dv[1], dv[2], dv[3]);
            *reinterpret_cast<uint4 *>(b.vis_plane + p) = make_uint4(rv[0], rv[1], rv[2], rv[3]);
        }
    } else {
        for (int i = threadIdx.x; i < vh * TILE_W; i += THREADS) {
            const int y = i >> 6, x = i & 63;
            if (x >= vw) continue;
            const unsigned long long key = sm.key[y * VIS_PITCH + x];
            const uint32_t lo = (uint32_t)key;
            uint32_t rv = VIS_NONE;
            if (lo != none_lo) {
                const uint32_t id = mode.first_wins ? lo - 1u : 0xFFFFFFFEu - lo;
                rv = __ldg(&b.group_base[id >> GROUP_SHIFT]) + (id & slot_mask);
            }
            const size_t p = (size_t)(py0 + y) * fb.width + px0 + x;
            if (planes & 2u) fb.depth[p] = ord_float((uint32_t)(key >> 32) ^ mode.inv);
            b.vis_plane[p] = rv;
        }
    }
}

void launch_vis_unordered(const BatchDev &b, const FrameTargets &fb, const ClearOp &clear, uint32_t planes, uint32_t depth_func,
                          bool all_range01, cudaStream_t s)
{
    const uint32_t tiles = (uint32_t)(fb.tiles_x * fb.tile_rows);
    static bool configured[64] = { false };
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !configured[dev]) {      /* more than 48 KB of shared memory is a per-device opt-in */
        cudaFuncSetAttribute(k_vis<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(VisSmem<256>));
        cudaFuncSetAttribute(k_vis<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(VisSmem<512>));
        configured[dev] = true;
    }
    if (small_grid(tiles)) k_vis<512><<<tiles, 512, sizeof(VisSmem<512>), s>>>(b, fb, clear, planes, depth_func, all_range01 ? 1u : 0u);
    else k_vis<256><<<tiles, 256, sizeof(VisSmem<256>), s>>>(b, fb, clear, planes, depth_func, all_range01 ? 1u : 0u);
    note_launch();
}

} // namespace mtgl_dev_impl
