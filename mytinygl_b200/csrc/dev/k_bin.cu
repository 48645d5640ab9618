/*
 * k_bin.cu -- K3: sort-middle binning of set-up sub-triangles into 64x64 screen tiles.
 *
 * No reference counterpart: the reference walks every triangle's bounding box in submission
 * order (src/raster.c:532-533).  Here each record is referenced from every tile its clamped
 * bounding box touches; the per-tile reference lists are built with count -> scan -> fill and
 * are NOT ordered by the binner -- every reference carries the record's submission-ordered id
 * and the tile kernel sorts its list, which is what preserves blend / stencil / depth-tie
 * semantics (hard part 2) without serialising this pass.
 *
 * Records whose box spans more than LARGE_TILES tiles (full-screen quads) are deferred to a
 * cooperative pass: one CTA per record, threads striding over its tiles, with a conservative
 * corner test of the three edge functions that drops tiles the triangle cannot touch.
 *
 * Algorithmic bytes: 8 B (packed bbox) read per record per pass + 4 B written per (record, tile).
 */
#include "dev_common.cuh"

#include <cstdlib>

namespace mtgl_dev_impl {

void note_launch();

__device__ __forceinline__ float edge_at(float ax, float ay, float bx, float by, float px, float py)   /* raster.c:299-302 */
{
    return (px - ax) * (by - ay) - (py - ay) * (bx - ax);
}

/* Can the triangle cover any pixel of the rectangle [x0,x1]x[y0,y1]?  The rounded float edge
 * function is monotone in px for fixed py and vice versa, so its extrema over the rectangle are
 * attained at corners: if one edge is strictly outside at all four corners no pixel passes the
 * inclusive test of raster.c:539-540. */
__device__ bool tile_may_overlap(const TriRecord *r, int x0, int y0, int x1, int y1)
{
    const float fx0 = (float)r->x0, fy0 = (float)r->y0, fx1 = (float)r->x1, fy1 = (float)r->y1, fx2 = (float)r->x2, fy2 = (float)r->y2;
    const float ax[3] = { fx1, fx2, fx0 }, ay[3] = { fy1, fy2, fy0 }, bx[3] = { fx2, fx0, fx1 }, by[3] = { fy2, fy0, fy1 };
    const bool pos = r->area > 0;
    const float cx[2] = { (float)x0, (float)x1 }, cy[2] = { (float)y0, (float)y1 };
#pragma unroll
    for (int e = 0; e < 3; e++) {
        bool all_out = true;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            float v = edge_at(ax[e], ay[e], bx[e], by[e], cx[k & 1], cy[k >> 1]);
            bool out = pos ? (v < 0) : (v > 0);
            all_out = all_out && out;
        }
        if (all_out) return false;
    }
    return true;
}

struct TileRange { int tx0, ty0, tx1, ty1; };

__device__ __forceinline__ TileRange tile_range(const TriRecord *r, const FrameTargets &fb)
{
    TileRange t;
    t.tx0 = (int)(r->bbox_min & 0xFFFFu) >> TILE_LOG; t.tx1 = (int)(r->bbox_max & 0xFFFFu) >> TILE_LOG;
    t.ty0 = ((int)(r->bbox_min >> 16) >> TILE_LOG) - fb.tile_y0; t.ty1 = ((int)(r->bbox_max >> 16) >> TILE_LOG) - fb.tile_y0;
    return t;
}

/* Pass 1 for the records that touch at most LARGE_TILES tiles: write their references through the per-tile cursors.
 * (Pass 0 -- counting -- is fused into k_setup.)
 *
 * What bounds this pass is the serialisation of atomics on the same cursor: a mesh puts hundreds of consecutive records
 * into the same tile.  So single-tile records are aggregated twice before they reach L2 -- per warp (match.any), then per
 * CTA in a small shared-memory table -- and a CTA issues ONE cursor atomic per distinct tile of its 256 records. */
constexpr uint32_t FILL_SLOTS = 64;             /* shared-memory table: distinct tiles per CTA iteration (more: straight to L2) */

__device__ __forceinline__ void bin_small_fill(const BatchDev &b, const FrameTargets &fb, uint32_t block, uint32_t nblocks)
{
    __shared__ uint32_t s_tile[FILL_SLOTS], s_count[FILL_SLOTS], s_base[FILL_SLOTS];
    const uint32_t n = b.counters->records;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t lt_mask = (1u << lane) - 1u;
    for (uint32_t base_r = block * blockDim.x; base_r < n; base_r += nblocks * blockDim.x) {       /* uniform per CTA */
        if (threadIdx.x < FILL_SLOTS) { s_tile[threadIdx.x] = 0xFFFFFFFFu; s_count[threadIdx.x] = 0u; }
        __syncthreads();
        const uint32_t r = base_r + threadIdx.x;
        int tx0 = 0, tx1 = -1, ty0 = 0, ty1 = -1, ntiles = 0;
        if (r < n) {
            const uint4 box = b.bin_rows[r];        /* bbox_min, bbox_max, state_flags, id */
            tx0 = (int)(box.x & 0xFFFFu) >> TILE_LOG; tx1 = (int)(box.y & 0xFFFFu) >> TILE_LOG;
            ty0 = ((int)(box.x >> 16) >> TILE_LOG) - fb.tile_y0; ty1 = ((int)(box.y >> 16) >> TILE_LOG) - fb.tile_y0;
            ntiles = (tx1 - tx0 + 1) * (ty1 - ty0 + 1);
            if (ntiles > LARGE_TILES) ntiles = 0;       /* the cooperative binner's */
        }
        /* ---- every (record, tile) reference takes a rank in its tile's table slot; single-tile records, the bulk, as
         * warp groups (match.any), the two to four references of a record across a tile edge or corner one by one ---- */
        auto table_slot = [&](uint32_t tile) -> uint32_t {
            uint32_t h = (tile * 2654435761u) >> 26;
            for (uint32_t probe = 0; probe < FILL_SLOTS; probe++, h = (h + 1u) & (FILL_SLOTS - 1u)) {
                const uint32_t prev = atomicCAS(&s_tile[h], 0xFFFFFFFFu, tile);
                if (prev == 0xFFFFFFFFu || prev == tile) return h;
            }
            return FILL_SLOTS;      /* table full: the reference goes straight to L2 */
        };
        const uint32_t single = __ballot_sync(0xFFFFFFFFu, ntiles == 1);
        uint32_t tl[4] = { 0u, 0u, 0u, 0u }, slot[4] = { FILL_SLOTS, FILL_SLOTS, FILL_SLOTS, FILL_SLOTS }, rank[4] = { 0u, 0u, 0u, 0u }, off[4] = { 0u, 0u, 0u, 0u };
        const int nref = (ntiles >= 1 && ntiles <= 4) ? ntiles : 0;
        if (ntiles == 1) {
            tl[0] = (uint32_t)(ty0 * fb.tiles_x + tx0);
            off[0] = __ldg(&b.tile_offset[tl[0]]);
            const uint32_t peers = __match_any_sync(single, tl[0]);
            const int leader = __ffs(peers) - 1;
            uint32_t sl = FILL_SLOTS, rk = 0;
            if ((int)lane == leader) {
                sl = table_slot(tl[0]);
                rk = (sl < FILL_SLOTS) ? atomicAdd(&s_count[sl], (uint32_t)__popc(peers)) : atomicAdd(&b.tile_cursor[tl[0]], (uint32_t)__popc(peers));
            }
            slot[0] = __shfl_sync(peers, sl, leader);
            rank[0] = __shfl_sync(peers, rk, leader) + (uint32_t)__popc(peers & lt_mask);
        } else if (nref > 1) {
            const int w = tx1 - tx0 + 1;            /* 1 or 2 columns when the record spans more than one row; up to 4 in one row */
#pragma unroll
            for (int k = 0; k < 4; k++) {           /* (static indices: the arrays stay in registers) */
                const int row = (w == 1) ? k : ((w == 2) ? (k >> 1) : 0), col = (w == 1) ? 0 : ((w == 2) ? (k & 1) : k);
                if (k < nref) {
                    tl[k] = (uint32_t)((ty0 + row) * fb.tiles_x + tx0 + col);
                    off[k] = __ldg(&b.tile_offset[tl[k]]);
                    slot[k] = table_slot(tl[k]);
                    rank[k] = (slot[k] < FILL_SLOTS) ? atomicAdd(&s_count[slot[k]], 1u) : atomicAdd(&b.tile_cursor[tl[k]], 1u);
                }
            }
        }
        __syncthreads();
        if (threadIdx.x < FILL_SLOTS && s_count[threadIdx.x]) s_base[threadIdx.x] = atomicAdd(&b.tile_cursor[s_tile[threadIdx.x]], s_count[threadIdx.x]);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (k < nref) b.tile_list[off[k] + (slot[k] < FILL_SLOTS ? s_base[slot[k]] : 0u) + rank[k]] = r;
        if (ntiles > 4) {
            for (int ty = ty0; ty <= ty1; ty++)
                for (int tx = tx0; tx <= tx1; tx++) {
                    const uint32_t t2 = (uint32_t)(ty * fb.tiles_x + tx);
                    const uint32_t o2 = __ldg(&b.tile_offset[t2]);
                    const uint32_t a2 = atomicAdd(&b.tile_cursor[t2], 1u);
                    b.tile_list[o2 + a2] = r;
                }
        }
        __syncthreads();        /* the table is reset by the next iteration */
    }
}

template <int PASS>
__device__ __forceinline__ void bin_large(const BatchDev &b, const FrameTargets &fb, uint32_t block, uint32_t nblocks)
{
    const uint32_t n = b.counters->large_count;
    for (uint32_t i = block; i < n; i += nblocks) {
        const uint32_t r = b.large_list[i];
        const TriRecord *rec = b.records + r;
        TileRange tr = tile_range(rec, fb);
        int w = tr.tx1 - tr.tx0 + 1, cnt = w * (tr.ty1 - tr.ty0 + 1);
        for (int k = threadIdx.x; k < cnt; k += blockDim.x) {
            int tx = tr.tx0 + k % w, ty = tr.ty0 + k / w;
            int bx0 = (int)(rec->bbox_min & 0xFFFFu), by0 = (int)(rec->bbox_min >> 16);
            int bx1 = (int)(rec->bbox_max & 0xFFFFu), by1 = (int)(rec->bbox_max >> 16);
            int px0 = max(tx << TILE_LOG, bx0), px1 = min((tx << TILE_LOG) + TILE_W - 1, bx1);
            int py0 = max((ty + fb.tile_y0) << TILE_LOG, by0), py1 = min(((ty + fb.tile_y0) << TILE_LOG) + TILE_H - 1, by1);
            /* the exact-coverage corner test only knows triangles; lines and points are binned by their box */
            if ((rec->state_flags & STATE_KIND_MASK) == (KIND_TRIANGLE << STATE_KIND_SHIFT) && !tile_may_overlap(rec, px0, py0, px1, py1)) continue;
            uint32_t tile = (uint32_t)(ty * fb.tiles_x + tx);
            if (PASS == 0) {
                atomicAdd(&b.tile_count[tile], 1u);
                const uint32_t tflags = tile_flag_bits(rec->state_flags);
                if (tflags) atomicOr(&b.tile_flags[tile], tflags);
            } else {
                uint32_t at = atomicAdd(&b.tile_cursor[tile], 1u);
                b.tile_list[b.tile_offset[tile] + at] = r;
            }
        }
    }
}

constexpr uint32_t BIN_SMALL_BLOCKS_MIN = 148 * 8, BIN_LARGE_BLOCKS = 148 * 4;


/* pass 1, one launch: the first BIN_SMALL_BLOCKS CTAs fill in the records that touch at most LARGE_TILES tiles, the
 * rest go through the list of large records together */
__global__ void __launch_bounds__(256) k_bin_fill(BatchDev b, FrameTargets fb, uint32_t small_blocks)
{
    if (!lists_fit(b)) return;
    if (blockIdx.x < small_blocks) bin_small_fill(b, fb, blockIdx.x, small_blocks);
    else bin_large<1>(b, fb, blockIdx.x - small_blocks, BIN_LARGE_BLOCKS);
}

/* One launch for the two steps between set-up and the fill pass:
 *   count  references of the LARGE records per tile (the cooperative binner's pass 0; the count pass of the small records
 *          is fused into k_setup), all CTAs;
 *   scan   exclusive scan of the per-tile counts + the launch order of the tile kernels, by the CTA that finishes its
 *          share of the counting last (ticket in DevCounters: no second launch, no idle GPU in between).
 * At most 256x256 tiles for a 16384^2 framebuffer. */
__global__ void __launch_bounds__(1024) k_bin_scan(BatchDev b, FrameTargets fb, uint32_t ntiles)
{
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry;
    __shared__ uint32_t ticket;
    bin_large<0>(b, fb, blockIdx.x, gridDim.x);
    __threadfence();                    /* this CTA's counts are visible before its ticket is */
    __syncthreads();
    if (threadIdx.x == 0) { ticket = atomicAdd(&b.counters->scan_ticket, 1u); carry = 0; }
    __syncthreads();
    if (ticket != gridDim.x - 1u) return;
    __threadfence();
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t base = 0; base < ntiles; base += blockDim.x) {
        uint32_t i = base + threadIdx.x;
        uint32_t v = (i < ntiles) ? __ldcg(&b.tile_count[i]) : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t n = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= (uint32_t)o) incl += n;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = warp_sums[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t n = __shfl_up_sync(0xFFFFFFFFu, wi, o);
                if (lane >= (uint32_t)o) wi += n;
            }
            warp_sums[lane] = wi - w;
        }
        __syncthreads();
        uint32_t excl = carry + warp_sums[warp] + incl - v;
        if (i < ntiles) { b.tile_offset[i] = excl; b.tile_cursor[i] = 0u; }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        b.counters->tile_refs = carry;
        /* the host sizes the reference lists from these: written straight into its pinned, device-mapped copy (a
         * device-to-host memcpy here costs a compute -> copy-engine -> compute round trip of ~17 us in mid-frame) */
        if (b.host_counters) {
            DevCounters c;
            c.records = __ldcg(&b.counters->records); c.large_count = __ldcg(&b.counters->large_count);
            c.triangles_in = __ldcg(&b.counters->triangles_in); c.overflow = __ldcg(&b.counters->overflow);
            c.culled_chunks = __ldcg(&b.counters->culled_chunks); c.scan_ticket = 0u; c.pad_[0] = 0u;
            c.tile_refs = carry;
            *b.host_counters = c;
            __threadfence_system();
        }
    }

    /* Launch order of the tile kernels: tiles by decreasing reference count (longest-processing-time-first), a
     * STABLE counting sort over 64 linear buckets -- equal loads keep their row-major order, so a uniform frame
     * launches exactly as before.  The hardware hands CTAs to SMs in blockIdx order, so the heavy tiles start first
     * and the light ones fill the tail; with plain row-major order the SMs' busy time differed by 14 % on C4.
     * Every warp owns a contiguous range of tiles; counts per (bucket, warp), scanned bucket-major, give each warp
     * its first slot in every bucket. */
    if (!b.tile_order) return;
    __shared__ uint32_t hist[64 * 32];
    __shared__ uint32_t maxc;
    for (uint32_t i = threadIdx.x; i < 64 * 32; i += blockDim.x) hist[i] = 0;
    if (threadIdx.x == 0) maxc = 0;
    __syncthreads();
    uint32_t m = 0;
    for (uint32_t i = threadIdx.x; i < ntiles; i += blockDim.x) m = max(m, __ldcg(&b.tile_count[i]));
    m = __reduce_max_sync(0xFFFFFFFFu, m);
    if (lane == 0 && m) atomicMax(&maxc, m);
    __syncthreads();
    const unsigned long long span = (unsigned long long)maxc + 1ull;
    const uint32_t per_warp = ((ntiles + 1023u) >> 10) << 5;           /* tiles per warp, a multiple of 32 */
    const uint32_t w_begin = min(warp * per_warp, ntiles), w_end = min(w_begin + per_warp, ntiles);
    for (uint32_t i0 = w_begin; i0 < w_end; i0 += 32) {
        const uint32_t i = i0 + lane;
        if (i < w_end) atomicAdd(&hist[(63u - (uint32_t)(((unsigned long long)__ldcg(&b.tile_count[i]) * 64ull) / span)) * 32u + warp], 1u);
    }
    __syncthreads();
    {   /* exclusive scan of the 2048 counters, two per thread */
        const uint32_t a0 = hist[2 * threadIdx.x], a1 = hist[2 * threadIdx.x + 1];
        uint32_t incl = a0 + a1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t n = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= (uint32_t)o) incl += n;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = warp_sums[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t n = __shfl_up_sync(0xFFFFFFFFu, wi, o);
                if (lane >= (uint32_t)o) wi += n;
            }
            warp_sums[lane] = wi - w;
        }
        __syncthreads();
        const uint32_t excl = warp_sums[warp] + incl - a0 - a1;
        hist[2 * threadIdx.x] = excl;
        hist[2 * threadIdx.x + 1] = excl + a0;
    }
    __syncthreads();
    const uint32_t lt_mask = (1u << lane) - 1u;
    for (uint32_t i0 = w_begin; i0 < w_end; i0 += 32) {
        const uint32_t i = i0 + lane;
        const bool on = i < w_end;
        const uint32_t act = __ballot_sync(0xFFFFFFFFu, on);
        if (on) {
            const uint32_t bucket = 63u - (uint32_t)(((unsigned long long)__ldcg(&b.tile_count[i]) * 64ull) / span);
            const uint32_t peers = __match_any_sync(act, bucket);
            const int leader = __ffs(peers) - 1;
            uint32_t base = 0;
            if ((int)lane == leader) { base = hist[bucket * 32u + warp]; hist[bucket * 32u + warp] = base + (uint32_t)__popc(peers); }
            base = __shfl_sync(peers, base, leader);
            b.tile_order[base + (uint32_t)__popc(peers & lt_mask)] = i;
        }
        __syncwarp();
    }
}

void launch_bin_scan(const BatchDev &b, const FrameTargets &fb, cudaStream_t s)
{
    k_bin_scan<<<148, 1024, 0, s>>>(b, fb, (uint32_t)(fb.tiles_x * fb.tile_rows));      /* one CTA per SM counts; the last one scans */
    note_launch();
}

void launch_bin_fill(const BatchDev &b, const FrameTargets &fb, cudaStream_t s)
{
    /* (more CTAs than this -- one record per thread -- measured slower: 0.047 against 0.034 ms on C4, the cursor atomics
     * of a tile then all arrive at once) */
    const uint32_t small_blocks = BIN_SMALL_BLOCKS_MIN;
    k_bin_fill<<<small_blocks + BIN_LARGE_BLOCKS, 256, 0, s>>>(b, fb, small_blocks);
    note_launch();
}

} // namespace mtgl_dev_impl
