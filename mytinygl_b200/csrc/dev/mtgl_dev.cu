/*
 * mtgl_dev.cu -- host side of the C ABI (include/mtgl_dev.h): device memory management, object
 * mirrors (buffer objects, textures + mip level 1), batch upload and kernel orchestration.
 *
 * HBM layout per context:
 *   framebuffer planes   colour u32 / depth f32 / stencil u8, row 0 = top, pitch = width -- the
 *                        reference's framebuffer_t (src/framebuffer.h:19-25), so read-back is a memcpy
 *   object mirrors       one allocation per buffer object; per texture level 0 and level 1 (RGBA8 words)
 *   batch arena          states | raster cfgs | staged vertices | draw records | prefix tables | index blob
 *                        (one pinned-host -> device copy per batch)
 *   vertex streams       3 (5 with per-fragment lighting) float4 SoA arrays, one slot per vertex
 *   triangle records     160 B per surviving sub-triangle, worst case 7 per input triangle
 *   tile tables          count / offset / cursor per 64x64 tile, and the reference lists
 * All of it is grow-only and reused across batches; nothing is allocated per frame in steady state.
 *
 * There is no CPU path in this file: every entry point either runs CUDA work or fails.
 */
#include "dev_common.cuh"
#include "dev_fill.cuh"
#include "dev_vertex.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <new>
#include <vector>

using namespace mtgl_dev_impl;

namespace {

constexpr uint32_t kMaxObjects = 257;               /* textures */
constexpr uint32_t kMaxBuffers = MTGL_MAX_BUFFER_IDS; /* glGenBuffers names + the front end's display-list buffers */
constexpr uint32_t kMaxTrisPerPass = 4u << 20;
constexpr size_t kKernelUploadMax = 1u << 20;   /* batch arenas up to this size are uploaded by k_upload (k_misc.cu) */

struct DevBuf {
    void *ptr = nullptr;
    size_t cap = 0;
};

struct TexObj { uint32_t *l0 = nullptr, *l1 = nullptr; float4 *f4 = nullptr; int w = 0, h = 0, w1 = 0, h1 = 0; };
constexpr int kFloatTexelsMax = 256 * 256;      /* textures up to this size also keep a float4 copy (16 B / texel + level 1) */
struct BufObj {
    uint8_t *ptr = nullptr; uint64_t size = 0;
    uint64_t gen = 0;               /* bumped by every write through the API */
    uint32_t draws_since_write = 0; /* array draws that sourced positions from the unchanged content */
    bool exposed = false;           /* mtgl_dev_buffer_pointer handed the storage out: content may change behind the API */
};
/* cached object-space boxes of one (buffer content, position layout, element range): k_cull.cu */
struct BoundsEntry {
    uint32_t buffer = 0; uint64_t gen = 0, offset = 0; uint32_t stride = 0, size = 0; int32_t first = 0; uint32_t nverts = 0;
    DevBuf boxes; uint64_t last_use = 0;
    uint64_t batch = 0;             /* serial of the last batch that refers to these boxes: not evictable while it is being built */
    /* whole-frame rendering of geometry that is entirely on screen: the pass drops nothing, so after a few such batches
     * it is left out for a while and probed again later (a band-limited device always runs it) */
    uint32_t zero_streak = 0, skip_left = 0;
};
constexpr uint32_t kCullZeroStreak = 4, kCullSkipBatches = 60;
constexpr size_t kBoundsEntries = 8;
constexpr uint32_t kBoundsMinTriangles = 2 * SETUP_THREADS;    /* smaller draws are not worth a culling pass (and do not own their chunks) */

} // namespace

struct mtgl_dev {
    int device = 0;
    cudaStream_t stream = nullptr;
    int32_t width = 0, height = 0, band_y0 = 0, band_y1 = 0;
    uint32_t *color = nullptr;
    float *depth = nullptr;
    uint8_t *stencil = nullptr;
    TexObj tex[kMaxObjects];
    BufObj buf[kMaxBuffers];
    float *unorm8 = nullptr;

    DevBuf arena, v_clip, v_color, v_tex, v_epos, v_enrm, records, rec_eye, group_base, large_list, bin_rows;
    DevBuf tile_count, tile_offset, tile_cursor, tile_flags, tile_order, tile_list, vis_plane, chunk_cull, pixel_stage;
    BoundsEntry bounds[kBoundsEntries];
    uint64_t bounds_clock = 0;
    uint64_t batch_serial = 0;
    DevCounters *counters = nullptr;
    DevCounters *h_counters = nullptr;      /* pinned */
    uint32_t *h_barrier_timed_out = nullptr; /* pinned, mapped: set by k_frame_barrier when a participant never arrived */
    unsigned long long barrier_timeout_ns = 10000000000ull;

    uint8_t *pinned[2] = { nullptr, nullptr };
    size_t pinned_cap[2] = { 0, 0 };
    cudaEvent_t pinned_ev[2] = { nullptr, nullptr };
    int pinned_next = 0;
    /* stage timing of the batches in flight: a ring, so that a caller that only glFlush()es (frames pipelined behind each
     * other, the host one batch ahead) still gets every batch's kernel times -- a set is folded into the cumulative
     * statistics when its slot comes round again or when the statistics are read */
    struct EvSet {
        cudaEvent_t start = nullptr, stop = nullptr;
        std::vector<cudaEvent_t> stage;     /* 8 per pass: K1 | K2 | count+scan | fill | raster, then [6] end of K4a, [7] end of K4b */
        size_t passes = 0;
        bool pending = false;
    };
    static constexpr int kEvSets = 4;
    EvSet evset[kEvSets];
    int ev_next = 0;
    cudaEvent_t counters_ev = nullptr;      /* the counters of the current pass have reached h_counters */
    bool timed = false;
    uint32_t *present = nullptr;            /* IPC-mapped colour plane of the presenting GPU (mtgl_dev_set_present_target) */
    unsigned long long barrier_epoch = 0;   /* frame barriers this context has taken part in since the plane was exported / mapped */
    cudaEvent_t mark_ev[2] = { nullptr, nullptr };

    /* The optimistic check of the last single-pass batch (tile-list guess, record overflow), deferred: the host does not
     * wait for the scan of frame i before it goes on to queue the upload of frame i+1 (resolve_pending) */
    struct Pending {
        bool active = false;
        BatchDev b; FrameTargets fb; ClearOp clr; uint32_t planes = 0; RasterPlan plan;
        cudaEvent_t sev3 = nullptr, sev4 = nullptr, sev5 = nullptr, sev6 = nullptr, sev7 = nullptr, stop = nullptr;
        uint64_t serial = 0;
        struct Readback { int32_t y0, y1; uint32_t *dst; };
        std::vector<Readback> readbacks;        /* asynchronous read-backs queued behind the batch: re-issued if it has to be redone */
    } pending;

    /* MTGL_PRESENT_COPY: band pushed to the presenting GPU by an asynchronous copy + barrier on a side stream */
    int present_mode = MTGL_PRESENT_STORES;
    cudaStream_t present_stream = nullptr;
    cudaEvent_t present_ev = nullptr, raster_ev = nullptr;
    bool present_busy = false;          /* a push is in flight: colour writers and readers wait for present_ev */

    /* pipelined transfers (mtgl_dev_buffer_data_pinned / mtgl_dev_read_color_async) */
    cudaStream_t upload_stream = nullptr, readback_stream = nullptr;
    cudaEvent_t upload_ev = nullptr, readback_ev = nullptr, render_ev = nullptr;
    bool upload_pending = false, readback_pending = false;
    struct Orphan { uint8_t *ptr; uint64_t size; cudaEvent_t ev; };     /* storage whose last readers are batches before `ev` */
    std::vector<Orphan> orphans;

    mtgl_dev_stats stats{};
    char err[256] = { 0 };
};

namespace {

int fail(mtgl_dev *d, int code, const char *what, cudaError_t ce = cudaSuccess)
{
    if (d) std::snprintf(d->err, sizeof d->err, "%s%s%s", what, ce != cudaSuccess ? ": " : "", ce != cudaSuccess ? cudaGetErrorString(ce) : "");
    return code;
}

#define CU(call)                                                                             \
    do {                                                                                     \
        cudaError_t ce_ = (call);                                                            \
        if (ce_ != cudaSuccess) return fail(d, ce_ == cudaErrorMemoryAllocation ? MTGL_E_OOM : MTGL_E_CUDA, #call, ce_); \
    } while (0)

/* work queued on the main stream from here on runs after the queued uploads have landed and after the queued read-backs
 * have left the colour plane */
int order_after_transfers(mtgl_dev *d)
{
    if (d->upload_pending) { CU(cudaStreamWaitEvent(d->stream, d->upload_ev, 0)); d->upload_pending = false; }
    if (d->readback_pending) { CU(cudaStreamWaitEvent(d->stream, d->readback_ev, 0)); d->readback_pending = false; }
    return MTGL_OK;
}

/* work queued on the main stream from here on sees the plane after the band push / frame barrier in flight on the side stream */
int order_after_present(mtgl_dev *d)
{
    if (d->present_busy) { CU(cudaStreamWaitEvent(d->stream, d->present_ev, 0)); d->present_busy = false; }
    return MTGL_OK;
}

int sync_all_streams(mtgl_dev *d)
{
    if (d->present_stream) CU(cudaStreamSynchronize(d->present_stream));
    d->present_busy = false;
    if (d->upload_stream) CU(cudaStreamSynchronize(d->upload_stream));
    CU(cudaStreamSynchronize(d->stream));
    if (d->readback_stream) CU(cudaStreamSynchronize(d->readback_stream));
    d->upload_pending = false; d->readback_pending = false;
    return MTGL_OK;
}

int reserve(mtgl_dev *d, DevBuf &b, size_t bytes)
{
    if (bytes <= b.cap) return MTGL_OK;
    size_t want = std::max(bytes, b.cap + b.cap / 2);
    want = (want + 255) & ~(size_t)255;
    /* queued kernels may still use the old allocation */
    CU(cudaStreamSynchronize(d->stream));
    if (b.ptr) CU(cudaFree(b.ptr));
    b.ptr = nullptr; b.cap = 0;
    CU(cudaMalloc(&b.ptr, want));
    b.cap = want;
    return MTGL_OK;
}

/* Settle the deferred check of the last optimistic batch.  Must run before anything else is queued on the main stream
 * (every entry point that queues work or waits calls it): when the guessed tile-list capacity was too small the batch's
 * fill and raster kernels have left everything untouched, and they are queued again here, in order. */
int resolve_pending(mtgl_dev *d)
{
    mtgl_dev::Pending &p = d->pending;
    if (!p.active) return MTGL_OK;
    p.active = false;
    CU(cudaEventSynchronize(d->counters_ev));
    const DevCounters hc = *d->h_counters;
    d->stats.triangles_setup = hc.records; d->stats.tile_refs = hc.tile_refs; d->stats.chunks_culled = hc.culled_chunks;
    for (BoundsEntry &e : d->bounds) {          /* did the culling pass pay for the boxes this batch used? */
        if (e.batch != p.serial) continue;
        if (hc.culled_chunks != 0) e.zero_streak = 0;
        else if (++e.zero_streak >= kCullZeroStreak) { e.zero_streak = 0; e.skip_left = kCullSkipBatches; }
    }
    if (hc.overflow) { p.readbacks.clear(); return fail(d, MTGL_E_OOM, "triangle record storage overflow"); }
    if (hc.tile_refs > p.b.list_capacity) {     /* the guess was too small: nothing ran; size exactly and redo */
        CU(cudaStreamSynchronize(d->stream));
        int rc = reserve(d, d->tile_list, (size_t)hc.tile_refs * 4);
        if (rc != MTGL_OK) return rc;
        p.b.tile_list = (uint32_t *)d->tile_list.ptr;
        p.b.list_capacity = (uint32_t)(d->tile_list.cap / 4);
        CU(cudaEventRecord(p.sev3, d->stream));
        launch_bin_fill(p.b, p.fb, d->stream);
        CU(cudaEventRecord(p.sev4, d->stream));
        launch_raster(p.b, p.fb, p.clr, p.planes, p.plan, d->stream, p.sev6, p.sev7);
        CU(cudaEventRecord(p.sev5, d->stream));
        CU(cudaEventRecord(p.stop, d->stream));
        for (const mtgl_dev::Pending::Readback &rb : p.readbacks) {     /* what was copied out before is not the frame */
            const size_t o = (size_t)rb.y0 * d->width, n = (size_t)(rb.y1 - rb.y0) * d->width;
            CU(cudaEventRecord(d->render_ev, d->stream));
            CU(cudaStreamWaitEvent(d->readback_stream, d->render_ev, 0));
            CU(cudaMemcpyAsync(rb.dst + o, d->color + o, n * 4, cudaMemcpyDeviceToHost, d->readback_stream));
            CU(cudaEventRecord(d->readback_ev, d->readback_stream));
            d->readback_pending = true;
        }
    }
    p.readbacks.clear();
    return MTGL_OK;
}

void release(DevBuf &b)
{
    if (b.ptr) cudaFree(b.ptr);
    b.ptr = nullptr; b.cap = 0;
}

uint32_t triangles_of(uint32_t mode, uint32_t n)   /* primitives per draw: loop bounds of flush_* (raster.c:288-296, 961-1231) */
{
    switch (mode) {
    case G_TRIANGLES: return n / 3;
    case G_QUADS: return (n / 4) * 2;
    case G_TRIANGLE_STRIP: case G_TRIANGLE_FAN: case G_POLYGON: return n >= 3 ? n - 2 : 0;
    case G_QUAD_STRIP: return n >= 4 ? ((n - 2) / 2) * 2 : 0;
    case G_POINTS: return n;                                /* raster.c:1044 */
    case G_LINES: return n / 2;                             /* raster.c:293 */
    case G_LINE_STRIP: return n >= 2 ? n - 1 : 0;           /* raster.c:1172-1176 */
    case G_LINE_LOOP: return n >= 2 ? n : 0;                /* raster.c:1185-1195: n - 1 segments + the closing one */
    default: return 0;
    }
}

void build_cfg(const mtgl_dev *d, const mtgl_state &s, RasterCfg &c)
{
    std::memset(&c, 0, sizeof c);
    uint32_t f = 0;
    if (s.caps & MTGL_CAP_DEPTH_TEST) f |= RC_DEPTH_TEST;
    if (s.depth_mask) f |= RC_DEPTH_WRITE;
    if (s.caps & MTGL_CAP_STENCIL_TEST) f |= RC_STENCIL;
    if (s.caps & MTGL_CAP_BLEND) f |= RC_BLEND;
    if (s.caps & MTGL_CAP_ALPHA_TEST) f |= RC_ALPHA_TEST;
    if (s.caps & MTGL_CAP_FOG) f |= RC_FOG;
    if (s.caps & MTGL_CAP_LIGHTING) f |= RC_LIGHTING;
    if (s.shade_model == G_FLAT) f |= RC_FLAT;
    if (s.shade_model == G_PHONG) f |= RC_PHONG;
    if (s.light_model_two_side) f |= RC_TWO_SIDE;
    if (s.perspective_hint != G_FASTEST) f |= RC_PERSPECTIVE;
    if (s.depth_near == 0.0 && s.depth_far == 1.0) f |= RC_DEPTH_RANGE_01;
    bool textured = false;
    /* raster.c:495-498, 618: texturing needs the cap, a bound texture and an uploaded image */
    if ((s.caps & MTGL_CAP_TEXTURE_2D) && s.texture_id != 0 && s.texture_id < kMaxObjects && d->tex[s.texture_id].l0) {
        const TexObj &t = d->tex[s.texture_id];
        f |= RC_TEXTURED;
        textured = true;
        c.tex_l0 = t.l0; c.tex_l1 = t.l1;
        c.tex_f4 = t.f4;
        const bool pow2_w = (t.w & (t.w - 1)) == 0, pow2_h = (t.h & (t.h - 1)) == 0;
        c.fast_tex = (t.f4 && (s.tex_wrap_s != G_REPEAT || pow2_w) && (s.tex_wrap_t != G_REPEAT || pow2_h)) ? 1u : 0u;
        c.tex_w = t.w; c.tex_h = t.h; c.tex_w1 = t.w1; c.tex_h1 = t.h1;
        c.tex_min = s.tex_min_filter; c.tex_mag = s.tex_mag_filter;
        c.tex_wrap_s = s.tex_wrap_s; c.tex_wrap_t = s.tex_wrap_t;
    }
    /* raster.c:640-643 / 712-721: the alpha test (textured only), blending and partial colour masks need in-order shading */
    if (!(s.caps & MTGL_CAP_BLEND) && !((s.caps & MTGL_CAP_ALPHA_TEST) && textured) && (s.color_mask & 0xFu) == 0xFu) f |= RC_DEFER;
    c.flags = f;
    c.depth_func = s.depth_func - G_NEVER; c.alpha_func = s.alpha_func - G_NEVER; c.stencil_func = s.stencil_func - G_NEVER;
    c.stencil_fail = s.stencil_fail; c.stencil_zfail = s.stencil_zfail; c.stencil_zpass = s.stencil_zpass;
    c.stencil_ref = s.stencil_ref; c.stencil_mask = s.stencil_mask; c.stencil_writemask = s.stencil_writemask;
    c.blend_src = s.blend_src; c.blend_dst = s.blend_dst;
    c.color_mask = s.color_mask & 0xFu;
    c.tex_env_mode = s.tex_env_mode; c.fog_mode = s.fog_mode;
    c.alpha_ref = s.alpha_ref;
    c.fog_density = s.fog_density; c.fog_start = s.fog_start; c.fog_end = s.fog_end;
    std::memcpy(c.fog_color, s.fog_color, 16);
    std::memcpy(c.tex_env_color, s.tex_env_color, 16);
    c.depth_near = s.depth_near; c.depth_far = s.depth_far;
}

void describe(const mtgl_dev *d, const mtgl_attrib &a, DevAttrib &o)
{
    std::memset(&o, 0, sizeof o);
    o.enabled = a.enabled;
    if (!a.enabled) return;
    o.stride = a.stride; o.size = a.size; o.type = a.type;
    if (a.buffer != 0 && a.buffer < kMaxBuffers && d->buf[a.buffer].ptr && a.offset <= d->buf[a.buffer].size) {
        o.ptr = d->buf[a.buffer].ptr + a.offset;
        o.avail = d->buf[a.buffer].size - a.offset;
    }
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) & ~(a - 1); }
size_t barrier_offset(size_t pixels) { return align_up(pixels * 4, 256); }     /* byte offset of the frame-barrier line in the colour allocation */

FrameTargets frame_targets(const mtgl_dev *d)
{
    FrameTargets fb;
    fb.color = d->color; fb.depth = d->depth; fb.stencil = d->stencil;
    fb.present = d->present_mode == MTGL_PRESENT_STORES ? d->present : nullptr;
    fb.width = d->width; fb.height = d->height;
    fb.band_y0 = d->band_y0; fb.band_y1 = d->band_y1;
    fb.tiles_x = (d->width + TILE_W - 1) / TILE_W;
    fb.tile_y0 = d->band_y0 >> TILE_LOG;
    fb.tile_rows = (d->band_y1 > d->band_y0) ? (((d->band_y1 - 1) >> TILE_LOG) - fb.tile_y0 + 1) : 0;
    return fb;
}

struct PassDraw { uint32_t draw; uint32_t tri_first, tri_count; };

/* wait for a batch's events and add its times to the statistics (last batch + cumulative) */
int fold_timing(mtgl_dev *d, mtgl_dev::EvSet &es)
{
    CU(cudaEventSynchronize(es.stop));
    float ms = 0.0f;
    CU(cudaEventElapsedTime(&ms, es.start, es.stop));
    mtgl_dev_stats &st = d->stats;
    st.last_batch_ms = ms;
    st.cum_batch_ms += ms;
    st.batches++;
    for (int k = 0; k < 5; k++) st.stage_ms[k] = 0.0f;
    for (int k = 0; k < 3; k++) st.raster_ms[k] = 0.0f;
    for (size_t p = 0; p < es.passes; p++) {
        cudaEvent_t *sev = &es.stage[p * 8];
        for (int k = 0; k < 5; k++) {
            CU(cudaEventElapsedTime(&ms, sev[k], sev[k + 1]));
            st.stage_ms[k] += ms;
        }
        CU(cudaEventElapsedTime(&ms, sev[4], sev[6])); st.raster_ms[0] += ms;
        CU(cudaEventElapsedTime(&ms, sev[6], sev[7])); st.raster_ms[1] += ms;
        CU(cudaEventElapsedTime(&ms, sev[7], sev[5])); st.raster_ms[2] += ms;
    }
    for (int k = 0; k < 5; k++) st.cum_stage_ms[k] += st.stage_ms[k];
    for (int k = 0; k < 3; k++) st.cum_raster_ms[k] += st.raster_ms[k];
    es.pending = false;
    return MTGL_OK;
}

/* Object-space chunk boxes for an array draw (k_cull.cu), or NULL when the draw does not qualify or a culling pass
 * is not expected to pay: the boxes cost one pass over the positions, so they are computed when this device renders
 * only a band of the frame, or when the buffer has already been drawn from unchanged (static geometry: computed once,
 * kept until the buffer is written). */
int chunk_bounds_for(mtgl_dev *d, const mtgl_draw &s, const float4 **out)
{
    *out = nullptr;
    static const bool disabled = std::getenv("MTGL_NO_CULL") != nullptr;       /* A/B switch for profiling and tests */
    if (disabled) return MTGL_OK;
    const mtgl_attrib &a = s.position;
    if (s.mode != G_TRIANGLES || s.source != MTGL_SRC_ARRAYS || s.index_type || !a.enabled || a.type != MTGL_TYPE_F32) return MTGL_OK;
    if (a.buffer == 0 || a.buffer >= kMaxBuffers || a.size < 2 || (a.stride & 3u) || (a.offset & 3u) || s.first < 0) return MTGL_OK;
    BufObj &bo = d->buf[a.buffer];
    const uint32_t ntris = s.count / 3u, nverts = ntris * 3u;
    if (!bo.ptr || ntris < kBoundsMinTriangles || (((uintptr_t)bo.ptr) & 3u)) return MTGL_OK;
    const uint64_t last = a.offset + (uint64_t)((uint32_t)s.first + nverts - 1u) * a.stride + (uint64_t)std::min<uint32_t>(a.size, 3u) * 4u;
    if (last > bo.size) return MTGL_OK;
    const bool band = d->band_y0 > 0 || d->band_y1 < d->height;
    /* boxes an earlier draw of the batch being built refers to are neither evicted nor recomputed: its culling pass has
     * not run yet */
    BoundsEntry *hit = nullptr, *lru = nullptr;
    for (BoundsEntry &e : d->bounds) {
        if (e.buffer == a.buffer && e.gen == bo.gen && e.offset == a.offset && e.stride == a.stride && e.size == a.size &&
            e.first == s.first && e.nverts == nverts && e.boxes.ptr && (!bo.exposed || e.batch == d->batch_serial)) hit = &e;
        if (e.batch != d->batch_serial && (!lru || e.last_use < lru->last_use)) lru = &e;
    }
    const bool is_static = bo.draws_since_write >= 1 && !bo.exposed;
    bo.draws_since_write++;
    if (!hit) {
        if ((!band && !is_static) || !lru) return MTGL_OK;
        const uint32_t nchunks = (ntris + SETUP_THREADS - 1) / SETUP_THREADS;
        int rc = reserve(d, lru->boxes, (size_t)nchunks * 32);
        if (rc != MTGL_OK) return rc;
        launch_chunk_bounds(bo.ptr + a.offset, a.stride, a.size, s.first, nverts, (float4 *)lru->boxes.ptr, d->stream);
        lru->buffer = a.buffer; lru->gen = bo.gen; lru->offset = a.offset; lru->stride = a.stride; lru->size = a.size;
        lru->first = s.first; lru->nverts = nverts;
        lru->zero_streak = 0; lru->skip_left = 0;          /* a new occupant does not inherit the previous one's skip window */
        hit = lru;
    }
    hit->last_use = ++d->bounds_clock;
    if (!band && hit->skip_left > 0) { hit->skip_left--; return MTGL_OK; }      /* recently useless: not this time */
    hit->batch = d->batch_serial;
    *out = (const float4 *)hit->boxes.ptr;
    return MTGL_OK;
}

} // namespace

extern "C" {

int mtgl_dev_abi_version(void) { return MTGL_DEV_ABI_VERSION; }

int mtgl_dev_create(int32_t width, int32_t height, int32_t device, mtgl_dev **out)
{
    if (!out || width <= 0 || height <= 0 || width > 16384 || height > 16384) return MTGL_E_INVALID;   /* framebuffer.h:30-36 */
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return MTGL_E_NO_DEVICE;
    if (device < 0 && cudaGetDevice(&device) != cudaSuccess) return MTGL_E_NO_DEVICE;
    if (device >= count) return MTGL_E_INVALID;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return MTGL_E_NO_DEVICE;
    if (prop.major < 10) return MTGL_E_NO_DEVICE;    /* the kernels are built for sm_100a only */

    mtgl_dev *d = new (std::nothrow) mtgl_dev();
    if (!d) return MTGL_E_OOM;
    d->device = device;
    d->width = width; d->height = height; d->band_y0 = 0; d->band_y1 = height;
    size_t n = (size_t)width * (size_t)height;
    cudaError_t ce = cudaSetDevice(device);
    if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking);
    /* + one 256-byte line behind the last pixel: the frame-barrier counter of a multi-GPU frame (mtgl_dev_frame_barrier),
     * inside the allocation that mtgl_dev_export_color_plane shares, so that one IPC mapping covers both */
    if (ce == cudaSuccess) ce = cudaMalloc(&d->color, barrier_offset(n) + 256);
    if (ce == cudaSuccess) ce = cudaMemset((uint8_t *)d->color + barrier_offset(n), 0, 256);
    if (ce == cudaSuccess) ce = cudaMalloc(&d->depth, n * 4);
    if (ce == cudaSuccess) ce = cudaMalloc(&d->stencil, n);
    if (ce == cudaSuccess) ce = cudaMalloc(&d->unorm8, 256 * sizeof(float));
    if (ce == cudaSuccess) ce = cudaMalloc(&d->counters, sizeof(DevCounters));
    if (ce == cudaSuccess) ce = cudaMallocHost(&d->h_counters, sizeof(DevCounters));
    if (ce == cudaSuccess) ce = cudaMallocHost(&d->h_barrier_timed_out, sizeof(uint32_t));
    if (ce == cudaSuccess) *d->h_barrier_timed_out = 0u;
    if (const char *e = std::getenv("MTGL_BARRIER_TIMEOUT_S")) d->barrier_timeout_ns = (unsigned long long)(std::atof(e) * 1e9);
    for (int i = 0; i < 2 && ce == cudaSuccess; i++) ce = cudaEventCreateWithFlags(&d->pinned_ev[i], cudaEventDisableTiming);
    for (mtgl_dev::EvSet &es : d->evset) {
        if (ce == cudaSuccess) ce = cudaEventCreate(&es.start);
        if (ce == cudaSuccess) ce = cudaEventCreate(&es.stop);
    }
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&d->counters_ev, cudaEventDisableTiming);
    if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&d->present_stream, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&d->present_ev, cudaEventDisableTiming);
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&d->raster_ev, cudaEventDisableTiming);
    if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&d->upload_stream, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&d->readback_stream, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&d->upload_ev, cudaEventDisableTiming);
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&d->readback_ev, cudaEventDisableTiming);
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&d->render_ev, cudaEventDisableTiming);
    for (int i = 0; i < 2 && ce == cudaSuccess; i++) ce = cudaEventCreate(&d->mark_ev[i]);
    if (ce != cudaSuccess) {
        mtgl_dev_destroy(d);
        return ce == cudaErrorMemoryAllocation ? MTGL_E_OOM : MTGL_E_CUDA;
    }
    launch_fill_unorm8(d->unorm8, d->stream);
    *out = d;
    return MTGL_OK;
}

void mtgl_dev_destroy(mtgl_dev *d)
{
    if (!d) return;
    cudaSetDevice(d->device);
    if (d->stream) resolve_pending(d);
    if (d->upload_stream) cudaStreamSynchronize(d->upload_stream);
    if (d->stream) cudaStreamSynchronize(d->stream);
    if (d->readback_stream) cudaStreamSynchronize(d->readback_stream);
    if (d->present_stream) { cudaStreamSynchronize(d->present_stream); cudaStreamDestroy(d->present_stream); }
    if (d->present_ev) cudaEventDestroy(d->present_ev);
    if (d->raster_ev) cudaEventDestroy(d->raster_ev);
    for (mtgl_dev::Orphan &o : d->orphans) { if (o.ptr) cudaFree(o.ptr); if (o.ev) cudaEventDestroy(o.ev); }
    if (d->upload_stream) cudaStreamDestroy(d->upload_stream);
    if (d->readback_stream) cudaStreamDestroy(d->readback_stream);
    if (d->upload_ev) cudaEventDestroy(d->upload_ev);
    if (d->readback_ev) cudaEventDestroy(d->readback_ev);
    if (d->render_ev) cudaEventDestroy(d->render_ev);
    for (uint32_t i = 0; i < kMaxObjects; i++) {
        if (d->tex[i].l0) cudaFree(d->tex[i].l0);
        if (d->tex[i].l1) cudaFree(d->tex[i].l1);
        if (d->tex[i].f4) cudaFree(d->tex[i].f4);
    }
    for (uint32_t i = 0; i < kMaxBuffers; i++)
        if (d->buf[i].ptr) cudaFree(d->buf[i].ptr);
    DevBuf *bufs[] = { &d->arena, &d->v_clip, &d->v_color, &d->v_tex, &d->v_epos, &d->v_enrm, &d->records, &d->rec_eye,
                       &d->group_base, &d->large_list, &d->bin_rows, &d->tile_count, &d->tile_offset, &d->tile_cursor, &d->tile_flags, &d->tile_order, &d->tile_list, &d->vis_plane, &d->chunk_cull, &d->pixel_stage };
    for (DevBuf *b : bufs) release(*b);
    for (BoundsEntry &e : d->bounds) release(e.boxes);
    if (d->color) cudaFree(d->color);
    if (d->depth) cudaFree(d->depth);
    if (d->stencil) cudaFree(d->stencil);
    if (d->unorm8) cudaFree(d->unorm8);
    if (d->counters) cudaFree(d->counters);
    if (d->h_counters) cudaFreeHost(d->h_counters);
    if (d->h_barrier_timed_out) cudaFreeHost(d->h_barrier_timed_out);
    for (int i = 0; i < 2; i++) {
        if (d->pinned[i]) cudaFreeHost(d->pinned[i]);
        if (d->pinned_ev[i]) cudaEventDestroy(d->pinned_ev[i]);
    }
    for (mtgl_dev::EvSet &es : d->evset) {
        if (es.start) cudaEventDestroy(es.start);
        if (es.stop) cudaEventDestroy(es.stop);
        for (cudaEvent_t e : es.stage) cudaEventDestroy(e);
    }
    if (d->counters_ev) cudaEventDestroy(d->counters_ev);
    if (d->present) cudaIpcCloseMemHandle(d->present);
    for (int i = 0; i < 2; i++) if (d->mark_ev[i]) cudaEventDestroy(d->mark_ev[i]);
    if (d->stream) cudaStreamDestroy(d->stream);
    delete d;
}

int mtgl_dev_set_band(mtgl_dev *d, int32_t y0, int32_t y1)
{
    if (!d || y0 < 0 || y1 > d->height || y0 > y1) return MTGL_E_INVALID;
    d->band_y0 = y0; d->band_y1 = y1;
    return MTGL_OK;
}

int mtgl_dev_buffer_data(mtgl_dev *d, uint32_t id, uint64_t size, const void *data)
{
    if (!d || id == 0 || id >= kMaxBuffers) return MTGL_E_INVALID;
    CU(cudaSetDevice(d->device));
    if (int prc_ = resolve_pending(d)) return prc_;
    BufObj &b = d->buf[id];
    b.gen++; b.draws_since_write = 0;
    if (d->upload_pending) CU(cudaStreamSynchronize(d->upload_stream));     /* a queued pinned upload may target this storage */
    if (b.size != size || !b.ptr) {
        b.exposed = false;
        CU(cudaStreamSynchronize(d->stream));
        if (b.ptr) CU(cudaFree(b.ptr));
        b.ptr = nullptr; b.size = 0;
        if (size == 0) return MTGL_OK;
        CU(cudaMalloc(&b.ptr, size));
        b.size = size;
    }
    if (size && data) {
        CU(cudaMemcpyAsync(b.ptr, data, size, cudaMemcpyHostToDevice, d->stream));
        /* the caller may reuse 'data' immediately (glBufferData copies at call time, vbo.c:120-145) */
        CU(cudaStreamSynchronize(d->stream));
    }
    return MTGL_OK;
}

/* the buffer name gets storage of `size` bytes that nobody reads.  The storage it has now may still be read by batches
 * already submitted: it is orphaned behind an event on the main stream (recycled by a later call once the event has
 * passed) */
static int fresh_storage(mtgl_dev *d, BufObj &b, uint64_t size)
{
    if (b.ptr) {
        mtgl_dev::Orphan o{ b.ptr, b.size, nullptr };
        for (mtgl_dev::Orphan &f : d->orphans) if (!f.ptr && f.ev) { o.ev = f.ev; f.ev = nullptr; break; }
        if (!o.ev) CU(cudaEventCreateWithFlags(&o.ev, cudaEventDisableTiming));
        CU(cudaEventRecord(o.ev, d->stream));
        bool placed = false;
        for (mtgl_dev::Orphan &f : d->orphans) if (!f.ptr && !f.ev) { f = o; placed = true; break; }
        if (!placed) d->orphans.push_back(o);
        b.ptr = nullptr; b.size = 0;
    }
    uint8_t *fresh = nullptr;
    for (mtgl_dev::Orphan &f : d->orphans) {
        if (!f.ptr || f.size != size || cudaEventQuery(f.ev) != cudaSuccess) continue;
        fresh = f.ptr; f.ptr = nullptr;         /* (its event stays in the slot for reuse) */
        break;
    }
    cudaGetLastError();                         /* cudaEventQuery reports "not ready" as an error code: not an error here */
    if (!fresh) {
        /* nothing to recycle yet (the first frames of a loop): a plain allocation; orphans of other sizes whose readers
         * have finished are released on the way */
        for (mtgl_dev::Orphan &f : d->orphans)
            if (f.ptr && f.size != size && cudaEventQuery(f.ev) == cudaSuccess) { CU(cudaFree(f.ptr)); f.ptr = nullptr; }
        cudaGetLastError();
        CU(cudaMalloc(&fresh, size));
    }
    b.ptr = fresh; b.size = size;
    b.gen++; b.draws_since_write = 0; b.exposed = false;
    return MTGL_OK;
}

int mtgl_dev_buffer_data_pinned(mtgl_dev *d, uint32_t id, uint64_t size, const void *data)
{
    if (!d || id == 0 || id >= kMaxBuffers) return MTGL_E_INVALID;
    if (size == 0 || !data) return mtgl_dev_buffer_data(d, id, size, data);
    CU(cudaSetDevice(d->device));
    BufObj &b = d->buf[id];
    if (int rc = fresh_storage(d, b, size)) return rc;
    CU(cudaMemcpyAsync(b.ptr, data, size, cudaMemcpyHostToDevice, d->upload_stream));
    CU(cudaEventRecord(d->upload_ev, d->upload_stream));
    d->upload_pending = true;
    return MTGL_OK;
}

int mtgl_dev_buffer_orphan(mtgl_dev *d, uint32_t id, uint64_t size, void **ptr)
{
    if (!d || id == 0 || id >= kMaxBuffers || size == 0 || !ptr) return MTGL_E_INVALID;
    CU(cudaSetDevice(d->device));
    BufObj &b = d->buf[id];
    if (d->upload_pending) CU(cudaStreamSynchronize(d->upload_stream));     /* a queued pinned upload may target the storage being orphaned */
    if (int rc = fresh_storage(d, b, size)) return rc;
    b.exposed = true;                   /* the caller writes the storage directly */
    *ptr = b.ptr;
    return MTGL_OK;
}

int mtgl_dev_read_color_async(mtgl_dev *d, int32_t y0, int32_t y1, uint32_t *color)
{
    if (!d || !color || y0 < 0 || y1 > d->height || y0 > y1) return MTGL_E_INVALID;
    CU(cudaSetDevice(d->device));
    const size_t o = (size_t)y0 * d->width, n = (size_t)(y1 - y0) * d->width;
    if (n == 0) return MTGL_OK;
    if (int prc = order_after_present(d)) return prc;
    CU(cudaEventRecord(d->render_ev, d->stream));
    CU(cudaStreamWaitEvent(d->readback_stream, d->render_ev, 0));
    CU(cudaMemcpyAsync(color + o, d->color + o, n * 4, cudaMemcpyDeviceToHost, d->readback_stream));
    CU(cudaEventRecord(d->readback_ev, d->readback_stream));
    d->readback_pending = true;
    if (d->pending.active) d->pending.readbacks.push_back({ y0, y1, color });     /* re-issued should the batch have to be redone */
    return MTGL_OK;
}

int mtgl_dev_buffer_pointer(mtgl_dev *d, uint32_t id, void **ptr, uint64_t *size)
{
    if (!d || id == 0 || id >= kMaxBuffers) return MTGL_E_INVALID;
    if (ptr) *ptr = d->buf[id].ptr;
    if (size) *size = d->buf[id].size;
    d->buf[id].exposed = true;          /* the caller may write the storage directly (NCCL all-gather of a sharded upload) */
    d->buf[id].gen++;
    return MTGL_OK;
}

int mtgl_dev_buffer_sub_data(mtgl_dev *d, uint32_t id, uint64_t offset, uint64_t size, const void *data)
{
    if (!d || id == 0 || id >= kMaxBuffers || !data) return MTGL_E_INVALID;
    BufObj &b = d->buf[id];
    if (!b.ptr || offset + size > b.size) return MTGL_E_INVALID;
    CU(cudaSetDevice(d->device));
    if (int prc_ = resolve_pending(d)) return prc_;
    b.gen++; b.draws_since_write = 0;
    if (int orc = order_after_transfers(d)) return orc;
    if (size) {
        CU(cudaMemcpyAsync(b.ptr + offset, data, size, cudaMemcpyHostToDevice, d->stream));
        CU(cudaStreamSynchronize(d->stream));
    }
    return MTGL_OK;
}

int mtgl_dev_buffer_delete(mtgl_dev *d, uint32_t id)
{
    if (!d || id == 0 || id >= kMaxBuffers) return MTGL_E_INVALID;
    CU(cudaSetDevice(d->device));
    if (int prc_ = resolve_pending(d)) return prc_;
    if (int orc = order_after_transfers(d)) return orc;
    BufObj &b = d->buf[id];
    if (b.ptr) {
        CU(cudaStreamSynchronize(d->stream));
        CU(cudaFree(b.ptr));
    }
    b.ptr = nullptr; b.size = 0;
    b.gen++; b.draws_since_write = 0; b.exposed = false;
    return MTGL_OK;
}

int mtgl_dev_buffer_read(mtgl_dev *d, uint32_t id, uint64_t offset, uint64_t size, void *out)
{
    if (!d || id == 0 || id >= kMaxBuffers || !out) return MTGL_E_INVALID;
    BufObj &b = d->buf[id];
    if (!b.ptr || offset + size > b.size) return MTGL_E_INVALID;
    CU(cudaSetDevice(d->device));
    if (int prc_ = resolve_pending(d)) return prc_;
    if (int orc = order_after_transfers(d)) return orc;
    if (size) {
        CU(cudaMemcpyAsync(out, b.ptr + offset, size, cudaMemcpyDeviceToHost, d->stream));
        CU(cudaStreamSynchronize(d->stream));
    }
    return MTGL_OK;
}

int mtgl_dev_texture_image(mtgl_dev *d, uint32_t id, int32_t w, int32_t h, const uint32_t *rgba8)
{
    if (!d || id == 0 || id >= kMaxObjects || w <= 0 || h <= 0 || w > 2048 || h > 2048 || !rgba8) return MTGL_E_INVALID;
    CU(cudaSetDevice(d->device));
    if (int prc_ = resolve_pending(d)) return prc_;
    TexObj &t = d->tex[id];
    CU(cudaStreamSynchronize(d->stream));
    if (t.l0) CU(cudaFree(t.l0));
    if (t.l1) CU(cudaFree(t.l1));
    if (t.f4) CU(cudaFree(t.f4));
    t = TexObj();
    size_t n = (size_t)w * (size_t)h;
    CU(cudaMalloc(&t.l0, n * 4));
    t.w = w; t.h = h;
    CU(cudaMemcpyAsync(t.l0, rgba8, n * 4, cudaMemcpyHostToDevice, d->stream));
    if (w >= 2 && h >= 2) {      /* textures.c:317: level 1 needs both dimensions >= 2 */
        t.w1 = w / 2; t.h1 = h / 2;
        CU(cudaMalloc(&t.l1, (size_t)t.w1 * t.h1 * 4));
        launch_mip1(t.l0, w, h, t.l1, d->unorm8, d->stream);
    }
    if (n <= (size_t)kFloatTexelsMax) {
        const int n1 = t.l1 ? t.w1 * t.h1 : 0;
        CU(cudaMalloc(&t.f4, (n + (size_t)n1) * sizeof(float4)));
        launch_tex_f4(t.l0, (int)n, t.l1, n1, t.f4, d->unorm8, d->stream);
    }
    CU(cudaStreamSynchronize(d->stream));
    return MTGL_OK;
}

int mtgl_dev_texture_delete(mtgl_dev *d, uint32_t id)
{
    if (!d || id == 0 || id >= kMaxObjects) return MTGL_E_INVALID;
    CU(cudaSetDevice(d->device));
    if (int prc_ = resolve_pending(d)) return prc_;
    TexObj &t = d->tex[id];
    if (t.l0 || t.l1) CU(cudaStreamSynchronize(d->stream));
    if (t.l0) CU(cudaFree(t.l0));
    if (t.l1) CU(cudaFree(t.l1));
    if (t.f4) CU(cudaFree(t.f4));
    t = TexObj();
    return MTGL_OK;
}

int mtgl_dev_submit(mtgl_dev *d, const mtgl_batch *bt)
{
    if (!d || !bt) return MTGL_E_INVALID;
    /* MTGL_TRACE_SUBMIT=1: host microseconds from entry to (a) the arena copy queued, (b) the last launch queued */
    static const bool trace = std::getenv("MTGL_TRACE_SUBMIT") != nullptr;
    const auto t_in = std::chrono::steady_clock::now();
    auto t_copy = t_in, t_launched = t_in;
    struct TraceOut {
        const bool on; const std::chrono::steady_clock::time_point &t0, &t1, &t2;
        ~TraceOut() {
            if (!on) return;
            const auto t3 = std::chrono::steady_clock::now();
            auto us = [&](const std::chrono::steady_clock::time_point &t) { return std::chrono::duration<double, std::micro>(t - t0).count(); };
            static std::chrono::steady_clock::time_point last_out;
            std::fprintf(stderr, "[mtgl submit] since last return %.1f us | arena queued +%.1f, kernels queued +%.1f, counters read +%.1f us\n",
                         std::chrono::duration<double, std::micro>(t0 - last_out).count(), us(t1), us(t2), us(t3));
            last_out = std::chrono::steady_clock::now();
        }
    } trace_out{ trace, t_in, t_copy, t_launched };
    CU(cudaSetDevice(d->device));
    if (int prc_ = resolve_pending(d)) return prc_;
    if (int orc = order_after_transfers(d)) return orc;
    const FrameTargets fb = frame_targets(d);
    const uint32_t ntiles = (uint32_t)(fb.tiles_x * fb.tile_rows);

    d->batch_serial++;
    /* ---- validate + per-draw prefix tables ---- */
    std::vector<DevDraw> draws;
    draws.reserve(bt->n_draws);
    uint32_t planes = 0;
    bool need_eye = false;
    if (bt->n_states > STATE_INDEX_MASK) return fail(d, MTGL_E_INVALID, "too many state blocks in one batch");
    for (uint32_t i = 0; i < bt->n_draws; i++) {
        const mtgl_draw &s = bt->draws[i];
        if (s.raster_state >= bt->n_states) return fail(d, MTGL_E_INVALID, "draw refers to a missing state block");
        if (s.source == MTGL_SRC_STAGED && (uint64_t)s.first_staged + s.count > bt->n_vertices) return fail(d, MTGL_E_INVALID, "draw exceeds the staged vertices");
        if (s.source == MTGL_SRC_ARRAYS && s.vertex_state >= bt->n_states) return fail(d, MTGL_E_INVALID, "draw refers to a missing vertex state");
        DevDraw o;
        std::memset(&o, 0, sizeof o);
        o.mode = s.mode; o.count = s.count; o.raster_state = s.raster_state; o.source = s.source;
        o.first_staged = s.first_staged; o.vertex_state = s.vertex_state; o.first = s.first; o.index_type = s.index_type;
        std::memcpy(o.cur_color, s.cur_color, 16); std::memcpy(o.cur_texcoord, s.cur_texcoord, 8); std::memcpy(o.cur_normal, s.cur_normal, 12);
        describe(d, s.position, o.position); describe(d, s.color, o.color);
        describe(d, s.texcoord, o.texcoord); describe(d, s.normal, o.normal);
        o.ntris = triangles_of(s.mode, s.count);
        o.state_max = bt->n_states ? bt->n_states - 1u : 0u;
        if (int brc = chunk_bounds_for(d, s, &o.bounds)) return brc;
        draws.push_back(o);
        const mtgl_state &rs = bt->states[s.raster_state];
        planes |= 1u;
        if (rs.caps & MTGL_CAP_DEPTH_TEST) planes |= 2u;
        if (rs.caps & MTGL_CAP_STENCIL_TEST) planes |= 4u;
        if ((rs.caps & MTGL_CAP_LIGHTING) && (rs.shade_model == G_PHONG || rs.light_model_two_side)) need_eye = true;
    }
    bool outline_modes = false;
    for (uint32_t i = 0; i < bt->n_states; i++)
        if (bt->states[i].polygon_mode_front != G_FILL || bt->states[i].polygon_mode_back != G_FILL) outline_modes = true;
    if (bt->clear_mask & G_COLOR_BUFFER_BIT) planes |= 1u;
    if (bt->clear_mask & G_DEPTH_BUFFER_BIT) planes |= 2u;
    if (bt->clear_mask & G_STENCIL_BUFFER_BIT) planes |= 4u;
    if (planes == 0) return MTGL_OK;

    /* ---- split into passes that fit the worst-case record storage ---- */
    std::vector<std::vector<PassDraw>> passes(1);
    {
        uint32_t room = kMaxTrisPerPass;
        for (uint32_t i = 0; i < draws.size(); i++) {
            uint32_t left = draws[i].ntris, first = 0;
            while (left) {
                if (room == 0) { passes.emplace_back(); room = kMaxTrisPerPass; }
                uint32_t take = std::min(left, room);
                passes.back().push_back({ i, first, take });
                first += take; left -= take; room -= take;
            }
        }
    }

    /* ---- batch arena: states | cfgs | staged | blob (shared by all passes) ---- */
    std::vector<RasterCfg> cfgs(bt->n_states);
    for (uint32_t i = 0; i < bt->n_states; i++) build_cfg(d, bt->states[i], cfgs[i]);
    /* The unordered class (k_vis.cu): deferrable states with depth test + write, no stencil test and the depth
     * function of the first such state, when that is one of LESS / LEQUAL / GREATER / GEQUAL. */
    uint32_t unordered_func = 0;
    bool unordered_range01 = true;
    for (uint32_t i = 0; i < bt->n_states; i++) {
        RasterCfg &c = cfgs[i];
        const uint32_t need = RC_DEFER | RC_DEPTH_TEST | RC_DEPTH_WRITE;
        if ((c.flags & need) != need || (c.flags & RC_STENCIL)) continue;
        const uint32_t fn = c.depth_func;
        if (fn != 1u && fn != 3u && fn != 4u && fn != 6u) continue;
        if (!unordered_func) unordered_func = fn;
        if (fn != unordered_func) continue;
        c.flags |= RC_UNORDERED;
        if (!(c.flags & RC_DEPTH_RANGE_01)) unordered_range01 = false;
    }
    const size_t sz_states = align_up((size_t)bt->n_states * sizeof(mtgl_state), 256);
    const size_t sz_cfgs = align_up((size_t)bt->n_states * sizeof(RasterCfg), 256);
    const size_t sz_staged = align_up((size_t)bt->n_vertices * sizeof(mtgl_in_vertex), 256);
    const size_t sz_blob = align_up((size_t)bt->blob_size, 256);
    size_t max_pass_draws = 0;
    for (auto &p : passes) max_pass_draws = std::max(max_pass_draws, p.size());
    const size_t sz_draws = align_up(max_pass_draws * sizeof(DevDraw), 256);
    const size_t sz_prefix = align_up((max_pass_draws + 1) * sizeof(uint32_t), 256);
    const size_t arena_fixed = sz_states + sz_cfgs + sz_staged + sz_blob;
    const size_t arena_pass = sz_draws + 2 * sz_prefix;
    const size_t arena_total = arena_fixed + arena_pass * passes.size();

    int rc = reserve(d, d->arena, arena_total);
    if (rc != MTGL_OK) return rc;
    const int slot = d->pinned_next;
    d->pinned_next ^= 1;
    CU(cudaEventSynchronize(d->pinned_ev[slot]));       /* the previous copy out of this staging buffer is done */
    if (d->pinned_cap[slot] < arena_total) {
        if (d->pinned[slot]) CU(cudaFreeHost(d->pinned[slot]));
        d->pinned[slot] = nullptr; d->pinned_cap[slot] = 0;
        size_t want = align_up(arena_total + arena_total / 2, 4096);
        CU(cudaMallocHost(&d->pinned[slot], want));
        d->pinned_cap[slot] = want;
    }
    uint8_t *hp = d->pinned[slot];
    uint8_t *dp = (uint8_t *)d->arena.ptr;
    size_t off = 0;
    const size_t o_states = off; if (bt->n_states) std::memcpy(hp + off, bt->states, (size_t)bt->n_states * sizeof(mtgl_state)); off += sz_states;
    const size_t o_cfgs = off; if (bt->n_states) std::memcpy(hp + off, cfgs.data(), (size_t)bt->n_states * sizeof(RasterCfg)); off += sz_cfgs;
    const size_t o_staged = off; if (bt->n_vertices) std::memcpy(hp + off, bt->vertices, (size_t)bt->n_vertices * sizeof(mtgl_in_vertex)); off += sz_staged;
    const size_t o_blob = off; if (bt->blob_size) std::memcpy(hp + off, bt->blob, (size_t)bt->blob_size); off += sz_blob;

    struct PassInfo { size_t o_draws, o_vbase, o_tbase; uint32_t n_draws, n_vertices, n_triangles, n_unfused, real_triangles; bool any_bounds; };
    static const bool fuse_triangles = std::getenv("MTGL_NO_FUSE") == nullptr;     /* A/B switch for profiling */
    const bool share_indexed = std::getenv("MTGL_NO_SHARED_VERTS") == nullptr;      /* A/B switch (read per batch: the tests flip it) */
    std::vector<PassInfo> infos;
    for (auto &p : passes) {
        PassInfo pi{};
        pi.o_draws = off; off += sz_draws;
        pi.o_vbase = off; off += sz_prefix;
        pi.o_tbase = off; off += sz_prefix;
        DevDraw *pd = reinterpret_cast<DevDraw *>(hp + pi.o_draws);
        uint32_t *vb = reinterpret_cast<uint32_t *>(hp + pi.o_vbase), *tb = reinterpret_cast<uint32_t *>(hp + pi.o_tbase);
        uint32_t v = 0, t = 0, k = 0, unfused = 0;
        for (const PassDraw &q : p) {
            /* a draw of several chunks starts on a chunk boundary and leaves the rest of its last chunk empty, so that no
             * 256-triangle chunk of k_setup mixes two such draws (one fast attribute path, one state block, one set of chunk
             * boxes per CTA): many draws of one mesh under changing matrices cost a few idle slots instead of the slow path */
            const bool own_chunks = q.tri_count >= 2u * SETUP_THREADS;
            if (own_chunks) t = (uint32_t)align_up(t, SETUP_THREADS);
            if (draws[q.draw].bounds) pi.any_bounds = true;
            DevDraw o = draws[q.draw];
            const mtgl_draw &s = bt->draws[q.draw];
            if (o.index_type) {
                if (s.index_buffer) {
                    if (s.index_buffer < kMaxBuffers && d->buf[s.index_buffer].ptr && s.index_offset <= d->buf[s.index_buffer].size) {
                        o.index_ptr = d->buf[s.index_buffer].ptr + s.index_offset;
                        o.index_avail = d->buf[s.index_buffer].size - s.index_offset;
                    }
                } else if (s.index_offset <= bt->blob_size) {
                    o.index_ptr = dp + o_blob + s.index_offset;
                    o.index_avail = bt->blob_size - s.index_offset;
                }
            }
            o.vbase = v;
            o.fused = (fuse_triangles && o.mode == G_TRIANGLES) ? 1u : 0u;
            /* glDrawElements re-emits a vertex for every index (gl_api.c:1898-1939); the vertex stage is a pure function of
             * the element, so when the arrays hold fewer elements than the draw has indices, every ELEMENT is transformed
             * and lit once and set-up looks vertices up by index (a post-transform cache that cannot miss).  Elements
             * = the longest enabled array; indices beyond it read as the defaults, like fetch_attrib does per attribute. */
            if (o.index_type && o.index_ptr && s.source == MTGL_SRC_ARRAYS && share_indexed) {
                uint64_t elems = 0;
                for (const DevAttrib *a : { &o.position, &o.color, &o.texcoord, &o.normal }) {
                    if (!a->enabled || !a->ptr || !a->stride) continue;
                    const uint64_t bytes = (uint64_t)a->size * (a->type == MTGL_TYPE_F32 ? 4u : 1u);
                    if (a->avail >= bytes) elems = std::max<uint64_t>(elems, (a->avail - bytes) / a->stride + 1u);
                }
                if (elems > 0 && elems + 1 <= (uint64_t)o.count && elems < (1u << 24)) { o.shared_verts = (uint32_t)elems; o.fused = 0u; }
            }
            if (!o.fused) unfused++;
            o.tbase = t - q.tri_first;      /* triangle k of the draw has global index tbase + k */
            o.ntris = q.tri_count;
            vb[k] = v; tb[k] = t;
            fast_draw_init(o.fast, reinterpret_cast<const mtgl_state *>(dp + o_states), o, k, t, t + q.tri_count);
            pd[k++] = o;
            v += o.shared_verts ? o.shared_verts + 1u : o.count; t += q.tri_count;
            pi.real_triangles += q.tri_count;
            if (own_chunks) t = (uint32_t)align_up(t, SETUP_THREADS);
        }
        vb[k] = v; tb[k] = t;
        pi.n_draws = k; pi.n_vertices = v; pi.n_triangles = t; pi.n_unfused = unfused;
        infos.push_back(pi);
    }
    /* the pinned staging buffer is device-mapped (unified addressing): small arenas are read by a kernel, large ones
     * (immediate-mode batches with millions of staged vertices) go through the copy engine.  Sizes are multiples of 256. */
    /* the tables the (first) pass starts from zero are cleared by the same kernel: one launch less per frame */
    bool zeroed_by_upload = false;
    if (arena_total <= kKernelUploadMax) {
        void *zero = nullptr;
        size_t zero_bytes = 0;
        if (!infos.empty() && infos[0].n_triangles > 0 && ntiles > 0) {
            if ((rc = reserve(d, d->tile_count, 256 + (size_t)ntiles * 8))) return rc;
            zero = d->tile_count.ptr; zero_bytes = align_up(256 + (size_t)ntiles * 8, 16);
            zeroed_by_upload = true;
        }
        launch_upload(hp, dp, arena_total, zero, zero_bytes, d->stream);
    } else CU(cudaMemcpyAsync(dp, hp, arena_total, cudaMemcpyHostToDevice, d->stream));
    CU(cudaEventRecord(d->pinned_ev[slot], d->stream));
    t_copy = std::chrono::steady_clock::now();

    mtgl_dev::EvSet &es = d->evset[d->ev_next];
    d->ev_next = (d->ev_next + 1) % mtgl_dev::kEvSets;
    if (es.pending && (rc = fold_timing(d, es))) return rc;     /* four batches ago: long finished */
    CU(cudaEventRecord(es.start, d->stream));
    d->timed = true;
    uint64_t tot_v = 0, tot_t = 0, tot_r = 0, tot_refs = 0, tot_culled = 0;
    while (es.stage.size() < passes.size() * 8) {
        cudaEvent_t e;
        CU(cudaEventCreate(&e));
        es.stage.push_back(e);
    }
    es.passes = passes.size();

    for (size_t pidx = 0; pidx < passes.size(); pidx++) {
        const PassInfo &pi = infos[pidx];
        BatchDev b;
        std::memset(&b, 0, sizeof b);
        b.states = reinterpret_cast<const mtgl_state *>(dp + o_states);
        b.cfgs = reinterpret_cast<const RasterCfg *>(dp + o_cfgs);
        b.n_states = bt->n_states;
        b.staged = reinterpret_cast<const mtgl_in_vertex *>(dp + o_staged);
        b.draws = reinterpret_cast<const DevDraw *>(dp + pi.o_draws);
        b.draw_vbase = reinterpret_cast<const uint32_t *>(dp + pi.o_vbase);
        b.draw_tbase = reinterpret_cast<const uint32_t *>(dp + pi.o_tbase);
        b.n_draws = pi.n_draws; b.n_vertices = pi.n_vertices; b.n_triangles = pi.n_triangles;
        b.n_unfused_draws = pi.n_unfused;
        b.need_eye = need_eye ? 1u : 0u;
        b.unorm8 = d->unorm8;
        b.counters = d->counters;

        cudaEvent_t *sev = &es.stage[pidx * 8];
        bool optimistic = false, had_triangles = false;
        CU(cudaEventRecord(sev[0], d->stream));
        ClearOp clr;
        std::memset(&clr, 0, sizeof clr);
        if (pidx == 0 && bt->clear_mask) {
            clr.mask = bt->clear_mask;
            clr.x0 = bt->clear_rect[0]; clr.y0 = bt->clear_rect[1]; clr.x1 = bt->clear_rect[2]; clr.y1 = bt->clear_rect[3];
            clr.color = bt->clear_color; clr.depth = bt->clear_depth; clr.stencil = bt->clear_stencil;
        }

        if (pi.n_triangles > 0 && ntiles > 0) {
            const size_t nv = pi.n_vertices;
            const uint32_t chunks = (pi.n_triangles + SETUP_THREADS - 1) / SETUP_THREADS;
            /* worst case per primitive: 7 fan sub-triangles of a clipped triangle, each 3 outline records in GL_LINE/GL_POINT mode */
            const size_t rec_cap = (size_t)pi.n_triangles * (outline_modes ? 21 : 7);
            if ((rc = reserve(d, d->v_clip, nv * 16)) || (rc = reserve(d, d->v_color, nv * 16)) || (rc = reserve(d, d->v_tex, nv * 16))) return rc;
            if (need_eye && ((rc = reserve(d, d->v_epos, nv * 16)) || (rc = reserve(d, d->v_enrm, nv * 16)))) return rc;
            if ((rc = reserve(d, d->records, rec_cap * sizeof(TriRecord)))) return rc;
            if (need_eye && (rc = reserve(d, d->rec_eye, rec_cap * sizeof(TriEye)))) return rc;
            if ((rc = reserve(d, d->group_base, (size_t)chunks * (SETUP_THREADS / 32) * 4)) || (rc = reserve(d, d->large_list, rec_cap * 4)) ||
                (rc = reserve(d, d->bin_rows, rec_cap * 16))) return rc;
            /* everything a pass starts from zero lives in one allocation (one memset per frame, not three):
             * counters (256 B) | tile_count | tile_flags */
            if ((rc = reserve(d, d->tile_count, 256 + (size_t)ntiles * 8)) || (rc = reserve(d, d->tile_offset, (size_t)ntiles * 4)) ||
                (rc = reserve(d, d->tile_cursor, (size_t)ntiles * 4)) ||
                (rc = reserve(d, d->tile_order, (size_t)ntiles * 4))) return rc;
            if ((rc = reserve(d, d->vis_plane, (size_t)d->width * d->height * 4))) return rc;
            if (pi.any_bounds && (rc = reserve(d, d->chunk_cull, chunks))) return rc;
            b.chunk_cull = pi.any_bounds ? (uint8_t *)d->chunk_cull.ptr : nullptr;
            b.v_clip = (float4 *)d->v_clip.ptr; b.v_color = (float4 *)d->v_color.ptr; b.v_tex = (float4 *)d->v_tex.ptr;
            b.v_epos = (float4 *)d->v_epos.ptr; b.v_enrm = (float4 *)d->v_enrm.ptr;
            b.records = (TriRecord *)d->records.ptr; b.rec_eye = (TriEye *)d->rec_eye.ptr;
            b.record_capacity = (uint32_t)std::min<size_t>(rec_cap, 0xFFFFFFFFu);
            b.group_base = (uint32_t *)d->group_base.ptr; b.large_list = (uint32_t *)d->large_list.ptr;
            b.bin_rows = (uint4 *)d->bin_rows.ptr;
            b.counters = (DevCounters *)d->tile_count.ptr;
            b.host_counters = d->h_counters;
            b.tile_count = (uint32_t *)((uint8_t *)d->tile_count.ptr + 256); b.tile_offset = (uint32_t *)d->tile_offset.ptr;
            b.tile_cursor = (uint32_t *)d->tile_cursor.ptr;
            b.tile_flags = b.tile_count + ntiles;
            b.tile_order = (uint32_t *)d->tile_order.ptr;
            b.vis_plane = (uint32_t *)d->vis_plane.ptr;

            if (!(pidx == 0 && zeroed_by_upload)) CU(cudaMemsetAsync(d->tile_count.ptr, 0, 256 + (size_t)ntiles * 8, d->stream));
            launch_vertex_stage(b, d->stream);
            launch_chunk_cull(b, fb, d->stream);
            CU(cudaEventRecord(sev[1], d->stream));
            launch_setup(b, fb, d->stream);
            CU(cudaEventRecord(sev[2], d->stream));
            launch_bin_scan(b, fb, d->stream);
            CU(cudaEventRecord(sev[3], d->stream));
            /* The list length is only known on the device.  Single-pass batches (the normal frame) run optimistically:
             * the list buffer is sized from a guess that only ever grows, the fill and raster kernels refuse to run
             * when the scan result does not fit (BatchDev::guard), and the host checks the counters AFTER queueing
             * everything -- so the GPU never idles waiting for the host; a miss re-queues fill + raster.  Multi-pass
             * batches read the count back first. */
            /* (k_bin_scan has written the counters into d->h_counters, BatchDev::host_counters) */
            static const bool sync_lists = std::getenv("MTGL_SYNC_LISTS") != nullptr;      /* A/B switch for profiling */
            optimistic = passes.size() == 1 && !sync_lists;
            if (optimistic) {
                CU(cudaEventRecord(d->counters_ev, d->stream));
                size_t guess = std::max<size_t>(d->tile_list.cap / 4, (size_t)pi.n_triangles * 2 + 65536);
                if (const char *e = std::getenv("MTGL_LIST_GUESS")) guess = std::max<size_t>(d->tile_list.cap / 4, (size_t)std::atoll(e));   /* test knob: force the re-queue path */
                if ((rc = reserve(d, d->tile_list, guess * 4))) return rc;
                b.guard = 1u;
            } else {
                CU(cudaStreamSynchronize(d->stream));
                if (d->h_counters->overflow) return fail(d, MTGL_E_OOM, "triangle record storage overflow");
                if ((rc = reserve(d, d->tile_list, std::max<size_t>((size_t)d->h_counters->tile_refs * 4, 1024)))) return rc;
            }
            b.tile_list = (uint32_t *)d->tile_list.ptr;
            b.list_capacity = (uint32_t)(d->tile_list.cap / 4);
            if (optimistic || d->h_counters->tile_refs) launch_bin_fill(b, fb, d->stream);
            had_triangles = true;
        } else {
            CU(cudaEventRecord(sev[1], d->stream));
            CU(cudaEventRecord(sev[2], d->stream));
            CU(cudaEventRecord(sev[3], d->stream));
        }
        CU(cudaEventRecord(sev[4], d->stream));
        bool any_defer = false, any_in_order = false, any_ordered_vis = false, any_not_plain = false;
        uint32_t flags_all = 0xFFFFFFFFu, flags_any = 0u;
        for (const PassDraw &q : passes[pidx]) {
            const mtgl_draw &dq = bt->draws[q.draw];
            const mtgl_state &sq = bt->states[dq.raster_state];
            uint32_t cf = cfgs[dq.raster_state].flags;
            cf |= fill_pseudo_flags(cfgs[dq.raster_state]);         /* (for launch_fill's choice of kernel instance) */
            const bool filled = dq.mode >= G_TRIANGLES && sq.polygon_mode_front == G_FILL && sq.polygon_mode_back == G_FILL;
            if ((cf & RC_DEFER) && filled) any_defer = true; else any_in_order = true;
            if ((cf & RC_DEFER) && !filled && dq.mode >= G_TRIANGLES) any_defer = true;   /* mixed fill/outline faces */
            if ((cf & RC_DEFER) && !(cf & RC_UNORDERED) && dq.mode >= G_TRIANGLES) any_ordered_vis = true;
            if ((cf & RC_DEFER) || !filled) any_not_plain = true;
            flags_all &= cf; flags_any |= cf;
        }
        RasterPlan plan;
        plan.any_deferrable = any_defer && pi.n_triangles > 0;
        plan.any_ordered_vis = any_ordered_vis;
        plan.any_in_order = any_in_order;
        plan.plain_in_order = any_in_order && !any_not_plain;
        plan.unordered_func = unordered_func;
        plan.unordered_range01 = unordered_range01;
        plan.color_gate = nullptr;
        if (d->present_busy) { plan.color_gate = d->present_ev; d->present_busy = false; }      /* (the stream waits inside launch_raster) */
        /* the pixel-owner kernel for in-order tiles of large triangles (k_fill.cu); it has no per-fragment lighting.
         * MTGL_FILL=never|always: A/B switch for profiling and tests (always: every eligible in-order tile, whatever its triangles' size) */
        uint32_t fill_env = FILL_AUTO;       /* read per batch (not cached): the tests switch it between frames */
        if (const char *e = std::getenv("MTGL_FILL")) fill_env = !std::strcmp(e, "never") ? FILL_OFF : (!std::strcmp(e, "always") ? FILL_ALWAYS : FILL_AUTO);
        plan.fill_mode = (need_eye || !had_triangles) ? FILL_OFF : fill_env;
        plan.in_order_all = flags_all; plan.in_order_any = flags_any;
        launch_raster(b, fb, clr, planes, plan, d->stream, sev[6], sev[7]);
        CU(cudaEventRecord(sev[5], d->stream));
        t_launched = std::chrono::steady_clock::now();
        if (had_triangles) {
            tot_v += pi.n_vertices; tot_t += pi.real_triangles;
            if (optimistic) {
                /* The scan's counters are looked at later (resolve_pending): at the next call that queues work or waits.
                 * The host is free to queue the next frame's upload in the meantime. */
                mtgl_dev::Pending &pd = d->pending;
                pd.active = true; pd.b = b; pd.fb = fb; pd.clr = clr; pd.planes = planes; pd.plan = plan; pd.plan.color_gate = nullptr;
                pd.sev3 = sev[3]; pd.sev4 = sev[4]; pd.sev5 = sev[5]; pd.sev6 = sev[6]; pd.sev7 = sev[7]; pd.stop = es.stop;
                pd.serial = d->batch_serial;
                pd.readbacks.clear();
            } else {
                tot_r += d->h_counters->records; tot_refs += d->h_counters->tile_refs; tot_culled += d->h_counters->culled_chunks;
            }
        }
    }
    CU(cudaEventRecord(es.stop, d->stream));
    es.pending = true;      /* only a batch that was queued completely is folded into the statistics */
    CU(cudaGetLastError());
    d->stats.vertices = tot_v; d->stats.triangles_in = tot_t;
    if (!d->pending.active) {
        d->stats.triangles_setup = tot_r; d->stats.tile_refs = tot_refs; d->stats.chunks_culled = tot_culled;
        for (BoundsEntry &e : d->bounds) {          /* did the culling pass pay for the boxes this batch used? */
            if (e.batch != d->batch_serial) continue;
            if (tot_culled != 0) e.zero_streak = 0;
            else if (++e.zero_streak >= kCullZeroStreak) { e.zero_streak = 0; e.skip_left = kCullSkipBatches; }
        }
    }
    return MTGL_OK;
}

int mtgl_dev_finish(mtgl_dev *d)
{
    if (!d) return MTGL_E_INVALID;
    CU(cudaSetDevice(d->device));
    if (int prc_ = resolve_pending(d)) return prc_;
    if (int rc = sync_all_streams(d)) return rc;
    if (*d->h_barrier_timed_out) {
        *d->h_barrier_timed_out = 0u;
        return fail(d, MTGL_E_CUDA, "frame barrier timed out: a participant of the frame never arrived (MTGL_BARRIER_TIMEOUT_S)");
    }
    return MTGL_OK;
}

int mtgl_dev_read_framebuffer(mtgl_dev *d, int32_t y0, int32_t y1, uint32_t *color, float *depth, uint8_t *stencil)
{
    if (!d || y0 < 0 || y1 > d->height || y0 > y1) return MTGL_E_INVALID;
    CU(cudaSetDevice(d->device));
    if (int prc_ = resolve_pending(d)) return prc_;
    size_t o = (size_t)y0 * d->width, n = (size_t)(y1 - y0) * d->width;
    if (n == 0) return MTGL_OK;
    if (int prc = order_after_present(d)) return prc;
    if (color) CU(cudaMemcpyAsync(color + o, d->color + o, n * 4, cudaMemcpyDeviceToHost, d->stream));
    if (depth) CU(cudaMemcpyAsync(depth + o, d->depth + o, n * 4, cudaMemcpyDeviceToHost, d->stream));
    if (stencil) CU(cudaMemcpyAsync(stencil + o, d->stencil + o, n, cudaMemcpyDeviceToHost, d->stream));
    CU(cudaStreamSynchronize(d->stream));
    return MTGL_OK;
}

int mtgl_dev_write_framebuffer(mtgl_dev *d, int32_t y0, int32_t y1, const uint32_t *color, const float *depth, const uint8_t *stencil)
{
    if (!d || y0 < 0 || y1 > d->height || y0 > y1) return MTGL_E_INVALID;
    CU(cudaSetDevice(d->device));
    if (int prc_ = resolve_pending(d)) return prc_;
    size_t o = (size_t)y0 * d->width, n = (size_t)(y1 - y0) * d->width;
    if (n == 0) return MTGL_OK;
    if (int orc = order_after_transfers(d)) return orc;
    if (int prc = order_after_present(d)) return prc;
    if (color) CU(cudaMemcpyAsync(d->color + o, color + o, n * 4, cudaMemcpyHostToDevice, d->stream));
    if (color && d->present) CU(cudaMemcpyAsync(d->present + o, color + o, n * 4, cudaMemcpyHostToDevice, d->stream));
    if (depth) CU(cudaMemcpyAsync(d->depth + o, depth + o, n * 4, cudaMemcpyHostToDevice, d->stream));
    if (stencil) CU(cudaMemcpyAsync(d->stencil + o, stencil + o, n, cudaMemcpyHostToDevice, d->stream));
    CU(cudaStreamSynchronize(d->stream));
    return MTGL_OK;
}

int mtgl_dev_plane_pointers(mtgl_dev *d, void **color, void **depth, void **stencil)
{
    if (!d) return MTGL_E_INVALID;
    if (color) *color = d->color;
    if (depth) *depth = d->depth;
    if (stencil) *stencil = d->stencil;
    return MTGL_OK;
}

int mtgl_dev_export_color_plane(mtgl_dev *d, void *handle_out)
{
    static_assert(sizeof(cudaIpcMemHandle_t) <= MTGL_IPC_HANDLE_BYTES, "IPC handle size");
    if (!d || !handle_out) return MTGL_E_INVALID;
    CU(cudaSetDevice(d->device));
    if (int prc_ = resolve_pending(d)) return prc_;
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, d->color));
    CU(cudaStreamSynchronize(d->stream));
    CU(cudaMemset((uint8_t *)d->color + barrier_offset((size_t)d->width * d->height), 0, 256));   /* a fresh barrier counter for the new group */
    d->barrier_epoch = 0;
    std::memset(handle_out, 0, MTGL_IPC_HANDLE_BYTES);
    std::memcpy(handle_out, &h, sizeof h);
    return MTGL_OK;
}

int mtgl_dev_set_present_target(mtgl_dev *d, const void *handle)
{
    if (!d) return MTGL_E_INVALID;
    CU(cudaSetDevice(d->device));
    if (int prc_ = resolve_pending(d)) return prc_;
    CU(cudaStreamSynchronize(d->stream));
    if (d->present) { CU(cudaIpcCloseMemHandle(d->present)); d->present = nullptr; }
    d->barrier_epoch = 0;
    if (!handle) return MTGL_OK;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof h);
    void *p = nullptr;
    CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    d->present = static_cast<uint32_t *>(p);
    return MTGL_OK;
}

int mtgl_dev_set_present_mode(mtgl_dev *d, int mode)
{
    if (!d || (mode != MTGL_PRESENT_STORES && mode != MTGL_PRESENT_COPY)) return MTGL_E_INVALID;
    CU(cudaSetDevice(d->device));
    if (int prc_ = resolve_pending(d)) return prc_;
    if (int rc = sync_all_streams(d)) return rc;
    d->present_mode = mode;
    return MTGL_OK;
}

static uint32_t pixel_bpp(uint32_t format)
{
    switch (format) {
    case 0x1908: return 4;      /* GL_RGBA */
    case 0x1907: return 3;      /* GL_RGB */
    case 0x1909: return 1;      /* GL_LUMINANCE */
    case 0x190A: return 2;      /* GL_LUMINANCE_ALPHA */
    default: return 0;
    }
}

int mtgl_dev_draw_pixels(mtgl_dev *d, const mtgl_pixel_rect *rect, const void *pixels)
{
    if (!d || !rect || !pixels) return MTGL_E_INVALID;
    const uint32_t bpp = pixel_bpp(rect->format);
    if (bpp == 0 || rect->width <= 0 || rect->height <= 0) return MTGL_OK;       /* gl_api.c:1336-1338: unknown formats draw nothing */
    CU(cudaSetDevice(d->device));
    if (int prc_ = resolve_pending(d)) return prc_;
    if (int orc = order_after_transfers(d)) return orc;
    if (int prc = order_after_present(d)) return prc;
    /* Only the part of the rectangle that lands on this device's rows and inside the framebuffer's columns is staged and
     * launched (the reference skips the other pixels one by one, gl_api.c:1304-1312): rectangle row r lands on
     * framebuffer row height - 1 - (y + r), column c on x + c. */
    const int64_t fy_lo = std::max<int64_t>(d->band_y0, 0), fy_hi = std::min<int64_t>(d->band_y1, d->height);     /* [lo, hi) */
    const int64_t r_lo = std::max<int64_t>(0, (int64_t)d->height - 1 - rect->y - (fy_hi - 1));
    const int64_t r_hi = std::min<int64_t>(rect->height, (int64_t)d->height - 1 - rect->y - fy_lo + 1);
    const int64_t c_lo = std::max<int64_t>(0, -(int64_t)rect->x), c_hi = std::min<int64_t>(rect->width, (int64_t)d->width - rect->x);
    if (r_lo >= r_hi || c_lo >= c_hi) return MTGL_OK;
    mtgl_pixel_rect part = *rect;
    part.x = (int32_t)(rect->x + c_lo); part.y = (int32_t)(rect->y + r_lo);
    part.width = (int32_t)(c_hi - c_lo); part.height = (int32_t)(r_hi - r_lo);
    const size_t row_bytes = (size_t)part.width * bpp, bytes = row_bytes * (size_t)part.height;
    /* the staging buffer is reused in stream order (copy, kernel, next copy); growing it waits for the stream */
    int rc = reserve(d, d->pixel_stage, bytes);
    if (rc != MTGL_OK) return rc;
    /* from pageable memory this returns once the driver has taken its copy: the caller may reuse 'pixels' */
    const uint8_t *src = (const uint8_t *)pixels + ((size_t)r_lo * (size_t)rect->width + (size_t)c_lo) * bpp;
    CU(cudaMemcpy2DAsync(d->pixel_stage.ptr, row_bytes, src, (size_t)rect->width * bpp, row_bytes, (size_t)part.height, cudaMemcpyHostToDevice, d->stream));
    launch_draw_pixels(part, (const uint8_t *)d->pixel_stage.ptr, frame_targets(d), d->unorm8, d->stream);
    CU(cudaGetLastError());
    return MTGL_OK;
}

int mtgl_dev_read_pixels(mtgl_dev *d, int32_t x, int32_t y, int32_t width, int32_t height, uint32_t format, void *out)
{
    if (!d || !out) return MTGL_E_INVALID;
    const uint32_t bpp = (format == 0x1908) ? 4u : (format == 0x1907 ? 3u : 0u);
    if (bpp == 0 || width <= 0 || height <= 0) return MTGL_OK;                   /* other formats leave 'out' untouched */
    CU(cudaSetDevice(d->device));
    if (int prc_ = resolve_pending(d)) return prc_;
    if (int prc = order_after_present(d)) return prc;
    /* rows outside the framebuffer read as zeros, columns outside it as (0, 0, 0, 255) (gl_api.c:1193-1214): filled on
     * the host; only the visible part of the rectangle is gathered on the device and crosses PCIe */
    const int64_t r_lo = std::max<int64_t>(0, -(int64_t)y), r_hi = std::min<int64_t>(height, (int64_t)d->height - y);
    const int64_t c_lo = std::max<int64_t>(0, -(int64_t)x), c_hi = std::min<int64_t>(width, (int64_t)d->width - x);
    const bool any = r_lo < r_hi && c_lo < c_hi;
    uint8_t *o = (uint8_t *)out;
    const size_t out_row = (size_t)width * bpp;
    for (int64_t r = 0; r < height; r++) {
        uint8_t *row = o + (size_t)r * out_row;
        if (r < r_lo || r >= r_hi) { std::memset(row, 0, out_row); continue; }
        if (!any || c_lo > 0 || c_hi < width) {
            std::memset(row, 0, out_row);
            if (bpp == 4) for (int64_t c = 0; c < width; c++) if (c < c_lo || c >= c_hi) row[(size_t)c * 4 + 3] = 255;
        }
    }
    if (!any) { CU(cudaStreamSynchronize(d->stream)); return MTGL_OK; }
    const int32_t pw = (int32_t)(c_hi - c_lo), ph = (int32_t)(r_hi - r_lo);
    const size_t row_bytes = (size_t)pw * bpp, bytes = row_bytes * (size_t)ph;
    int rc = reserve(d, d->pixel_stage, bytes);
    if (rc != MTGL_OK) return rc;
    launch_read_pixels(frame_targets(d), (int32_t)(x + c_lo), (int32_t)(y + r_lo), pw, ph, bpp, (uint8_t *)d->pixel_stage.ptr, d->stream);
    CU(cudaMemcpy2DAsync(o + ((size_t)r_lo * (size_t)width + (size_t)c_lo) * bpp, out_row, d->pixel_stage.ptr, row_bytes, row_bytes, (size_t)ph,
                         cudaMemcpyDeviceToHost, d->stream));
    CU(cudaStreamSynchronize(d->stream));
    return MTGL_OK;
}

int mtgl_dev_frame_barrier(mtgl_dev *d, uint32_t participants)
{
    if (!d || participants == 0) return MTGL_E_INVALID;
    if (participants == 1) return MTGL_OK;
    CU(cudaSetDevice(d->device));
    if (int prc_ = resolve_pending(d)) return prc_;
    /* the counter lives behind the presenting GPU's colour plane: local for the presenter, NVLink-mapped for the others */
    uint32_t *plane = d->present ? d->present : d->color;
    unsigned long long *ctr = reinterpret_cast<unsigned long long *>((uint8_t *)plane + barrier_offset((size_t)d->width * d->height));
    d->barrier_epoch++;
    if (d->present_mode == MTGL_PRESENT_COPY) {
        /* behind this frame's raster kernels, on the side stream: push the band (peer-to-peer copy over NVLink), then the
         * barrier.  The main stream goes on with the next frame's geometry and visibility; its colour writers wait. */
        if (int orc = order_after_present(d)) return orc;       /* (two barriers without a frame in between) */
        CU(cudaEventRecord(d->raster_ev, d->stream));
        CU(cudaStreamWaitEvent(d->present_stream, d->raster_ev, 0));
        if (d->present && d->band_y1 > d->band_y0) {
            const size_t o = (size_t)d->band_y0 * d->width, n = (size_t)(d->band_y1 - d->band_y0) * d->width;
            CU(cudaMemcpyAsync(d->present + o, d->color + o, n * 4, cudaMemcpyDeviceToDevice, d->present_stream));
        }
        launch_frame_barrier(ctr, d->barrier_epoch * participants, d->barrier_timeout_ns, d->h_barrier_timed_out, d->present_stream);
        CU(cudaEventRecord(d->present_ev, d->present_stream));
        d->present_busy = true;
    } else launch_frame_barrier(ctr, d->barrier_epoch * participants, d->barrier_timeout_ns, d->h_barrier_timed_out, d->stream);
    CU(cudaGetLastError());
    return MTGL_OK;
}

int mtgl_dev_get_stats(mtgl_dev *d, mtgl_dev_stats *out)
{
    if (!d || !out) return MTGL_E_INVALID;
    CU(cudaSetDevice(d->device));
    if (int prc_ = resolve_pending(d)) return prc_;
    for (int i = 0; i < mtgl_dev::kEvSets; i++) {           /* oldest first: the last one folded is the last batch */
        mtgl_dev::EvSet &es = d->evset[(d->ev_next + i) % mtgl_dev::kEvSets];
        if (es.pending) { int rc = fold_timing(d, es); if (rc != MTGL_OK) return rc; }
    }
    d->stats.kernel_launches = kernel_launch_count();
    *out = d->stats;
    return MTGL_OK;
}

void *mtgl_dev_stream(mtgl_dev *d) { return d ? (void *)d->stream : nullptr; }

int mtgl_dev_timer_mark(mtgl_dev *d, int which)
{
    if (!d || which < 0 || which > 1) return MTGL_E_INVALID;
    CU(cudaSetDevice(d->device));
    if (int prc_ = resolve_pending(d)) return prc_;
    if (int prc = order_after_present(d)) return prc;           /* a band push in flight belongs to the time before the mark */
    CU(cudaEventRecord(d->mark_ev[which], d->stream));
    return MTGL_OK;
}

int mtgl_dev_timer_elapsed_ms(mtgl_dev *d, float *ms)
{
    if (!d || !ms) return MTGL_E_INVALID;
    CU(cudaSetDevice(d->device));
    CU(cudaEventSynchronize(d->mark_ev[1]));
    CU(cudaEventElapsedTime(ms, d->mark_ev[0], d->mark_ev[1]));
    return MTGL_OK;
}

const char *mtgl_dev_last_error(mtgl_dev *d) { return d ? d->err : "no device"; }

} // extern "C"
