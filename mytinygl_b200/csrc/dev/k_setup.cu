/*
 * k_setup.cu -- K2: primitive assembly, frustum clipping, perspective divide, integer snapping,
 * culling and triangle set-up, one thread per assembled triangle.
 *
 * Replaces flush_triangles/quads/triangle_strip/triangle_fan/polygon/quad_strip
 * (src/raster.c:961-1017, 1199-1231), render_triangle (901-958), clip_triangle and friends
 * (src/clipping.h:27-126), perspective_divide (raster.c:729-746), ndc_to_screen (59-63),
 * should_cull (751-774) and the set-up half of rasterize_triangle_smooth (458-529).
 *
 * Work decomposition: a CTA owns a chunk of 256 consecutive input triangles.  Every thread first
 * counts how many sub-triangles of its triangle survive (0 or 1 without clipping, up to 7 with),
 * a block scan turns the counts into dense, submission-ordered slots inside the chunk, one
 * atomicAdd per CTA reserves the chunk's slots in the record array, and a second evaluation
 * writes the 160-byte records.  A record's id = chunk << 11 | slot-in-chunk is therefore ordered
 * exactly like the reference's sequential loop; the tile kernel sorts by it.
 *
 * Algorithmic bytes per input triangle: 3 x 48 B gathered (+ 3 x 32 B for per-fragment lighting),
 * 160 B written per surviving sub-triangle.
 */
#include "dev_common.cuh"

namespace mtgl_dev_impl {

void note_launch();

struct SVert {              /* vertex_t (graphics.h:438-446) minus the unused object normal */
    float x, y, z, w;
    float r, g, b, a;
    float u, v, ez;
    float epx, epy, epz, enx, eny, enz;
};

__device__ __forceinline__ SVert load_vertex(const BatchDev &b, uint32_t i)
{
    SVert o;
    float4 p = b.v_clip[i], c = b.v_color[i], t = b.v_tex[i];
    o.x = p.x; o.y = p.y; o.z = p.z; o.w = p.w;
    o.r = c.x; o.g = c.y; o.b = c.z; o.a = c.w;
    o.u = t.x; o.v = t.y; o.ez = t.z;
    if (b.need_eye) {
        float4 e = b.v_epos[i], n = b.v_enrm[i];
        o.epx = e.x; o.epy = e.y; o.epz = e.z; o.enx = n.x; o.eny = n.y; o.enz = n.z;
    } else { o.epx = o.epy = o.epz = 0.0f; o.enx = o.eny = 0.0f; o.enz = 1.0f; }
    return o;
}

__device__ __forceinline__ float plane_dist(const SVert &v, int plane)   /* clipping.h:27-32: near far left right bottom top */
{
    switch (plane) {
    case 0: return v.z + v.w;
    case 1: return v.w - v.z;
    case 2: return v.x + v.w;
    case 3: return v.w - v.x;
    case 4: return v.y + v.w;
    default: return v.w - v.y;
    }
}

__device__ __forceinline__ bool inside_all(const SVert &v)
{
    return (v.z + v.w) >= 0 && (v.w - v.z) >= 0 && (v.x + v.w) >= 0 && (v.w - v.x) >= 0 && (v.y + v.w) >= 0 && (v.w - v.y) >= 0;
}

__device__ __forceinline__ float lerp1(float a, float b, float t) { return a + t * (b - a); }   /* graphics.h:369-371 */

__device__ SVert vertex_lerp(const SVert &a, const SVert &b, float t)   /* graphics.h:465-475 */
{
    SVert o;
    o.x = lerp1(a.x, b.x, t); o.y = lerp1(a.y, b.y, t); o.z = lerp1(a.z, b.z, t); o.w = lerp1(a.w, b.w, t);
    float s = 1.0f - t;                                     /* colour uses a*(1-t) + b*t (graphics.h:293-295) */
    o.r = a.r * s + b.r * t; o.g = a.g * s + b.g * t; o.b = a.b * s + b.b * t; o.a = a.a * s + b.a * t;
    o.u = lerp1(a.u, b.u, t); o.v = lerp1(a.v, b.v, t);
    o.ez = lerp1(a.ez, b.ez, t);
    o.epx = lerp1(a.epx, b.epx, t); o.epy = lerp1(a.epy, b.epy, t); o.epz = lerp1(a.epz, b.epz, t);
    o.enx = lerp1(a.enx, b.enx, t); o.eny = lerp1(a.eny, b.eny, t); o.enz = lerp1(a.enz, b.enz, t);
    return o;
}

__device__ __forceinline__ void snap(SVert &v, int plane)   /* clipping.h:37-47 */
{
    switch (plane) {
    case 0: v.z = -v.w; break;
    case 1: v.z = v.w; break;
    case 2: v.x = -v.w; break;
    case 3: v.x = v.w; break;
    case 4: v.y = -v.w; break;
    default: v.y = v.w; break;
    }
}

#define MAX_CLIP 12

__device__ int clip_plane(const SVert *in, int n, SVert *out, int plane)   /* clipping.h:50-99 */
{
    if (n == 0) return 0;
    int m = 0;
    int prev = n - 1;
    float pd = plane_dist(in[prev], plane);
    for (int i = 0; i < n; i++) {
        float cd = plane_dist(in[i], plane);
        if (pd >= 0) {
            if (cd >= 0) out[m++] = in[i];
            else {
                float den = pd - cd;
                if (fabsf(den) > 1e-10f) {
                    out[m] = vertex_lerp(in[prev], in[i], pd / den);
                    snap(out[m], plane);
                    m++;
                }
            }
        } else if (cd >= 0) {
            float den = pd - cd;
            if (fabsf(den) > 1e-10f) {
                out[m] = vertex_lerp(in[prev], in[i], pd / den);
                snap(out[m], plane);
                m++;
            }
            out[m++] = in[i];
        }
        prev = i;
        pd = cd;
    }
    return m;
}

__device__ __forceinline__ void persp_divide(SVert &v)   /* raster.c:729-746: w becomes 1/w */
{
    if (fabsf(v.w) < 1e-6f) { v.x = 0.0f; v.y = 0.0f; v.z = 0.0f; v.w = 1.0f; return; }
    float iw = 1.0f / v.w;
    v.x *= iw; v.y *= iw; v.z *= iw; v.w = iw;
}

__device__ __forceinline__ int imin3(int a, int b, int c) { return min(a, min(b, c)); }
__device__ __forceinline__ int imax3(int a, int b, int c) { return max(a, max(b, c)); }

/* One fan sub-triangle after the divide: snap, cull, set up.  Returns false when nothing is to be
 * rasterised; otherwise fills *rec (and *eye).  Mirrors raster.c:916-956 and 458-529. */
__device__ bool setup_subtri(const mtgl_state *st, const RasterCfg *cfg, const FrameTargets &fb, const SVert &a,
                             const SVert &b, const SVert &c, uint32_t state_index, TriRecord *rec, TriEye *eye)
{
    const float vw = (float)st->viewport[2], vh = (float)st->viewport[3];
    const float vx = (float)st->viewport[0], vy = (float)st->viewport[1];
    int32_t x0 = f2i_x86((a.x + 1.0f) * 0.5f * vw + vx), y0 = f2i_x86((1.0f - a.y) * 0.5f * vh + vy);
    int32_t x1 = f2i_x86((b.x + 1.0f) * 0.5f * vw + vx), y1 = f2i_x86((1.0f - b.y) * 0.5f * vh + vy);
    int32_t x2 = f2i_x86((c.x + 1.0f) * 0.5f * vw + vx), y2 = f2i_x86((1.0f - c.y) * 0.5f * vh + vy);

    /* signed area of the snapped triangle decides culling and facing (raster.c:923-935) */
    float sa = (float)(x1 - x0) * (float)(y2 - y0) - (float)(x2 - x0) * (float)(y1 - y0);
    if (st->caps & MTGL_CAP_CULL_FACE) {
        bool front = (st->front_face == G_CCW) ? (sa < 0) : (sa > 0);
        bool cull = (st->cull_face_mode == G_FRONT) ? front : (st->cull_face_mode == G_BACK) ? !front : true;
        if (cull) return false;
    }
    bool back = (st->front_face == G_CCW) ? (sa >= 0) : (sa < 0);
    uint32_t pm = back ? st->polygon_mode_back : st->polygon_mode_front;
    if (pm != G_FILL) return false;     /* TODO(next): GL_LINE / GL_POINT polygon modes */

    int32_t minX = imin3(x0, x1, x2), minY = imin3(y0, y1, y2), maxX = imax3(x0, x1, x2), maxY = imax3(y0, y1, y2);
    const int32_t *vp = st->viewport, *sc = st->scissor;
    if (minX < vp[0]) minX = vp[0];
    if (minY < vp[1]) minY = vp[1];
    if (maxX >= vp[0] + vp[2]) maxX = vp[0] + vp[2] - 1;
    if (maxY >= vp[1] + vp[3]) maxY = vp[1] + vp[3] - 1;
    if (st->caps & MTGL_CAP_SCISSOR_TEST) {
        if (minX < sc[0]) minX = sc[0];
        if (minY < sc[1]) minY = sc[1];
        if (maxX >= sc[0] + sc[2]) maxX = sc[0] + sc[2] - 1;
        if (maxY >= sc[1] + sc[3]) maxY = sc[1] + sc[3] - 1;
    }
    if (minX > maxX || minY > maxY) return false;

    /* edge_function(x0,y0,x1,y1,x2,y2) (raster.c:483, 299-302) */
    float area = ((float)x2 - (float)x0) * ((float)y1 - (float)y0) - ((float)y2 - (float)y0) * ((float)x1 - (float)x0);
    if (fabsf(area) < 0.5f) return false;

    /* pixels outside the framebuffer are dropped by the bounds-checked accessors (framebuffer.h:92-134);
     * rows outside this device's band belong to another GPU */
    if (minX < 0) minX = 0;
    if (maxX >= fb.width) maxX = fb.width - 1;
    if (minY < fb.band_y0) minY = fb.band_y0;
    if (maxY >= fb.band_y1) maxY = fb.band_y1 - 1;
    if (minX > maxX || minY > maxY) return false;
    if (!rec) return true;

    rec->x0 = x0; rec->y0 = y0; rec->x1 = x1; rec->y1 = y1; rec->x2 = x2; rec->y2 = y2;
    rec->state_flags = state_index | (back ? STATE_BACK_BIT : 0u) | ((cfg->flags & RC_DEFER) ? STATE_DEFER_BIT : 0u);
    rec->bbox_min = (uint32_t)minX | ((uint32_t)minY << 16);
    rec->bbox_max = (uint32_t)maxX | ((uint32_t)maxY << 16);
    rec->z0 = a.z; rec->z1 = b.z; rec->z2 = c.z;
    rec->w0 = a.w; rec->w1 = b.w; rec->w2 = c.w;
    rec->area = area;
    rec->inv_area = 1.0f / area;
    rec->c0[0] = a.r; rec->c0[1] = a.g; rec->c0[2] = a.b; rec->c0[3] = a.a;
    rec->c1[0] = b.r; rec->c1[1] = b.g; rec->c1[2] = b.b; rec->c1[3] = b.a;
    rec->c2[0] = c.r; rec->c2[1] = c.g; rec->c2[2] = c.b; rec->c2[3] = c.a;
    rec->u0 = a.u; rec->v0 = a.v; rec->u1 = b.u; rec->v1 = b.v; rec->u2 = c.u; rec->v2 = c.v;
    rec->ez0 = a.ez; rec->ez1 = b.ez; rec->ez2 = c.ez;

    float lod = 0.0f;                   /* one LOD per triangle from non-perspective UV deltas (raster.c:505-529) */
    if (cfg->flags & RC_TEXTURED) {
        float screen_area = fabsf(area) * 0.5f;
        float tw = (float)cfg->tex_w, th = (float)cfg->tex_h;
        float du1 = (b.u - a.u) * tw, dv1 = (b.v - a.v) * th;
        float du2 = (c.u - a.u) * tw, dv2 = (c.v - a.v) * th;
        float texel_area = fabsf(du1 * dv2 - du2 * dv1) * 0.5f;
        if (screen_area > 0.0f) {
            float tpp = texel_area / screen_area;
            if (tpp > 0.0f) {
                lod = log2f(tpp) * 0.5f;
                if (lod < 0.0f) lod = 0.0f;
            }
        }
    }
    rec->lod = lod;
    if (eye) {
        eye->ep0[0] = a.epx; eye->ep0[1] = a.epy; eye->ep0[2] = a.epz; eye->ep0[3] = 0.0f;
        eye->ep1[0] = b.epx; eye->ep1[1] = b.epy; eye->ep1[2] = b.epz; eye->ep1[3] = 0.0f;
        eye->ep2[0] = c.epx; eye->ep2[1] = c.epy; eye->ep2[2] = c.epz; eye->ep2[3] = 0.0f;
        eye->en0[0] = a.enx; eye->en0[1] = a.eny; eye->en0[2] = a.enz; eye->en0[3] = 0.0f;
        eye->en1[0] = b.enx; eye->en1[1] = b.eny; eye->en1[2] = b.enz; eye->en1[3] = 0.0f;
        eye->en2[0] = c.enx; eye->en2[1] = c.eny; eye->en2[2] = c.enz; eye->en2[3] = 0.0f;
    }
    return true;
}

__device__ __forceinline__ void store_record(TriRecord *dst, const TriRecord &r)
{
    const int4 *s = reinterpret_cast<const int4 *>(&r);
    int4 *d = reinterpret_cast<int4 *>(dst);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(TriRecord) / 16); k++) d[k] = s[k];
}

__device__ __forceinline__ uint32_t find_draw_tri(const uint32_t *base, uint32_t n, uint32_t g)
{
    uint32_t lo = 0, hi = n;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (__ldg(base + mid) <= g) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(SETUP_THREADS) k_setup(BatchDev b, FrameTargets fb)
{
    __shared__ uint32_t warp_sums[SETUP_THREADS / 32];
    __shared__ uint32_t chunk_slot0;

    const uint32_t chunk = blockIdx.x;
    const uint32_t t = chunk * SETUP_THREADS + threadIdx.x;
    const bool valid = t < b.n_triangles;

    SVert v0, v1, v2;
    SVert poly_a[MAX_CLIP], poly_b[MAX_CLIP];
    SVert *poly = nullptr;
    int npoly = 0;
    bool clipped = false;
    const mtgl_state *st = nullptr;
    const RasterCfg *cfg = nullptr;
    uint32_t state_index = 0;
    uint32_t count = 0;

    if (valid) {
        uint32_t d = (b.n_draws == 1) ? 0u : find_draw_tri(b.draw_tbase, b.n_draws, t);
        const DevDraw &dr = b.draws[d];
        uint32_t k = t - dr.tbase, i0, i1, i2;
        switch (dr.mode) {                                  /* raster.c:961-1017, 1199-1231 */
        case G_TRIANGLES: i0 = 3 * k; i1 = i0 + 1; i2 = i0 + 2; break;
        case G_QUADS: { uint32_t q = 4 * (k >> 1); i0 = q; if (k & 1) { i1 = q + 2; i2 = q + 3; } else { i1 = q + 1; i2 = q + 2; } break; }
        case G_TRIANGLE_STRIP: if (k & 1) { i0 = k + 1; i1 = k; } else { i0 = k; i1 = k + 1; } i2 = k + 2; break;
        case G_QUAD_STRIP: { uint32_t q = 2 * (k >> 1); i0 = q; if (k & 1) { i1 = q + 3; i2 = q + 2; } else { i1 = q + 1; i2 = q + 3; } break; }
        default: i0 = 0; i1 = k + 1; i2 = k + 2; break;     /* fan, polygon */
        }
        state_index = dr.raster_state;
        st = b.states + state_index;
        cfg = b.cfgs + state_index;
        v0 = load_vertex(b, dr.vbase + i0);
        v1 = load_vertex(b, dr.vbase + i1);
        v2 = load_vertex(b, dr.vbase + i2);

        if (inside_all(v0) && inside_all(v1) && inside_all(v2)) {
            /* Sutherland-Hodgman returns its input unchanged when every vertex passes every plane */
            persp_divide(v0); persp_divide(v1); persp_divide(v2);
            count = setup_subtri(st, cfg, fb, v0, v1, v2, state_index, nullptr, nullptr) ? 1u : 0u;
        } else {
            clipped = true;
            poly_a[0] = v0; poly_a[1] = v1; poly_a[2] = v2;
            int n = clip_plane(poly_a, 3, poly_b, 0);       /* clipping.h:106-126 */
            if (n) n = clip_plane(poly_b, n, poly_a, 1);
            if (n) n = clip_plane(poly_a, n, poly_b, 2);
            if (n) n = clip_plane(poly_b, n, poly_a, 3);
            if (n) n = clip_plane(poly_a, n, poly_b, 4);
            if (n) n = clip_plane(poly_b, n, poly_a, 5);
            poly = poly_a;
            npoly = (n >= 3) ? n : 0;
            for (int j = 0; j < npoly; j++) persp_divide(poly[j]);
            for (int j = 1; j + 1 < npoly; j++)
                if (setup_subtri(st, cfg, fb, poly[0], poly[j], poly[j + 1], state_index, nullptr, nullptr)) count++;
        }
    }

    /* block-wide exclusive scan of the survivor counts -> submission-ordered slots */
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = count;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t n = __shfl_up_sync(0xFFFFFFFFu, incl, o);
        if (lane >= (uint32_t)o) incl += n;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = (lane < SETUP_THREADS / 32) ? warp_sums[lane] : 0u;
        uint32_t wi = w;
#pragma unroll
        for (int o = 1; o < SETUP_THREADS / 32; o <<= 1) {
            uint32_t n = __shfl_up_sync(0xFFFFFFFFu, wi, o);
            if (lane >= (uint32_t)o) wi += n;
        }
        if (lane < SETUP_THREADS / 32) warp_sums[lane] = wi - w;
        if (lane == SETUP_THREADS / 32 - 1) {
            uint32_t total = wi;
            uint32_t base = total ? atomicAdd(&b.counters->records, total) : 0u;
            if (base + total > b.record_capacity) { atomicExch(&b.counters->overflow, 1u); base = 0xFFFFFFFFu; }
            chunk_slot0 = base;
            b.chunk_base[chunk] = base;
        }
    }
    __syncthreads();
    if (count == 0 || chunk_slot0 == 0xFFFFFFFFu) return;
    uint32_t slot = warp_sums[warp] + incl - count;         /* index inside the chunk */

    TriRecord rec;
    TriEye eye;
    TriEye *eyep = b.need_eye ? &eye : nullptr;
    if (!clipped) {
        setup_subtri(st, cfg, fb, v0, v1, v2, state_index, &rec, eyep);
        rec.id = (chunk << CHUNK_SHIFT) | slot;
        store_record(b.records + chunk_slot0 + slot, rec);
        if (eyep) b.rec_eye[chunk_slot0 + slot] = eye;
    } else {
        for (int j = 1; j + 1 < npoly; j++) {
            if (!setup_subtri(st, cfg, fb, poly[0], poly[j], poly[j + 1], state_index, &rec, eyep)) continue;
            rec.id = (chunk << CHUNK_SHIFT) | slot;
            store_record(b.records + chunk_slot0 + slot, rec);
            if (eyep) b.rec_eye[chunk_slot0 + slot] = eye;
            slot++;
        }
    }
}

void launch_setup(const BatchDev &b, const FrameTargets &fb, cudaStream_t s)
{
    if (b.n_triangles == 0) return;
    uint32_t chunks = (b.n_triangles + SETUP_THREADS - 1) / SETUP_THREADS;
    k_setup<<<chunks, SETUP_THREADS, 0, s>>>(b, fb);
    note_launch();
}

} // namespace mtgl_dev_impl
