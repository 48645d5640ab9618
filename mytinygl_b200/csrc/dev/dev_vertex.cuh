/*
 * dev_vertex.cuh -- the vertex stage as device functions: attribute fetch, modelview / projection / texture
 * transforms, per-vertex lighting.  Used by k_vertex (one thread per vertex, results to the SoA vertex streams) and
 * by k_setup's fused path for independent triangles (vertices are shaded inside the set-up thread of the one
 * triangle that uses them, only if the triangle survives culling -- no 48 B/vertex round trip through HBM).
 *
 * Replaces emit_vertex (src/gl_api.c:263-348), compute_lighting (src/lighting.h:53-142) and the attribute fetch
 * loop of glDrawArrays / glDrawElements (src/gl_api.c:1745-1941).
 */
#ifndef MTGL_DEV_VERTEX_CUH
#define MTGL_DEV_VERTEX_CUH

#include "dev_common.cuh"

namespace mtgl_dev_impl {

/* get_array_element (gl_api.c:1758-1797) against a buffer mirror; out-of-range reads (undefined
 * behaviour in the reference) return the defaults instead of faulting */
__device__ __forceinline__ void fetch_attrib(const DevAttrib &a, int32_t index, float *out, int want)
{
    bool ok = a.ptr != nullptr && index >= 0;
    uint64_t off = 0;
    if (ok) {
        off = (uint64_t)(uint32_t)index * a.stride;
        uint32_t bytes = a.size * (a.type == MTGL_TYPE_F32 ? 4u : 1u);
        ok = off + bytes <= a.avail;
    }
    if (!ok) {
        for (int i = 0; i < want; i++) out[i] = (i < 3) ? 0.0f : 1.0f;
        return;
    }
    const uint8_t *p = a.ptr + off;
    if (a.type == MTGL_TYPE_F32 && (((uintptr_t)p) & 3u) == 0) {       /* the common case: aligned floats, no per-component branching */
        const float *f = reinterpret_cast<const float *>(p);
        const int n = (int)a.size;
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (i < want) out[i] = (i < n) ? __ldg(f + i) : ((i == 3) ? 1.0f : 0.0f);
        return;
    }
    for (int i = 0; i < want; i++) {
        if (i < (int)a.size) {
            if (a.type == MTGL_TYPE_F32) {
                uint32_t w = (uint32_t)p[4 * i] | ((uint32_t)p[4 * i + 1] << 8) | ((uint32_t)p[4 * i + 2] << 16) |
                             ((uint32_t)p[4 * i + 3] << 24);
                out[i] = __uint_as_float(w);
            } else out[i] = (float)p[i] / 255.0f;
        } else out[i] = (i == 3) ? 1.0f : 0.0f;
    }
}

/* what a vertex is made from: raw attributes + the state block current at glVertex* */
struct VertexIn {
    float px, py, pz, nx, ny, nz, s, t;
    Color4 cur;
    const mtgl_state *st;
};

/* element index of vertex i of an array draw (glDrawElements index fetch, gl_api.c:1880-1941) */
__device__ __forceinline__ int32_t element_index(const DevDraw &dr, uint32_t i)
{
    if (!dr.index_type) return dr.first + (int32_t)i;
    uint32_t u = 0;
    if (dr.index_type == G_UNSIGNED_SHORT) {
        if (2ull * i + 2 <= dr.index_avail) u = (uint32_t)dr.index_ptr[2 * i] | ((uint32_t)dr.index_ptr[2 * i + 1] << 8);
    } else if (dr.index_type == G_UNSIGNED_INT) {
        if (4ull * i + 4 <= dr.index_avail)
            u = (uint32_t)dr.index_ptr[4 * i] | ((uint32_t)dr.index_ptr[4 * i + 1] << 8) |
                ((uint32_t)dr.index_ptr[4 * i + 2] << 16) | ((uint32_t)dr.index_ptr[4 * i + 3] << 24);
    } else if ((uint64_t)i < dr.index_avail) u = dr.index_ptr[i];
    return (int32_t)u;
}

/* position + state of vertex i of a draw: all the clip / cull decision needs */
__device__ __forceinline__ void fetch_position(const mtgl_in_vertex *staged, const mtgl_state *states, const DevDraw &dr, uint32_t i,
                                               float &px, float &py, float &pz, const mtgl_state *&st)
{
    if (dr.source == MTGL_SRC_STAGED) {
        const mtgl_in_vertex *iv = staged + dr.first_staged + i;
        px = iv->position[0]; py = iv->position[1]; pz = iv->position[2];
        st = states + min(iv->state, dr.state_max);
    } else {
        float p[4];
        fetch_attrib(dr.position, element_index(dr, i), p, 4);
        px = p[0]; py = p[1]; pz = (dr.position.size == 2) ? 0.0f : p[2];
        st = states + dr.vertex_state;
    }
}

/* texture coordinate of vertex i (before the texture matrix) */
__device__ __forceinline__ void fetch_texcoord(const mtgl_in_vertex *staged, const DevDraw &dr, uint32_t i, float &s, float &t)
{
    if (dr.source == MTGL_SRC_STAGED) {
        const mtgl_in_vertex *iv = staged + dr.first_staged + i;
        s = iv->texcoord[0]; t = iv->texcoord[1];
    } else {
        s = dr.cur_texcoord[0]; t = dr.cur_texcoord[1];
        if (dr.texcoord.enabled) {
            float tc[2];
            fetch_attrib(dr.texcoord, element_index(dr, i), tc, 2);
            s = tc[0]; t = tc[1];
        }
    }
}

/* texture matrix, divide by q only when q is neither 0 nor 1 (gl_api.c:327-335) */
__device__ __forceinline__ void tex_transform(const mtgl_state *st, float s, float t, float &tu, float &tv)
{
    const float *tm = st->texture;
    tu = tm[0] * s + tm[4] * t + tm[8] * 0.0f + tm[12] * 1.0f;
    tv = tm[1] * s + tm[5] * t + tm[9] * 0.0f + tm[13] * 1.0f;
    float tq = tm[3] * s + tm[7] * t + tm[11] * 0.0f + tm[15] * 1.0f;
    if (tq != 0.0f && tq != 1.0f) { tu = tu / tq; tv = tv / tq; }
}

/* slot of element index e of a shared-vertex draw (DevDraw::shared_verts): the element itself, or the default slot */
__device__ __forceinline__ uint32_t shared_slot(const DevDraw &dr, int32_t e)
{
    return (e >= 0 && (uint32_t)e < dr.shared_verts) ? (uint32_t)e : dr.shared_verts;
}

/* i: position in the draw's vertex sequence -- or, with by_element, the buffer element itself (negative: the defaults) */
__device__ __forceinline__ void fetch_vertex(const mtgl_in_vertex *staged, const mtgl_state *states, const DevDraw &dr, uint32_t i, VertexIn &v,
                                             bool by_element = false)
{
    if (dr.source == MTGL_SRC_STAGED) {
        const mtgl_in_vertex *iv = staged + dr.first_staged + i;
        v.px = iv->position[0]; v.py = iv->position[1]; v.pz = iv->position[2];
        v.cur = { iv->color[0], iv->color[1], iv->color[2], iv->color[3] };
        v.s = iv->texcoord[0]; v.t = iv->texcoord[1];
        v.nx = iv->normal[0]; v.ny = iv->normal[1]; v.nz = iv->normal[2];
        v.st = states + min(iv->state, dr.state_max);
    } else {
        const int32_t idx = by_element ? (int32_t)i : element_index(dr, i);
        float p[4], c[4], tc[2], n[3];
        fetch_attrib(dr.position, idx, p, 4);
        v.px = p[0]; v.py = p[1]; v.pz = (dr.position.size == 2) ? 0.0f : p[2];
        v.cur = { dr.cur_color[0], dr.cur_color[1], dr.cur_color[2], dr.cur_color[3] };
        if (dr.color.enabled) {            /* glColor4f sanitising, gl_api.c:708-722 */
            fetch_attrib(dr.color, idx, c, 4);
            float r = c[0], gg = c[1], bb = c[2], a = c[3];
            if (isnan(r) || isinf(r)) r = 0.0f;
            if (isnan(gg) || isinf(gg)) gg = 0.0f;
            if (isnan(bb) || isinf(bb)) bb = 0.0f;
            if (isnan(a) || isinf(a)) a = 1.0f;
            v.cur = { sat01(r), sat01(gg), sat01(bb), sat01(a) };
        }
        v.s = dr.cur_texcoord[0]; v.t = dr.cur_texcoord[1];
        if (dr.texcoord.enabled) { fetch_attrib(dr.texcoord, idx, tc, 2); v.s = tc[0]; v.t = tc[1]; }
        v.nx = dr.cur_normal[0]; v.ny = dr.cur_normal[1]; v.nz = dr.cur_normal[2];
        if (dr.normal.enabled) { fetch_attrib(dr.normal, idx, n, 3); v.nx = n[0]; v.ny = n[1]; v.nz = n[2]; }
        v.st = states + dr.vertex_state;
    }
}

__host__ __device__ __forceinline__ bool attrib_range_ok(const DevAttrib &a, int32_t first, uint32_t count)
{
    if (!a.ptr || a.type != MTGL_TYPE_F32 || (((uintptr_t)a.ptr) & 3u) || (a.stride & 3u) || first < 0 || count == 0) return false;
    const uint64_t last = (uint64_t)((uint32_t)first + count - 1u) * a.stride + (uint64_t)a.size * 4u;
    return last <= a.avail;
}

/* the descriptor of draw d (filled on the host once per draw and pass, DevDraw::fast); valid = 0 when the draw does not qualify */
__host__ __device__ __forceinline__ void fast_draw_init(FastDraw &f, const mtgl_state *states, const DevDraw &dr, uint32_t d, uint32_t tri_begin, uint32_t tri_end)
{
    f.valid = 0;
    f.draw = d; f.tri_begin = tri_begin; f.tri_end = tri_end; f.tbase = dr.tbase;
    if (dr.source != MTGL_SRC_ARRAYS || dr.index_type || dr.color.enabled || !dr.fused) return;
    if (!attrib_range_ok(dr.position, dr.first, dr.count) || dr.position.size < 2) return;
    if (dr.normal.enabled && (!attrib_range_ok(dr.normal, dr.first, dr.count) || dr.normal.size != 3)) return;
    if (dr.texcoord.enabled && (!attrib_range_ok(dr.texcoord, dr.first, dr.count) || dr.texcoord.size < 2)) return;
    f.pos = dr.position.ptr; f.pos_stride = dr.position.stride; f.pos_size = dr.position.size;
    f.nrm = dr.normal.enabled ? dr.normal.ptr : nullptr; f.nrm_stride = dr.normal.stride;
    f.tex = dr.texcoord.enabled ? dr.texcoord.ptr : nullptr; f.tex_stride = dr.texcoord.stride;
    f.first = dr.first;
#pragma unroll
    for (int k = 0; k < 4; k++) f.cur_color[k] = dr.cur_color[k];
#pragma unroll
    for (int k = 0; k < 3; k++) f.cur_normal[k] = dr.cur_normal[k];
    f.cur_texcoord[0] = dr.cur_texcoord[0]; f.cur_texcoord[1] = dr.cur_texcoord[1];
    f.st = states + dr.vertex_state;
    f.valid = 1;
}

__device__ __forceinline__ void fast_position(const FastDraw &f, uint32_t i, float &px, float &py, float &pz)
{
    const float *p = reinterpret_cast<const float *>(f.pos + (uint64_t)((uint32_t)f.first + i) * f.pos_stride);
    px = __ldg(p); py = __ldg(p + 1);
    pz = (f.pos_size >= 3) ? __ldg(p + 2) : 0.0f;
}

__device__ __forceinline__ void fast_texcoord(const FastDraw &f, uint32_t i, float &s, float &t)
{
    s = f.cur_texcoord[0]; t = f.cur_texcoord[1];
    if (f.tex) {
        const float *p = reinterpret_cast<const float *>(f.tex + (uint64_t)((uint32_t)f.first + i) * f.tex_stride);
        s = __ldg(p); t = __ldg(p + 1);
    }
}

__device__ __forceinline__ void fast_vertex(const FastDraw &f, uint32_t i, VertexIn &v)
{
    fast_position(f, i, v.px, v.py, v.pz);
    fast_texcoord(f, i, v.s, v.t);
    v.nx = f.cur_normal[0]; v.ny = f.cur_normal[1]; v.nz = f.cur_normal[2];
    if (f.nrm) {
        const float *p = reinterpret_cast<const float *>(f.nrm + (uint64_t)((uint32_t)f.first + i) * f.nrm_stride);
        v.nx = __ldg(p); v.ny = __ldg(p + 1); v.nz = __ldg(p + 2);
    }
    v.cur = { f.cur_color[0], f.cur_color[1], f.cur_color[2], f.cur_color[3] };
    v.st = f.st;
}

/* eye = MV * (x, y, z, 1)   (graphics.h:131-138: ((m0*x + m4*y) + m8*z) + m12*w), clip = P * eye (raster.c:48-56) */
__device__ __forceinline__ void to_eye(const mtgl_state *st, float px, float py, float pz, float &ex, float &ey, float &ez, float &ew)
{
    const float *mv = st->modelview;
    ex = mv[0] * px + mv[4] * py + mv[8] * pz + mv[12] * 1.0f;
    ey = mv[1] * px + mv[5] * py + mv[9] * pz + mv[13] * 1.0f;
    ez = mv[2] * px + mv[6] * py + mv[10] * pz + mv[14] * 1.0f;
    ew = mv[3] * px + mv[7] * py + mv[11] * pz + mv[15] * 1.0f;
}
__device__ __forceinline__ float4 to_clip(const mtgl_state *st, float ex, float ey, float ez, float ew)
{
    const float *pr = st->projection;
    return make_float4(pr[0] * ex + pr[4] * ey + pr[8] * ez + pr[12] * ew, pr[1] * ex + pr[5] * ey + pr[9] * ez + pr[13] * ew,
                       pr[2] * ex + pr[6] * ey + pr[10] * ez + pr[14] * ew, pr[3] * ex + pr[7] * ey + pr[11] * ez + pr[15] * ew);
}

/* the post-transform vertex: what k_vertex stores as five float4 streams */
struct VertexOut {
    float4 clip, color, tex;        /* tex = (u, v, eye_z, 0) */
    float4 epos, enrm;              /* eye-space position / unit normal (per-fragment lighting only) */
};

__device__ __forceinline__ void shade_vertex(const VertexIn &v, VertexOut &o)
{
    const mtgl_state *st = v.st;
    float ex, ey, ez, ew;
    to_eye(st, v.px, v.py, v.pz, ex, ey, ez, ew);

    /* eye normal = normalize(N * (n, 0)); the fourth column of N is zero (graphics.h:247-257) */
    const float *nm = st->normal;
    float enx = nm[0] * v.nx + nm[4] * v.ny + nm[8] * v.nz + 0.0f;
    float eny = nm[1] * v.nx + nm[5] * v.ny + nm[9] * v.nz + 0.0f;
    float enz = nm[2] * v.nx + nm[6] * v.ny + nm[10] * v.nz + 0.0f;
    normalize3(enx, eny, enz);

    Color4 vc = v.cur;
    if ((st->caps & MTGL_CAP_LIGHTING) && st->shade_model != G_PHONG) {
        MaterialRegs mat;
        load_material(mat, &st->material_front);
        if (st->caps & MTGL_CAP_COLOR_MATERIAL) {          /* gl_api.c:285-312, front material only matters here */
            uint32_t face = st->color_material_face, mode = st->color_material_mode;
            if (face == G_FRONT || face == G_FRONT_AND_BACK) {
                Color4 k = color_clamp(v.cur);
                float kv[4] = { k.r, k.g, k.b, k.a };
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    if (mode == G_AMBIENT || mode == G_AMBIENT_AND_DIFFUSE) mat.ambient[q] = kv[q];
                    if (mode == G_DIFFUSE || mode == G_AMBIENT_AND_DIFFUSE) mat.diffuse[q] = kv[q];
                    if (mode == G_SPECULAR) mat.specular[q] = kv[q];
                    if (mode == G_EMISSION) mat.emission[q] = kv[q];
                }
            }
        }
        vc = lighting_body(st, ex, ey, ez, enx, eny, enz, mat);
    }

    o.clip = to_clip(st, ex, ey, ez, ew);

    float tu, tv;
    tex_transform(st, v.s, v.t, tu, tv);

    o.color = make_float4(vc.r, vc.g, vc.b, vc.a);
    o.tex = make_float4(tu, tv, -ez, 0.0f);
    o.epos = make_float4(ex, ey, ez, 0.0f);
    o.enrm = make_float4(enx, eny, enz, 0.0f);
}

} // namespace mtgl_dev_impl

#endif
