/*
 * k_vertex.cu -- K1: the vertex stage, one thread per vertex.
 *
 * Replaces emit_vertex (src/gl_api.c:263-348) + the per-vertex compute_lighting call
 * (src/lighting.h:53-142) + the attribute fetch loop of glDrawArrays / glDrawElements
 * (src/gl_api.c:1745-1941).  4x4 matrix-vector products and an 8-light loop are not a dense
 * contraction: CUDA cores only, tensor cores are not used (north_star, subsystem 1).
 *
 * Data layout: input is either the staged immediate-mode stream (52 B AoS records, read once,
 * coalesced enough at 13 words/thread) or buffer-object mirrors in HBM addressed per attribute;
 * output is three (optionally five) float4 SoA streams so that the set-up kernel's gathers are
 * 128-bit transactions:
 *     v_clip  = clip-space position
 *     v_color = lit (or current) colour
 *     v_tex   = (u, v, eye_z, 0)
 *     v_epos / v_enrm = eye-space position / normal (only when a state of the batch needs
 *                       per-fragment lighting: GL_PHONG or two-sided)
 * Algorithmic bytes per vertex: enabled attribute bytes in (C4: 32 B) + 48 B out.
 */
#include "dev_common.cuh"

namespace mtgl_dev_impl {

static unsigned long long g_launches = 0;
uint64_t kernel_launch_count() { return g_launches; }
void note_launch() { g_launches++; }

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
    for (int i = 0; i < want; i++) {
        if (i < (int)a.size) {
            if (a.type == MTGL_TYPE_F32) {
                if ((((uintptr_t)p) & 3u) == 0) out[i] = __ldg((const float *)p + i);
                else {
                    uint32_t w = (uint32_t)p[4 * i] | ((uint32_t)p[4 * i + 1] << 8) | ((uint32_t)p[4 * i + 2] << 16) |
                                 ((uint32_t)p[4 * i + 3] << 24);
                    out[i] = __uint_as_float(w);
                }
            } else out[i] = (float)p[i] / 255.0f;
        } else out[i] = (i == 3) ? 1.0f : 0.0f;
    }
}

__device__ __forceinline__ uint32_t find_draw(const uint32_t *base, uint32_t n, uint32_t g)
{
    uint32_t lo = 0, hi = n;            /* base[lo] <= g < base[hi] */
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (__ldg(base + mid) <= g) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256) k_vertex(BatchDev b)
{
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= b.n_vertices) return;
    uint32_t d = (b.n_draws == 1) ? 0u : find_draw(b.draw_vbase, b.n_draws, g);
    const DevDraw &dr = b.draws[d];
    uint32_t i = g - dr.vbase;

    float px, py, pz, nx, ny, nz, s, t;
    Color4 cur;
    const mtgl_state *st;
    if (dr.source == MTGL_SRC_STAGED) {
        const mtgl_in_vertex *iv = b.staged + dr.first_staged + i;
        px = iv->position[0]; py = iv->position[1]; pz = iv->position[2];
        cur = { iv->color[0], iv->color[1], iv->color[2], iv->color[3] };
        s = iv->texcoord[0]; t = iv->texcoord[1];
        nx = iv->normal[0]; ny = iv->normal[1]; nz = iv->normal[2];
        st = b.states + iv->state;
    } else {
        int32_t idx;
        if (dr.index_type) {
            uint32_t u = 0;
            if (dr.index_type == G_UNSIGNED_SHORT) {
                if (2ull * i + 2 <= dr.index_avail) u = (uint32_t)dr.index_ptr[2 * i] | ((uint32_t)dr.index_ptr[2 * i + 1] << 8);
            } else if (dr.index_type == G_UNSIGNED_INT) {
                if (4ull * i + 4 <= dr.index_avail)
                    u = (uint32_t)dr.index_ptr[4 * i] | ((uint32_t)dr.index_ptr[4 * i + 1] << 8) |
                        ((uint32_t)dr.index_ptr[4 * i + 2] << 16) | ((uint32_t)dr.index_ptr[4 * i + 3] << 24);
            } else if ((uint64_t)i < dr.index_avail) u = dr.index_ptr[i];
            idx = (int32_t)u;
        } else idx = dr.first + (int32_t)i;
        float p[4], c[4], tc[2], n[3];
        fetch_attrib(dr.position, idx, p, 4);
        px = p[0]; py = p[1]; pz = (dr.position.size == 2) ? 0.0f : p[2];
        cur = { dr.cur_color[0], dr.cur_color[1], dr.cur_color[2], dr.cur_color[3] };
        if (dr.color.enabled) {            /* glColor4f sanitising, gl_api.c:708-722 */
            fetch_attrib(dr.color, idx, c, 4);
            float r = c[0], gg = c[1], bb = c[2], a = c[3];
            if (isnan(r) || isinf(r)) r = 0.0f;
            if (isnan(gg) || isinf(gg)) gg = 0.0f;
            if (isnan(bb) || isinf(bb)) bb = 0.0f;
            if (isnan(a) || isinf(a)) a = 1.0f;
            cur = { sat01(r), sat01(gg), sat01(bb), sat01(a) };
        }
        s = dr.cur_texcoord[0]; t = dr.cur_texcoord[1];
        if (dr.texcoord.enabled) { fetch_attrib(dr.texcoord, idx, tc, 2); s = tc[0]; t = tc[1]; }
        nx = dr.cur_normal[0]; ny = dr.cur_normal[1]; nz = dr.cur_normal[2];
        if (dr.normal.enabled) { fetch_attrib(dr.normal, idx, n, 3); nx = n[0]; ny = n[1]; nz = n[2]; }
        st = b.states + dr.vertex_state;
    }

    /* eye = MV * (x, y, z, 1)   (graphics.h:131-138: ((m0*x + m4*y) + m8*z) + m12*w) */
    const float *mv = st->modelview;
    float ex = mv[0] * px + mv[4] * py + mv[8] * pz + mv[12] * 1.0f;
    float ey = mv[1] * px + mv[5] * py + mv[9] * pz + mv[13] * 1.0f;
    float ez = mv[2] * px + mv[6] * py + mv[10] * pz + mv[14] * 1.0f;
    float ew = mv[3] * px + mv[7] * py + mv[11] * pz + mv[15] * 1.0f;

    /* eye normal = normalize(N * (n, 0)); the fourth column of N is zero (graphics.h:247-257) */
    const float *nm = st->normal;
    float enx = nm[0] * nx + nm[4] * ny + nm[8] * nz + 0.0f;
    float eny = nm[1] * nx + nm[5] * ny + nm[9] * nz + 0.0f;
    float enz = nm[2] * nx + nm[6] * ny + nm[10] * nz + 0.0f;
    normalize3(enx, eny, enz);

    Color4 vc = cur;
    if ((st->caps & MTGL_CAP_LIGHTING) && st->shade_model != G_PHONG) {
        MaterialRegs mat;
        load_material(mat, &st->material_front);
        if (st->caps & MTGL_CAP_COLOR_MATERIAL) {          /* gl_api.c:285-312, front material only matters here */
            uint32_t face = st->color_material_face, mode = st->color_material_mode;
            if (face == G_FRONT || face == G_FRONT_AND_BACK) {
                Color4 k = color_clamp(cur);
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

    /* clip = P * eye   (raster.c:48-56) */
    const float *pr = st->projection;
    float cx = pr[0] * ex + pr[4] * ey + pr[8] * ez + pr[12] * ew;
    float cy = pr[1] * ex + pr[5] * ey + pr[9] * ez + pr[13] * ew;
    float cz = pr[2] * ex + pr[6] * ey + pr[10] * ez + pr[14] * ew;
    float cw = pr[3] * ex + pr[7] * ey + pr[11] * ez + pr[15] * ew;

    /* texture matrix, divide by q only when q is neither 0 nor 1 (gl_api.c:327-335) */
    const float *tm = st->texture;
    float tu = tm[0] * s + tm[4] * t + tm[8] * 0.0f + tm[12] * 1.0f;
    float tv = tm[1] * s + tm[5] * t + tm[9] * 0.0f + tm[13] * 1.0f;
    float tq = tm[3] * s + tm[7] * t + tm[11] * 0.0f + tm[15] * 1.0f;
    if (tq != 0.0f && tq != 1.0f) { tu = tu / tq; tv = tv / tq; }

    b.v_clip[g] = make_float4(cx, cy, cz, cw);
    b.v_color[g] = make_float4(vc.r, vc.g, vc.b, vc.a);
    b.v_tex[g] = make_float4(tu, tv, -ez, 0.0f);
    if (b.need_eye) {
        b.v_epos[g] = make_float4(ex, ey, ez, 0.0f);
        b.v_enrm[g] = make_float4(enx, eny, enz, 0.0f);
    }
}

void launch_vertex_stage(const BatchDev &b, cudaStream_t s)
{
    if (b.n_vertices == 0) return;
    uint32_t blocks = (b.n_vertices + 255) / 256;
    k_vertex<<<blocks, 256, 0, s>>>(b);
    note_launch();
}

} // namespace mtgl_dev_impl
