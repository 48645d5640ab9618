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
#include "dev_vertex.cuh"

#include <cstdlib>

namespace mtgl_dev_impl {

static unsigned long long g_launches = 0;
uint64_t kernel_launch_count() { return g_launches; }
void note_launch() { g_launches++; }

bool small_grid(uint32_t tiles, uint32_t limit)
{
    if (const char *e = std::getenv("MTGL_GRID_SHAPE")) {
        if (e[0] == 's') return true;
        if (e[0] == 'l') return false;
    }
    return tiles <= limit;
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
    if (dr.fused) return;               /* independent triangles: k_setup shades the vertices of the survivors itself */
    VertexIn in;
    if (dr.shared_verts) {              /* one vertex per buffer element (glDrawElements reuses them); the last slot holds the defaults */
        const uint32_t e = g - dr.vbase;
        fetch_vertex(b.staged, b.states, dr, e < dr.shared_verts ? e : 0xFFFFFFFFu, in, true);
    } else fetch_vertex(b.staged, b.states, dr, g - dr.vbase, in);
    VertexOut o;
    shade_vertex(in, o);
    b.v_clip[g] = o.clip;
    b.v_color[g] = o.color;
    b.v_tex[g] = o.tex;
    if (b.need_eye) {
        b.v_epos[g] = o.epos;
        b.v_enrm[g] = o.enrm;
    }
}

void launch_vertex_stage(const BatchDev &b, cudaStream_t s)
{
    if (b.n_vertices == 0 || b.n_unfused_draws == 0) return;
    uint32_t blocks = (b.n_vertices + 255) / 256;
    k_vertex<<<blocks, 256, 0, s>>>(b);
    note_launch();
}

} // namespace mtgl_dev_impl
