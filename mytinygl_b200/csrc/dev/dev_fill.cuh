/*
 * dev_fill.cuh -- which in-order tiles the pixel-owner kernel (k_fill.cu) takes over from the general tile kernel
 * (k_raster.cu).  Both kernels are launched over the same grid of tiles and evaluate this one deterministic predicate
 * on the same data, so exactly one of them renders a tile.
 */
#ifndef MTGL_DEV_FILL_CUH
#define MTGL_DEV_FILL_CUH

#include "dev_common.cuh"

namespace mtgl_dev_impl {

constexpr uint32_t FILL_MAX_LIST = 1024;        /* longest tile list k_fill sorts in shared memory */
constexpr uint32_t FILL_MIN_AVG_AREA = 1024;    /* mean clamped box area (pixels of the 64x64 tile) per list entry from which
                                                 * one-thread-per-pixel beats one-warp-per-8x4-block: a quarter of the tile */
enum : uint32_t { FILL_OFF = 0u, FILL_AUTO = 1u, FILL_ALWAYS = 2u };
/* pseudo-flags next to the RC_* bits in RasterPlan::in_order_all -- properties every state of the pass has, so that the
 * kernel instance for the fill-rate state mix (BASELINE C3, and the usual cut-out / transparency set-up) does not decode
 * them per triangle: */
constexpr uint32_t FILL_ALPHA_GREATER = 1u << 30;   /* every alpha test is GL_GREATER */
constexpr uint32_t FILL_BLEND_ALPHA = 1u << 29;     /* every blend is (GL_SRC_ALPHA, GL_ONE_MINUS_SRC_ALPHA) */
constexpr uint32_t FILL_STENCIL_ALWAYS = 1u << 28;  /* every stencil test is GL_ALWAYS */
constexpr uint32_t FILL_FULL_MASK = 1u << 27;       /* every colour mask is (1, 1, 1, 1) */
constexpr uint32_t FILL_MODULATE = 1u << 26;        /* every texture environment is GL_MODULATE */
constexpr uint32_t FILL_PSEUDO = FILL_ALPHA_GREATER | FILL_BLEND_ALPHA | FILL_STENCIL_ALWAYS | FILL_FULL_MASK | FILL_MODULATE;

__host__ __device__ __forceinline__ uint32_t fill_pseudo_flags(const RasterCfg &c)
{
    uint32_t f = 0;
    if (!(c.flags & RC_ALPHA_TEST) || c.alpha_func == 4u) f |= FILL_ALPHA_GREATER;
    if (!(c.flags & RC_BLEND) || (c.blend_src == G_SRC_ALPHA && c.blend_dst == G_ONE_MINUS_SRC_ALPHA)) f |= FILL_BLEND_ALPHA;
    if (!(c.flags & RC_STENCIL) || c.stencil_func == 7u) f |= FILL_STENCIL_ALWAYS;
    if (c.color_mask == 0xFu) f |= FILL_FULL_MASK;
    if (!(c.flags & RC_TEXTURED) || c.tex_env_mode == G_MODULATE) f |= FILL_MODULATE;
    return f;
}

/* Called by every thread of the CTA (blockDim.x threads; *acc is a shared word nobody else uses).  The answer is uniform.
 * tflags: bit 0 = the tile holds a record that needs in-order shading, bit 2 = it holds a line or a point. */
__device__ __forceinline__ bool fill_owns_tile(const BatchDev &b, uint32_t fill_mode, uint32_t L, uint32_t tflags, const uint32_t *list,
                                               int px0, int py0, int vh, uint32_t *acc)
{
    if (fill_mode == FILL_OFF || L == 0u || L > FILL_MAX_LIST || (tflags & 5u) != 1u) return false;
    if (fill_mode == FILL_ALWAYS) return true;
    if (threadIdx.x == 0) *acc = 0u;
    __syncthreads();
    uint32_t area = 0;
    for (uint32_t i = threadIdx.x; i < L; i += blockDim.x) {
        const uint4 row = __ldg(b.bin_rows + list[i]);          /* bbox_min, bbox_max, state_flags, id */
        const int x0 = max((int)(row.x & 0xFFFFu) - px0, 0), y0 = max((int)(row.x >> 16) - py0, 0);
        const int x1 = min((int)(row.y & 0xFFFFu) - px0, TILE_W - 1), y1 = min((int)(row.y >> 16) - py0, vh - 1);
        if (x1 >= x0 && y1 >= y0) area += (uint32_t)((x1 - x0 + 1) * (y1 - y0 + 1));
    }
    area = __reduce_add_sync(0xFFFFFFFFu, area);
    if ((threadIdx.x & 31) == 0 && area) atomicAdd(acc, area);
    __syncthreads();
    const bool owns = *acc >= L * FILL_MIN_AVG_AREA;
    __syncthreads();        /* *acc may be reused by the caller */
    return owns;
}

} // namespace mtgl_dev_impl

#endif
