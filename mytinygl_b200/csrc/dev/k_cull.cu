/*
 * k_cull.cu -- hierarchical culling in front of K2: whole 256-triangle chunks of an array draw are dropped when the
 * object-space bounding box of their vertices projects entirely outside the rows (sort-first band) or columns this
 * device renders.
 *
 * The reference has no such stage: render_triangle (src/raster.c:901-958) clips, snaps and boxes every triangle, and
 * rasterize_triangle_smooth (458-499) then finds the clamped box empty.  Dropping a chunk here gives the same result
 * -- no record, no fragment -- as long as the test is conservative, so it is: a chunk is dropped only when
 *
 *   - it lies inside one non-indexed GL_TRIANGLES array draw whose two polygon modes are GL_FILL (a filled triangle
 *     touches nothing outside the box of its three snapped vertices; lines and points have a width);
 *   - all eight corners of its box have w > 0 by a wide margin (then the projected vertices of every triangle, and of
 *     every polygon Sutherland-Hodgman makes of it, lie inside the projected hull of the corners);
 *   - the corners' window coordinates, widened by a first-order bound of the float rounding error of the reference's
 *     transform chain (MV, P, divide, viewport map: raster.c:48-63, 729-746) plus two pixels for the truncating
 *     snap, all fall on one side of the band / framebuffer.
 *
 * Two kernels: k_chunk_bounds (one warp per chunk, once per buffer content -- the host caches the boxes until the
 * buffer changes) and k_chunk_cull (eight lanes per chunk, one box corner each, every frame).  K2's CTAs read one
 * byte and leave.  What it buys: on a band of 1/8 of the frame K2 runs for 1/6 of the chunks instead of deciding
 * all 1 029 952 triangles of C4 on every GPU (SURVEY.md 8e "per-draw bounding-box band cull").
 */
#include "dev_common.cuh"

#include <cfloat>

namespace mtgl_dev_impl {

void note_launch();

/* box of chunk c (vertices 768 c .. 768 c + 767 of the draw) as two float4: (min.xyz, valid) (max.xyz, 0);
 * valid = 0 when a coordinate is not finite */
__global__ void __launch_bounds__(256) k_chunk_bounds(const uint8_t *pos, uint32_t stride, uint32_t size, int32_t first,
                                                      uint32_t nverts, uint32_t nchunks, float4 *out)
{
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= nchunks) return;
    const uint32_t v0 = warp * 3u * SETUP_THREADS, v1 = min(v0 + 3u * SETUP_THREADS, nverts);
    float lo[3] = { FLT_MAX, FLT_MAX, FLT_MAX }, hi[3] = { -FLT_MAX, -FLT_MAX, -FLT_MAX };
    bool finite = true;
    for (uint32_t v = v0 + lane; v < v1; v += 32) {
        const float *p = reinterpret_cast<const float *>(pos + (uint64_t)((uint32_t)first + v) * stride);
        float c[3] = { __ldg(p), __ldg(p + 1), size >= 3 ? __ldg(p + 2) : 0.0f };
#pragma unroll
        for (int k = 0; k < 3; k++) {
            finite = finite && (fabsf(c[k]) <= FLT_MAX);        /* false for NaN and infinities */
            lo[k] = fminf(lo[k], c[k]); hi[k] = fmaxf(hi[k], c[k]);
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1)
#pragma unroll
        for (int k = 0; k < 3; k++) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xFFFFFFFFu, lo[k], o));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xFFFFFFFFu, hi[k], o));
        }
    finite = __all_sync(0xFFFFFFFFu, finite);
    if (lane == 0) {
        out[2 * warp] = make_float4(lo[0], lo[1], lo[2], (finite && v1 > v0) ? 1.0f : 0.0f);
        out[2 * warp + 1] = make_float4(hi[0], hi[1], hi[2], 0.0f);
    }
}

__device__ __forceinline__ uint32_t draw_of_triangle(const uint32_t *base, uint32_t n, uint32_t g)
{
    uint32_t lo = 0, hi = n;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(base + mid) <= g) lo = mid; else hi = mid;
    }
    return lo;
}

/* chunk_cull[c] = 1 when chunk c of the pass cannot produce a record on this device.  Eight lanes per chunk, one box
 * corner each (the corner transform is a dependent chain of ~150 operations: spread out, the whole pass is a few
 * microseconds in front of K2). */
__global__ void __launch_bounds__(256) k_chunk_cull(BatchDev b, FrameTargets fb, uint32_t nchunks)
{
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t c = g >> 3, k = g & 7u;
    const uint32_t group = 0xFFu << (threadIdx.x & 24u);        /* the 8 lanes of this chunk */
    bool ok = c < nchunks;
    float sx = 0.0f, sy = 0.0f, dsx = 0.0f, dsy = 0.0f;
    if (ok) {
        ok = false;
        do {
            const uint32_t t0 = c * SETUP_THREADS, t1 = min(t0 + SETUP_THREADS, b.n_triangles) - 1u;
            const uint32_t d0 = (b.n_draws == 1) ? 0u : draw_of_triangle(b.draw_tbase, b.n_draws, t0);
            if (t1 >= b.draw_tbase[d0 + 1]) break;              /* the chunk spans two draws */
            const DevDraw &dr = b.draws[d0];
            if (!dr.bounds) break;
            const mtgl_state *rs = b.states + dr.raster_state;
            if (rs->polygon_mode_front != G_FILL || rs->polygon_mode_back != G_FILL) break;
            /* the pass's chunk may straddle two of the draw's own chunks */
            const uint32_t lc0 = (t0 - dr.tbase) / SETUP_THREADS, lc1 = (t1 - dr.tbase) / SETUP_THREADS;
            float4 lo = __ldg(dr.bounds + 2 * lc0), hi = __ldg(dr.bounds + 2 * lc0 + 1);
            if (lo.w == 0.0f) break;
            if (lc1 != lc0) {
                const float4 lo1 = __ldg(dr.bounds + 2 * lc1), hi1 = __ldg(dr.bounds + 2 * lc1 + 1);
                if (lo1.w == 0.0f) break;
                lo.x = fminf(lo.x, lo1.x); lo.y = fminf(lo.y, lo1.y); lo.z = fminf(lo.z, lo1.z);
                hi.x = fmaxf(hi.x, hi1.x); hi.y = fmaxf(hi.y, hi1.y); hi.z = fmaxf(hi.z, hi1.z);
            }
            const mtgl_state *vs = b.states + dr.vertex_state;
            const float *mv = vs->modelview, *pr = vs->projection;
            const float vx = (float)rs->viewport[0], vy = (float)rs->viewport[1], vw = (float)rs->viewport[2], vh = (float)rs->viewport[3];
            const float x = (k & 1) ? hi.x : lo.x, y = (k & 2) ? hi.y : lo.y, z = (k & 4) ? hi.z : lo.z;
            float e[4], ae[4], cl[4], ac[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {                       /* eye = MV v, clip = P eye, each with the sum of |terms| */
                e[i] = mv[i] * x + mv[4 + i] * y + mv[8 + i] * z + mv[12 + i] * 1.0f;
                ae[i] = fabsf(mv[i]) * fabsf(x) + fabsf(mv[4 + i]) * fabsf(y) + fabsf(mv[8 + i]) * fabsf(z) + fabsf(mv[12 + i]);
            }
#pragma unroll
            for (int i = 0; i < 4; i++) {
                cl[i] = pr[i] * e[0] + pr[4 + i] * e[1] + pr[8 + i] * e[2] + pr[12 + i] * e[3];
                ac[i] = fabsf(pr[i]) * ae[0] + fabsf(pr[4 + i]) * ae[1] + fabsf(pr[8 + i]) * ae[2] + fabsf(pr[12 + i]) * ae[3];
            }
            /* a vertex inside the box differs from the exact affine image by at most ~8 roundings of the |term| sums;
             * 2^-19 is sixteen times that */
            const float ex = ac[0] * 1.9073486e-6f, ey = ac[1] * 1.9073486e-6f, ew = ac[3] * 1.9073486e-6f;
            const float w = cl[3];
            if (!(w > 1e-4f) || !(w > 8.0f * ew) || !(ac[0] <= FLT_MAX) || !(ac[1] <= FLT_MAX) || !(ac[3] <= FLT_MAX)) break;
            const float nx = cl[0] / w, ny = cl[1] / w;
            const float dnx = (ex + fabsf(nx) * ew) / (w - ew) * 1.01f, dny = (ey + fabsf(ny) * ew) / (w - ew) * 1.01f;
            sx = (nx + 1.0f) * 0.5f * vw + vx; sy = (1.0f - ny) * 0.5f * vh + vy;       /* raster.c:59-63 before the cast */
            /* + 2 pixels: the truncating snap moves a coordinate by less than one */
            dsx = dnx * 0.5f * fabsf(vw) + 1e-5f * (fabsf(sx) + fabsf(vw) + fabsf(vx)) + 2.0f;
            dsy = dny * 0.5f * fabsf(vh) + 1e-5f * (fabsf(sy) + fabsf(vh) + fabsf(vy)) + 2.0f;
            if (!(fabsf(sx) < 1e7f) || !(fabsf(sy) < 1e7f) || !(dsx < 4.0f) || !(dsy < 4.0f)) break;
            ok = true;
        } while (false);
    }
    const bool all_ok = (__ballot_sync(0xFFFFFFFFu, ok) & group) == group;
    float sx_min = sx - dsx, sx_max = sx + dsx, sy_min = sy - dsy, sy_max = sy + dsy;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        sx_min = fminf(sx_min, __shfl_xor_sync(0xFFFFFFFFu, sx_min, o)); sx_max = fmaxf(sx_max, __shfl_xor_sync(0xFFFFFFFFu, sx_max, o));
        sy_min = fminf(sy_min, __shfl_xor_sync(0xFFFFFFFFu, sy_min, o)); sy_max = fmaxf(sy_max, __shfl_xor_sync(0xFFFFFFFFu, sy_max, o));
    }
    if (k == 0 && c < nchunks) {
        const bool cull = all_ok && (sy_max < (float)fb.band_y0 || sy_min > (float)(fb.band_y1 - 1) ||
                                     sx_max < 0.0f || sx_min > (float)(fb.width - 1));
        b.chunk_cull[c] = cull ? 1 : 0;
        if (cull) atomicAdd(&b.counters->culled_chunks, 1u);        /* statistics only */
    }
}

void launch_chunk_bounds(const uint8_t *pos, uint32_t stride, uint32_t size, int32_t first, uint32_t nverts, float4 *out, cudaStream_t s)
{
    const uint32_t nchunks = (nverts + 3u * SETUP_THREADS - 1u) / (3u * SETUP_THREADS);
    if (nchunks == 0) return;
    k_chunk_bounds<<<(nchunks + 7u) / 8u, 256, 0, s>>>(pos, stride, size, first, nverts, nchunks, out);
    note_launch();
}

void launch_chunk_cull(const BatchDev &b, const FrameTargets &fb, cudaStream_t s)
{
    const uint32_t nchunks = (b.n_triangles + SETUP_THREADS - 1) / SETUP_THREADS;
    if (nchunks == 0 || !b.chunk_cull) return;
    k_chunk_cull<<<(nchunks * 8u + 255u) / 256u, 256, 0, s>>>(b, fb, nchunks);
    note_launch();
}

} // namespace mtgl_dev_impl
