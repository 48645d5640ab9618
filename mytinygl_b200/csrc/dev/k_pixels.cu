/*
 * k_pixels.cu -- pixel rectangles on the device: glDrawPixels (src/gl_api.c:1286-1373) and the gather half of
 * glReadPixels (1180-1230), SURVEY.md 8(f) rank 3.
 *
 * The reference loops over the rectangle on the CPU, touching the colour and depth planes directly; a back end whose
 * planes live in HBM would have to pull both planes to the host and push them back for every blit.  Here the
 * rectangle goes to the device (width x height x 1..4 bytes) and one thread per four destination pixels does the reference's
 * per-pixel sequence in the reference's arithmetic: format expansion, alpha test on a / 255, depth test of depth 0
 * against the stored depth, blend with the 8-bit destination, depth write, truncating pack.  Every destination pixel
 * is written by exactly one source pixel, so the rectangle's pixels are independent and the kernel runs in stream
 * order between the batches in front of and behind it -- no host synchronisation.
 *
 * glReadPixels needs the host to wait by definition, but only for its rectangle: the kernel flips, crops and
 * repacks (RGBA / RGB, zeros outside the framebuffer, alpha 255 outside its columns) into a staging buffer that is
 * then copied out -- width x height x bpp bytes over PCIe instead of whole planes.
 *
 * Algorithmic bytes per rectangle pixel: bpp in; 4 B colour write, + 4 B colour read when blending, + 4 B depth read
 * when depth testing, + 4 B depth write when the mask allows.
 */
#include "dev_common.cuh"

namespace mtgl_dev_impl {

void note_launch();

constexpr uint32_t PX_RGB = 0x1907, PX_RGBA = 0x1908, PX_LUMINANCE = 0x1909, PX_LUMINANCE_ALPHA = 0x190A;

struct PixelOp {
    int32_t x, y, width, height;    /* raster position (GL window coordinates, origin bottom-left) and size */
    uint32_t format;
    uint32_t alpha_on, depth_on, blend_on, depth_mask;
    uint32_t alpha_func, depth_func;        /* 0..7 = GL_NEVER..GL_ALWAYS, 7 for anything else (the helpers' default) */
    float alpha_ref;
    uint32_t blend_src, blend_dst;
};

/* one destination pixel, planes' values already loaded (the loads of a thread are issued together, not behind the tests) */
__device__ __forceinline__ void draw_one(const PixelOp &op, uint32_t r, uint32_t g, uint32_t b, uint32_t a, float stored_depth, uint32_t dst_color,
                                         const float *un, bool &write, float &depth_out, bool &depth_write, uint32_t &color_out)
{
    write = false; depth_write = false; depth_out = stored_depth; color_out = dst_color;
    if (op.alpha_on && !compare_f(op.alpha_func, un[a], op.alpha_ref)) return;           /* a / 255.0f, gl_api.c:1341-1346 */
    if (op.depth_on && !compare_f(op.depth_func, 0.0f, stored_depth)) return;            /* pixels sit at depth 0, 1297, 1349-1354 */
    Color4 s = { un[r], un[g], un[b], un[a] };
    if (op.blend_on) {                                       /* gl_api.c:1359-1365 */
        const Color4 d = color_unpack(dst_color, un);
        const Color4 sf = blend_factor(op.blend_src, s, d), df = blend_factor(op.blend_dst, s, d);
        s = color_clamp({ s.r * sf.r + d.r * df.r, s.g * sf.g + d.g * df.g, s.b * sf.b + d.b * df.b, s.a * sf.a + d.a * df.a });
    }
    if (op.depth_on && op.depth_mask) { depth_out = 0.0f; depth_write = true; }
    color_out = color_pack(s);
    write = true;
}

/* One thread per group of four destination pixels of a rectangle row.  Groups are aligned to the DESTINATION (16-byte
 * colour / depth accesses when the whole group is inside the framebuffer row); the rectangle's bytes are read per
 * pixel -- their alignment follows the raster position, not the framebuffer's. */
__global__ void __launch_bounds__(256) k_draw_pixels(PixelOp op, const uint8_t *__restrict__ src, FrameTargets fb, const float *__restrict__ unorm8)
{
    __shared__ float un[256];
    un[threadIdx.x] = unorm8[threadIdx.x];
    __syncthreads();
    const int row = blockIdx.y;
    const int fy = fb.height - 1 - (op.y + row);            /* gl_api.c:1304-1306: GL's origin is bottom-left */
    if (fy < fb.band_y0 || fy >= fb.band_y1 || fy < 0 || fy >= fb.height) return;
    /* destination columns [gx, gx + 4), gx a multiple of 4 */
    const int first_group = op.x >= 0 ? (op.x >> 2) : -((-op.x + 3) >> 2);
    const int gx = (first_group + (int)(blockIdx.x * blockDim.x + threadIdx.x)) * 4;
    if (gx >= op.x + op.width || gx >= fb.width || gx + 3 < 0) return;
    const size_t at = (size_t)fy * fb.width + gx;
    const bool full = gx >= 0 && gx + 3 < fb.width && (fb.width & 3) == 0;
    uint32_t dc[4] = { 0u, 0u, 0u, 0u };
    float dd[4] = { 0.0f, 0.0f, 0.0f, 0.0f };
    uint32_t pr[4], pg[4], pb[4], pa[4];
    bool in[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {                            /* the rectangle's bytes (gl_api.c:1318-1338) */
        const int dx = gx + k, col = dx - op.x;
        in[k] = col >= 0 && col < op.width && dx >= 0 && dx < fb.width;
        pr[k] = pg[k] = pb[k] = 0u; pa[k] = 255u;
        if (!in[k]) continue;
        const size_t i = (size_t)row * op.width + col;
        switch (op.format) {
        case PX_RGBA: { const uchar4 v = *reinterpret_cast<const uchar4 *>(src + i * 4); pr[k] = v.x; pg[k] = v.y; pb[k] = v.z; pa[k] = v.w; break; }
        case PX_RGB: pr[k] = src[i * 3]; pg[k] = src[i * 3 + 1]; pb[k] = src[i * 3 + 2]; break;
        case PX_LUMINANCE: pr[k] = pg[k] = pb[k] = src[i]; break;
        case PX_LUMINANCE_ALPHA: pr[k] = pg[k] = pb[k] = src[i * 2]; pa[k] = src[i * 2 + 1]; break;
        default: in[k] = false; break;
        }
    }
    if (full) {
        if (op.blend_on) { const uint4 v = *reinterpret_cast<const uint4 *>(fb.color + at); dc[0] = v.x; dc[1] = v.y; dc[2] = v.z; dc[3] = v.w; }
        if (op.depth_on) { const float4 v = *reinterpret_cast<const float4 *>(fb.depth + at); dd[0] = v.x; dd[1] = v.y; dd[2] = v.z; dd[3] = v.w; }
    } else {
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (in[k]) {
                if (op.blend_on) dc[k] = fb.color[at + k];
                if (op.depth_on) dd[k] = fb.depth[at + k];
            }
    }
    bool wr[4], dw[4];
    bool all_wr = true, any_dw = false, all_dw = true;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        wr[k] = dw[k] = false;
        if (in[k]) draw_one(op, pr[k], pg[k], pb[k], pa[k], dd[k], dc[k], un, wr[k], dd[k], dw[k], dc[k]);
        all_wr = all_wr && wr[k]; any_dw = any_dw || dw[k]; all_dw = all_dw && dw[k];
    }
    if (full && all_wr) {
        const uint4 v = make_uint4(dc[0], dc[1], dc[2], dc[3]);
        *reinterpret_cast<uint4 *>(fb.color + at) = v;
        if (fb.present) *reinterpret_cast<uint4 *>(fb.present + at) = v;     /* multi-GPU: the presenting GPU's plane */
        if (all_dw) *reinterpret_cast<float4 *>(fb.depth + at) = make_float4(dd[0], dd[1], dd[2], dd[3]);
        else if (any_dw) {
#pragma unroll
            for (int k = 0; k < 4; k++) if (dw[k]) fb.depth[at + k] = dd[k];
        }
    } else {
#pragma unroll
        for (int k = 0; k < 4; k++)
            if (wr[k]) {
                if (dw[k]) fb.depth[at + k] = dd[k];
                fb.color[at + k] = dc[k];
                if (fb.present) fb.present[at + k] = dc[k];
            }
    }
}

/* dst[(row * w + col) * bpp ..] for the rectangle whose lower-left corner is window pixel (x, y) */
__global__ void __launch_bounds__(256) k_read_pixels(FrameTargets fb, int32_t x, int32_t y, int32_t w, int32_t h, uint32_t bpp, uint8_t *__restrict__ dst)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x, row = blockIdx.y;
    if (col >= w || row >= h) return;
    const int fy = fb.height - 1 - (y + row), sx = x + col;
    uint32_t r = 0, g = 0, b = 0, a = 255u;
    const bool row_in = fy >= 0 && fy < fb.height;
    if (!row_in) a = 0u;                                     /* a row outside the framebuffer is all zeros (gl_api.c:1193-1201) */
    else if (sx >= 0 && sx < fb.width) {
        const uint32_t p = fb.color[(size_t)fy * fb.width + sx];
        r = p & 0xFFu; g = (p >> 8) & 0xFFu; b = (p >> 16) & 0xFFu; a = p >> 24;
    }
    uint8_t *o = dst + ((size_t)row * w + col) * bpp;
    o[0] = (uint8_t)r; o[1] = (uint8_t)g; o[2] = (uint8_t)b;
    if (bpp == 4) o[3] = (uint8_t)a;
}

static uint32_t func_index(uint32_t token) { return (token >= G_NEVER && token <= G_ALWAYS) ? token - G_NEVER : 7u; }

void launch_draw_pixels(const ::mtgl_pixel_rect &rect, const uint8_t *src, const FrameTargets &fb, const float *unorm8, cudaStream_t s)
{
    if (rect.width <= 0 || rect.height <= 0) return;
    PixelOp op;
    op.x = rect.x; op.y = rect.y; op.width = rect.width; op.height = rect.height; op.format = rect.format;
    op.alpha_on = (rect.caps & MTGL_CAP_ALPHA_TEST) ? 1u : 0u;
    op.depth_on = (rect.caps & MTGL_CAP_DEPTH_TEST) ? 1u : 0u;
    op.blend_on = (rect.caps & MTGL_CAP_BLEND) ? 1u : 0u;
    op.depth_mask = rect.depth_mask;
    op.alpha_func = func_index(rect.alpha_func); op.depth_func = func_index(rect.depth_func);
    op.alpha_ref = rect.alpha_ref;
    op.blend_src = rect.blend_src; op.blend_dst = rect.blend_dst;
    const uint32_t groups = (uint32_t)(rect.width + 3) / 4u + 1u;            /* destination-aligned groups of four columns */
    const dim3 grid((groups + 255u) / 256u, (uint32_t)rect.height);
    k_draw_pixels<<<grid, 256, 0, s>>>(op, src, fb, unorm8);
    note_launch();
}

void launch_read_pixels(const FrameTargets &fb, int32_t x, int32_t y, int32_t w, int32_t h, uint32_t bpp, uint8_t *dst, cudaStream_t s)
{
    if (w <= 0 || h <= 0) return;
    const dim3 grid((uint32_t)(w + 255) / 256u, (uint32_t)h);
    k_read_pixels<<<grid, 256, 0, s>>>(fb, x, y, w, h, bpp, dst);
    note_launch();
}

} // namespace mtgl_dev_impl
