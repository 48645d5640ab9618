/*
 * k_pixels.cu -- pixel rectangles on the device: glDrawPixels (src/gl_api.c:1286-1373) and the gather half of
 * glReadPixels (1180-1230), SURVEY.md 8(f) rank 3.
 *
 * The reference loops over the rectangle on the CPU, touching the colour and depth planes directly; a back end whose
 * planes live in HBM would have to pull both planes to the host and push them back for every blit.  Here the
 * rectangle goes to the device (width x height x 1..4 bytes) and one thread per source pixel does the reference's
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

__global__ void __launch_bounds__(256) k_draw_pixels(PixelOp op, const uint8_t *__restrict__ src, FrameTargets fb, const float *__restrict__ unorm8)
{
    const int col = blockIdx.x * blockDim.x + threadIdx.x, row = blockIdx.y;
    if (col >= op.width) return;
    const int fy = fb.height - 1 - (op.y + row);            /* gl_api.c:1304-1306: GL's origin is bottom-left */
    const int dx = op.x + col;
    if (fy < fb.band_y0 || fy >= fb.band_y1 || fy < 0 || fy >= fb.height || dx < 0 || dx >= fb.width) return;
    const size_t i = (size_t)row * op.width + col;
    uint32_t r, g, b, a = 255u;
    switch (op.format) {                                     /* gl_api.c:1318-1338 */
    case PX_RGBA: { const uchar4 v = *reinterpret_cast<const uchar4 *>(src + i * 4); r = v.x; g = v.y; b = v.z; a = v.w; break; }
    case PX_RGB: r = src[i * 3]; g = src[i * 3 + 1]; b = src[i * 3 + 2]; break;
    case PX_LUMINANCE: r = g = b = src[i]; break;
    case PX_LUMINANCE_ALPHA: r = g = b = src[i * 2]; a = src[i * 2 + 1]; break;
    default: return;
    }
    if (op.alpha_on && !compare_f(op.alpha_func, unorm8[a], op.alpha_ref)) return;      /* a / 255.0f, gl_api.c:1341-1346 */
    const size_t at = (size_t)fy * fb.width + dx;
    if (op.depth_on && !compare_f(op.depth_func, 0.0f, fb.depth[at])) return;           /* pixels sit at depth 0, 1297, 1349-1354 */
    Color4 s = { unorm8[r], unorm8[g], unorm8[b], unorm8[a] };
    if (op.blend_on) {                                       /* gl_api.c:1359-1365 */
        const Color4 d = color_unpack(fb.color[at], unorm8);
        const Color4 sf = blend_factor(op.blend_src, s, d), df = blend_factor(op.blend_dst, s, d);
        s = color_clamp({ s.r * sf.r + d.r * df.r, s.g * sf.g + d.g * df.g, s.b * sf.b + d.b * df.b, s.a * sf.a + d.a * df.a });
    }
    if (op.depth_on && op.depth_mask) fb.depth[at] = 0.0f;
    const uint32_t p = color_pack(s);
    fb.color[at] = p;
    if (fb.present) fb.present[at] = p;                      /* multi-GPU: the presenting GPU's plane */
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
    const dim3 grid((uint32_t)(rect.width + 255) / 256u, (uint32_t)rect.height);
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
