/*
 * dev_fragment.cuh -- per-fragment helpers shared by the tile kernels (k_raster.cu, k_fill.cu): stencil operation,
 * fog factor, the reference's edge function, and the arithmetic forms of the two 8-bit conversions every colour
 * goes through (color_to_rgba32 / color_from_rgba32, src/graphics.h:337-357) that need neither a conversion
 * instruction nor a table look-up.
 */
#ifndef MTGL_DEV_FRAGMENT_CUH
#define MTGL_DEV_FRAGMENT_CUH

#include "dev_common.cuh"

namespace mtgl_dev_impl {

__device__ __forceinline__ uint8_t stencil_apply(uint32_t op, uint8_t v, int32_t ref)   /* raster.c:425-438 */
{
    switch (op) {
    case G_KEEP: return v;
    case G_ZERO: return 0;
    case G_REPLACE: return (uint8_t)(ref & 0xFF);
    case G_INCR: return v < 255 ? (uint8_t)(v + 1) : (uint8_t)255;
    case G_INCR_WRAP: return (uint8_t)(v + 1);
    case G_DECR: return v > 0 ? (uint8_t)(v - 1) : (uint8_t)0;
    case G_DECR_WRAP: return (uint8_t)(v - 1);
    case G_INVERT: return (uint8_t)~v;
    default: return v;
    }
}

__device__ __forceinline__ float fog_factor(const RasterCfg *c, float coord)   /* raster.c:677-701 */
{
    float f;
    switch (c->fog_mode) {
    case G_LINEAR: f = (c->fog_end != c->fog_start) ? (c->fog_end - coord) / (c->fog_end - c->fog_start) : 1.0f; break;
    case G_EXP: f = expf(-c->fog_density * coord); break;
    case G_EXP2: { float d = c->fog_density * coord; f = expf(-d * d); break; }
    default: f = 1.0f; break;
    }
    if (f < 0.0f) f = 0.0f;
    if (f > 1.0f) f = 1.0f;
    return f;
}

__device__ __forceinline__ float edge_at(float ax, float ay, float bx, float by, float px, float py)   /* raster.c:299-302 */
{
    return (px - ax) * (by - ay) - (py - ay) * (bx - ax);
}

/* One channel of color_to_rgba32 (graphics.h:337-348) as a float: (float)(uint8_t)(clamp(x) * 255.0f).
 * v = sat(x) * 255 lies in [0, 255]; v + 2^23 rounded TOWARDS ZERO is 2^23 + floor(v) exactly (the sum has an ulp of
 * 1 there), so the subtraction returns the truncated byte as a float with two adds on the FP32 pipe and no F2I.
 * __saturatef maps NaN to 0 where the reference's ternaries keep it, but the reference's NaN * 255 converts to byte 0 as
 * well (cvttss2si yields 0x80000000), so the result is the same for every input. */
__device__ __forceinline__ float byte_of(float x)
{
    return __fadd_rz(__saturatef(x) * 255.0f, 8388608.0f) - 8388608.0f;
}

/* n / 255.0f for an integral float n in [0, 255] -- one channel of color_from_rgba32 (graphics.h:350-357) -- without
 * a division or a table: with R = RN(1/255) and r = RN(1/255 - R), RN(n*R + RN(n*r)) equals the correctly rounded
 * quotient for all 256 values (checked exhaustively: tests/test_numeric_identities.py). */
__device__ __forceinline__ float unorm_of(float n)
{
    return __fmaf_rn(n, __uint_as_float(0x3B808081u), n * __uint_as_float(0xAF7EFEFFu));
}

} // namespace mtgl_dev_impl

#endif
