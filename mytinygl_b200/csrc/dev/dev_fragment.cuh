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

/* The eight comparison functions GL_NEVER .. GL_ALWAYS (0x200 + f) as a mask: bit 0 = true when a < b, bit 1 when a == b,
 * bit 2 when a > b -- which is how the tokens are numbered -- and bit 3 when the operands are unordered (a NaN): the C
 * operators of depth_test / alpha_test / stencil_test (raster.c:344-357, 391-422) are all false then, except != and
 * "always".  One mask per triangle, then every test is branch-free. */
__device__ __forceinline__ uint32_t compare_mask(uint32_t f) { return (f & 7u) | (((f & 7u) == 5u || (f & 7u) == 7u) ? 8u : 0u); }

/* (the mask is the same for every fragment of a triangle: the tests on it are loop-invariant and the usual functions
 * -- always, greater, less, lequal -- cost one comparison per fragment) */
__device__ __forceinline__ bool compare_f_mask(uint32_t m, float a, float b)
{
    if (m == 15u) return true;
    if (m == 4u) return a > b;
    if (m == 1u) return a < b;
    if (m == 3u) return a <= b;
    const uint32_t idx = (a != a || b != b) ? 3u : ((a > b) ? 2u : ((a == b) ? 1u : 0u));
    return ((m >> idx) & 1u) != 0u;
}

__device__ __forceinline__ bool compare_i_mask(uint32_t m, int32_t a, int32_t b)
{
    if ((m & 7u) == 7u) return true;
    if ((m & 7u) == 2u) return a == b;
    const uint32_t idx = (a > b) ? 2u : ((a == b) ? 1u : 0u);
    return ((m >> idx) & 1u) != 0u;
}

/* A stencil operation (raster.c:425-438) as data: nv = clamp(((v & A) ^ X) + D, lo, hi) & 0xFF with
 * A = bits 0-7, X = bits 8-15, D = bits 16-17 (0, +1, 3 = -1), bit 18: lo = 0 (else none), bit 19: hi = 255 (else none). */
__device__ __forceinline__ uint32_t stencil_op_encode(uint32_t op, int32_t ref)
{
    switch (op) {
    case G_ZERO: return 0u;
    case G_REPLACE: return ((uint32_t)ref & 0xFFu) << 8;
    case G_INCR: return 0xFFu | (1u << 16) | (1u << 19);
    case G_INCR_WRAP: return 0xFFu | (1u << 16);
    case G_DECR: return 0xFFu | (3u << 16) | (1u << 18);
    case G_DECR_WRAP: return 0xFFu | (3u << 16);
    case G_INVERT: return 0xFFu | (0xFFu << 8);
    default: return 0xFFu;          /* GL_KEEP */
    }
}

struct StencilOp { uint32_t amask, xmask; int add, lo, hi; };

__device__ __forceinline__ StencilOp stencil_op_decode(uint32_t enc)
{
    StencilOp o;
    o.amask = enc & 0xFFu; o.xmask = (enc >> 8) & 0xFFu;
    o.add = (int)((enc >> 16) & 1u) - (int)((enc >> 16) & 2u);          /* 0, +1, -1 */
    o.lo = (enc & (1u << 18)) ? 0 : -256;
    o.hi = (enc & (1u << 19)) ? 255 : 511;
    return o;
}

__device__ __forceinline__ uint32_t stencil_op_apply(const StencilOp &o, uint32_t v)
{
    return (uint32_t)min(max((int)((v & o.amask) ^ o.xmask) + o.add, o.lo), o.hi) & 0xFFu;
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
