"""CPU checks of the arithmetic identities the tile kernels rely on (mytinygl_b200/csrc/dev/dev_fragment.cuh, k_fill.cu).

The device code replaces the reference's 8-bit conversions (src/graphics.h:337-357) and floorf/(int) pairs
(src/textures.c:379-451) by FP32 adds with directed rounding and an FMA pair.  Each replacement is an exact identity;
this file restates them with exact rational arithmetic and checks them against the reference's own expressions."""
from fractions import Fraction

import numpy as np

f32 = np.float32


def rn(fr: Fraction) -> np.float32:
    """Fraction -> nearest float32, ties to even."""
    y = f32(float(fr))
    cands = [y, np.nextafter(y, f32(np.inf)), np.nextafter(y, f32(-np.inf))]
    return f32(min(cands, key=lambda c: (abs(Fraction(float(c)) - fr), int(np.frombuffer(f32(c).tobytes(), dtype=np.uint32)[0]) & 1)))


def bits(x) -> np.float32:
    return np.frombuffer(np.uint32(x).tobytes(), dtype=np.float32)[0]


def test_unorm_of_is_the_ieee_quotient():
    """unorm_of(n) = fma(n, R, n * r) with R = RN(1/255), r = RN(1/255 - R) equals n / 255.0f for n = 0..255."""
    R, r = bits(0x3B808081), bits(0xAF7EFEFF)
    assert R == f32(1.0) / f32(255.0)
    assert r == rn(Fraction(1, 255) - Fraction(float(R)))
    for n in range(256):
        lo = rn(Fraction(n) * Fraction(float(r)))
        q = rn(Fraction(n) * Fraction(float(R)) + Fraction(float(lo)))
        assert q == f32(n) / f32(255.0), n


def test_byte_of_is_the_truncating_pack():
    """byte_of(x): (sat(x) * 255 + 2^23 rounded towards zero) - 2^23 == (float)(uint8_t)(clamp(x) * 255.0f)."""
    rng = np.random.default_rng(1234)
    xs = np.concatenate([rng.random(20000, dtype=np.float32), np.arange(256, dtype=np.float32) / f32(255.0),
                         np.nextafter(np.arange(256, dtype=np.float32) / f32(255.0), f32(0)), f32([0.0, 1.0, 0.999999, 1e-30, 0.5])])
    for x in xs:
        v = f32(x) * f32(255.0)                                   # 0 <= v <= 255
        exact = Fraction(float(v)) + Fraction(2 ** 23)
        toward_zero = Fraction(int(exact))                       # ulp is 1 in [2^23, 2^24): RZ = floor for a positive sum
        got = float(toward_zero - 2 ** 23)
        assert got == float(np.uint8(np.floor(v))), x


def test_round_down_add_is_floor():
    """floor(t) for |t| < 2^22 as the low mantissa bits of t + 1.5 * 2^23 rounded towards minus infinity."""
    rng = np.random.default_rng(99)
    ts = np.concatenate([rng.uniform(-0.5, 2047.5, 20000).astype(np.float32), f32([-0.5, -0.25, -0.0, 0.0, 0.5, 1.0, 2047.5, 63.5, 62.99999])])
    M = 12582912
    for t in ts:
        exact = Fraction(float(t)) + M
        down = exact.numerator // exact.denominator               # RM at ulp 1
        as_bits = 0x4B000000 + (down - 2 ** 23)                   # the float 2^23 + k has bits 0x4B000000 + k
        assert as_bits - 0x4B400000 == int(np.floor(t)), t
        assert float(f32(down) - f32(M)) == float(np.floor(t)), t


def test_exact_edge_form_matches_the_float_expression():
    """k_fill / k_vis evaluate e = A x + B y + C with FMAs when every product and every value stays below 2^24: then
    the reference's (px - ax) * (by - ay) - (py - ay) * (bx - ax) in float (raster.c:299-302) has no rounding either."""
    rng = np.random.default_rng(7)
    for _ in range(2000):
        ax, ay, bx, by = (int(v) for v in rng.integers(-200, 4000, 4))
        px0, py0 = int(rng.integers(0, 60)) * 64, int(rng.integers(0, 33)) * 64
        dx, dy = bx - ax, by - ay
        mx = max(abs(px0 - ax), abs(px0 + 63 - ax)); my = max(abs(py0 - ay), abs(py0 + 63 - ay))
        if mx * abs(dy) + my * abs(dx) >= 2 ** 24:
            continue
        X, Y = int(rng.integers(0, 64)), int(rng.integers(0, 64))
        px, py = f32(px0 + X), f32(py0 + Y)
        ref = (px - f32(ax)) * (f32(by) - f32(ay)) - (py - f32(ay)) * (f32(bx) - f32(ax))
        c = (px0 - ax) * dy - (py0 - ay) * dx
        assert float(ref) == float(dy * X - dx * Y + c)


def test_stencil_ops_as_data():
    """dev_fragment.cuh stencil_op_encode / stencil_op_apply against stencil_op (raster.c:425-438) for every op, value, ref."""
    KEEP, ZERO, REPLACE, INCR, DECR, INVERT, INCR_WRAP, DECR_WRAP = 0x1E00, 0, 0x1E01, 0x1E02, 0x1E03, 0x150A, 0x8507, 0x8508

    def reference(op, v, ref):
        return {KEEP: v, ZERO: 0, REPLACE: ref & 0xFF, INCR: min(v + 1, 255), INCR_WRAP: (v + 1) & 0xFF, DECR: max(v - 1, 0),
                DECR_WRAP: (v - 1) & 0xFF, INVERT: (~v) & 0xFF}[op]

    def encode(op, ref):
        return {ZERO: 0, REPLACE: (ref & 0xFF) << 8, INCR: 0xFF | (1 << 16) | (1 << 19), INCR_WRAP: 0xFF | (1 << 16),
                DECR: 0xFF | (3 << 16) | (1 << 18), DECR_WRAP: 0xFF | (3 << 16), INVERT: 0xFF | (0xFF << 8)}.get(op, 0xFF)

    def apply(enc, v):
        d = ((enc >> 16) & 1) - ((enc >> 16) & 2)
        r = ((v & enc & 0xFF) ^ ((enc >> 8) & 0xFF)) + d
        r = max(r, 0 if enc & (1 << 18) else -256)
        r = min(r, 255 if enc & (1 << 19) else 511)
        return r & 0xFF

    for op in (KEEP, ZERO, REPLACE, INCR, DECR, INVERT, INCR_WRAP, DECR_WRAP):
        for ref in (0, 1, 0x7F, 0xFF, 0x1234):
            for v in range(256):
                assert apply(encode(op, ref), v) == reference(op, v, ref), (hex(op), ref, v)


def test_comparison_masks():
    """compare_mask / compare_*_mask: GL_NEVER .. GL_ALWAYS are numbered lt = 1, eq = 2, gt = 4; unordered operands pass != and always."""
    import operator
    ops = [lambda a, b: False, operator.lt, operator.eq, operator.le, operator.gt, operator.ne, operator.ge, lambda a, b: True]
    vals = [-1.0, 0.0, 0.5, 1.0, float("inf"), float("nan")]
    for f in range(8):
        m = (f & 7) | (8 if f in (5, 7) else 0)
        for a in vals:
            for b in vals:
                code = 1 if a < b else 2 if a == b else 4 if a > b else 8
                assert bool(m & code) == bool(ops[f](a, b)), (f, a, b)
