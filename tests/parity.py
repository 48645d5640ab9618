"""Plane comparison used by the parity tests (BASELINE.json north_star gate)."""
import numpy as np


def compare_planes(ref, got):
    """ref/got: (color u32[h,w], depth f32[h,w], stencil u8[h,w]).  Returns a dict of statistics."""
    rc, rd, rs = ref[:3]
    gc, gd, gs = got[:3]
    rb = rc.view(np.uint8).reshape(rc.shape + (4,)).astype(np.int16)
    gb = gc.view(np.uint8).reshape(gc.shape + (4,)).astype(np.int16)
    chan = np.abs(rb - gb)
    ri = rd.view(np.int32).astype(np.int64)
    gi = gd.view(np.int32).astype(np.int64)
    # IEEE bit patterns ordered as integers (all depths here are >= 0 or equal in sign)
    ulp = np.abs(ri - gi)
    return {
        "pixels": int(rc.size),
        "color_identical_frac": float((rc == gc).mean()),
        "color_max_abs": int(chan.max()),
        "color_diff_pixels": int((rc != gc).sum()),
        "depth_max_ulp": int(ulp.max()),
        "depth_diff_pixels": int((ulp != 0).sum()),
        "stencil_diff_pixels": int((rs != gs).sum()),
    }


def assert_gate(stats, what=""):
    """north_star gate: stencil (and with it the covered-fragment set) bit-exact, colour within 1/255 per
    channel with >= 99.9 % of pixels identical, depth within 1 ulp."""
    assert stats["stencil_diff_pixels"] == 0, (what, stats)
    assert stats["color_max_abs"] <= 1, (what, stats)
    assert stats["color_identical_frac"] >= 0.999, (what, stats)
    assert stats["depth_max_ulp"] <= 1, (what, stats)


def assert_exact(stats, what=""):
    assert stats["color_diff_pixels"] == 0 and stats["depth_diff_pixels"] == 0 and stats["stencil_diff_pixels"] == 0, (what, stats)
