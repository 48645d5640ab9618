"""Measurement for the pixel-rectangle path (k_pixels.cu): K full-frame glDrawPixels blits and glReadPixels calls
through the gl* API on the B200 library, next to the unmodified reference on the host CPU.

    python tools/bench_pixels.py [--width 3840 --height 2160 --steps 20]

Prints one JSON line: ms per blit end to end (host rectangle -> device, PCIe inside), the kernel's CUDA-event time
(mtgl_dev_timer marks around the launch) with its algorithmic bytes (4 B in + 4 B colour read + 4 B colour write + 4 B
depth read + 4 B depth write per pixel: RGBA, blending and depth test on) against the measured HBM bandwidth, and the
reference's time for the same calls."""
import argparse
import ctypes
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from mytinygl_b200 import load_b200  # noqa: E402
from oracle_loader import load_reference  # noqa: E402

GL_RGBA, GL_UNSIGNED_BYTE = 0x1908, 0x1401
GL_DEPTH_TEST, GL_BLEND, GL_SRC_ALPHA, GL_ONE_MINUS_SRC_ALPHA, GL_ALWAYS = 0x0B71, 0x0BE2, 0x0302, 0x0303, 0x0207


def drive(lib, w, h, steps, px, out):
    L = lib.lib
    lib.create(w, h)
    L.glClearColor(ctypes.c_float(0.2), ctypes.c_float(0.3), ctypes.c_float(0.4), ctypes.c_float(1.0))
    L.glClear(0x4000 | 0x0100)
    L.glEnable(GL_DEPTH_TEST); L.glDepthFunc(GL_ALWAYS)
    L.glEnable(GL_BLEND); L.glBlendFunc(GL_SRC_ALPHA, GL_ONE_MINUS_SRC_ALPHA)
    L.glRasterPos2i.argtypes = [ctypes.c_int, ctypes.c_int]
    L.glDrawPixels.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_uint, ctypes.c_uint, ctypes.c_void_p]
    L.glReadPixels.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_uint, ctypes.c_uint, ctypes.c_void_p]
    L.glMatrixMode(0x1701); L.glLoadIdentity()
    L.glOrtho.argtypes = [ctypes.c_double] * 6
    L.glOrtho(0.0, float(w), 0.0, float(h), -1.0, 1.0)
    L.glMatrixMode(0x1700); L.glLoadIdentity()
    L.glRasterPos2i(0, 0)
    res = {}
    for name, call in (("draw", lambda: L.glDrawPixels(w, h, GL_RGBA, GL_UNSIGNED_BYTE, px.ctypes.data)),
                       ("read", lambda: L.glReadPixels(0, 0, w, h, GL_RGBA, GL_UNSIGNED_BYTE, out.ctypes.data))):
        call(); L.glFinish()
        if steps == 0:
            continue
        t0 = time.perf_counter()
        for _ in range(steps):
            call()
        L.glFinish()
        res[name] = (time.perf_counter() - t0) * 1e3 / steps
    return res


def one_blit(lib, w, h, px):
    """fresh context, one blended + depth-tested blit over a cleared frame, read back: the parity check"""
    out = np.empty((h, w, 4), np.uint8)
    drive(lib, w, h, 0, px, out)
    lib.destroy()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--steps", type=int, default=20)
    a = ap.parse_args()
    w, h = a.width, a.height
    rng = np.random.default_rng(1234)
    px = rng.integers(0, 256, size=(h, w, 4), dtype=np.uint8)
    out = np.empty((h, w, 4), np.uint8)
    gpu = load_b200()
    r = drive(gpu, w, h, a.steps, px, out)
    # kernel-only time: the same blit between two device marks, rectangle already staged by a first call
    L = gpu.lib
    dev = gpu.device()
    L.mtgl_dev_timer_mark.argtypes = [ctypes.c_void_p, ctypes.c_int]
    L.mtgl_dev_timer_elapsed_ms.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_float)]
    ms = ctypes.c_float()
    L.mtgl_dev_timer_mark(dev, 0)
    for _ in range(a.steps):
        L.glDrawPixels(w, h, GL_RGBA, GL_UNSIGNED_BYTE, px.ctypes.data)
    L.mtgl_dev_timer_mark(dev, 1)
    L.mtgl_dev_timer_elapsed_ms(dev, ctypes.byref(ms))
    gpu.destroy()
    gpu_out = one_blit(gpu, w, h, px)
    ref = load_reference("shipped")
    rr = drive(ref, w, h, max(a.steps // 10, 1), px, out)
    ref.destroy()
    same = bool(np.array_equal(gpu_out, one_blit(load_reference("strict"), w, h, px)))      # the canonical IEEE build
    peak = 6650.0
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        peak = float(json.loads(p.read_text())["hbm_gbs"])
    print(json.dumps({"workload": f"pixel rectangles {w}x{h} RGBA, blend + depth test", "steps": a.steps,
                      "draw_pixels_ms": r["draw"], "read_pixels_ms": r["read"], "draw_pixels_device_ms_incl_h2d": ms.value / a.steps,
                      "h2d_bytes_per_draw": w * h * 4, "d2h_bytes_per_read": w * h * 4,
                      "algorithmic_bytes_per_draw": w * h * 20, "hbm_peak_gbs": peak,
                      "reference_draw_pixels_ms": rr["draw"], "reference_read_pixels_ms": rr["read"],
                      "one_blit_identical_to_strict_reference": same}))


if __name__ == "__main__":
    main()
