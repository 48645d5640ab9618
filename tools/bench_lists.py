"""Measurement for compiled display-list geometry: frames of N glCallList(Suzanne) under different
transforms, 800x600, lit -- with the list's geometry compiled into array draws (default) and replayed call by call
(MTGL_NO_LIST_RUNS=1), next to the unmodified reference on the host CPU.

    python tools/bench_lists.py [--calls 64 --frames 50]
Prints one JSON line per mode (run the script once per mode: the switch is read when the library is loaded)."""
import argparse
import ctypes
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from mytinygl_b200 import load_b200  # noqa: E402
from oracle_loader import load_reference  # noqa: E402

GL_COMPILE, GL_TRIANGLES, GL_DEPTH_TEST, GL_LIGHTING, GL_LIGHT0 = 0x1300, 4, 0x0B71, 0x0B50, 0x4000


def drive(lib, calls, frames):
    from mytinygl_b200 import load_suzanne
    pos, nrm, faces = load_suzanne()
    L = lib.lib
    lib.create(800, 600)
    f = ctypes.c_float
    for fn in ("glNormal3f", "glVertex3f", "glTranslatef"):
        getattr(L, fn).argtypes = [f, f, f]
    L.glRotatef.argtypes = [f, f, f, f]
    L.glFrustum.argtypes = [ctypes.c_double] * 6
    L.glMatrixMode(0x1701); L.glLoadIdentity(); L.glFrustum(-0.1333, 0.1333, -0.1, 0.1, 0.1, 100.0); L.glMatrixMode(0x1700)
    L.glEnable(GL_DEPTH_TEST); L.glEnable(GL_LIGHTING); L.glEnable(GL_LIGHT0)
    L.glGenLists.restype = ctypes.c_uint
    lst = L.glGenLists(1)
    L.glNewList(lst, GL_COMPILE)
    L.glBegin(GL_TRIANGLES)
    for tri in faces.reshape(-1, 3):
        for vi in tri:
            L.glNormal3f(*[float(x) for x in nrm[vi]])
            L.glVertex3f(*[float(x) for x in pos[vi]])
    L.glEnd()
    L.glEndList()

    def frame():
        L.glClear(0x4000 | 0x0100)
        for k in range(calls):
            L.glLoadIdentity()
            L.glTranslatef(-3.5 + (k % 8), -3.0 + (k // 8) * 0.9, -9.0)
            L.glRotatef(11.0 * k, 0.0, 1.0, 0.0)
            L.glCallList(lst)
        L.glFlush()

    for _ in range(3):
        frame()
    L.glFinish()
    t0 = time.perf_counter()
    for _ in range(frames):
        frame()
    L.glFinish()
    dt = (time.perf_counter() - t0) / frames
    lib.destroy()
    return dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--calls", type=int, default=64)
    ap.add_argument("--frames", type=int, default=50)
    ap.add_argument("--reference", action="store_true")
    a = ap.parse_args()
    tris = a.calls * 968
    if a.reference:
        dt = drive(load_reference("shipped"), a.calls, max(a.frames // 10, 2))
        mode = "reference (CPU, as-shipped flags)"
    else:
        dt = drive(load_b200(), a.calls, a.frames)
        mode = "replayed call by call" if os.environ.get("MTGL_NO_LIST_RUNS") else "compiled runs"
    print(json.dumps({"workload": f"{a.calls} x glCallList(Suzanne, 968 triangles) per frame, 800x600, lit", "mode": mode,
                      "ms_per_frame": dt * 1e3, "triangles_per_s": tris / dt}))


if __name__ == "__main__":
    main()
