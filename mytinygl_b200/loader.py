"""ctypes loader for the three scene libraries that share one entry-point set.

* ``load_b200()``          tests/scenes/_build/libscenes_b200.so -> mytinygl_b200/lib/libMyTinyGL_b200.so
                           (the product: CUDA back end; raises LibraryMissing when it was not built)
* ``load_front_oracle()``  oracle/_build/libfront_oracle.so (front end + CPU restatement; tests only)
* ``load_reference(kind)`` oracle/_ref/libref_{strict,shipped,shipped_v3}.so (the unmodified reference
                           compiled from /root/reference; tests and the CPU baseline only)

All of them export scene_set_mesh / scene_render / scene_c4_* (tests/scenes/scenes.c), the four
mtgl_harness_* functions and the public gl* API.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys
from pathlib import Path

import numpy as np

REPO_ROOT = Path(__file__).resolve().parent.parent


class LibraryMissing(RuntimeError):
    """Raised when a required native library has not been built.  Never papered over."""


def build(targets=("product", "oracle", "ref"), jobs: int = 8) -> None:
    """Run the top-level Makefile (nvcc cross-compiles sm_100a without a GPU)."""
    subprocess.run(["make", f"-j{jobs}", *targets], cwd=REPO_ROOT, check=True)


def load_suzanne():
    """The Suzanne mesh fixture (tests/golden/suzanne.npz)."""
    z = np.load(REPO_ROOT / "tests" / "golden" / "suzanne.npz")
    return (np.ascontiguousarray(z["pos"], dtype=np.float32), np.ascontiguousarray(z["nrm"], dtype=np.float32),
            np.ascontiguousarray(z["faces"], dtype=np.int32))


class SceneLibrary:
    """One loaded library + a current GL context created through its harness."""

    def __init__(self, path: Path, kind: str):
        if not Path(path).exists():
            raise LibraryMissing(f"{path} is missing: run `make` (or __graft_entry__.build()) first")
        self.path = Path(path)
        self.kind = kind
        self.lib = ctypes.CDLL(str(path), mode=ctypes.RTLD_LOCAL)
        L = self.lib
        L.mtgl_harness_create.restype = ctypes.c_void_p
        L.mtgl_harness_create.argtypes = [ctypes.c_int, ctypes.c_int]
        L.mtgl_harness_destroy.argtypes = [ctypes.c_void_p]
        L.mtgl_harness_make_current.argtypes = [ctypes.c_void_p]
        L.mtgl_harness_read.restype = ctypes.c_int
        L.mtgl_harness_read.argtypes = [ctypes.c_void_p] * 4
        L.mtgl_harness_kind.restype = ctypes.c_char_p
        L.scene_set_mesh.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        L.scene_render.restype = ctypes.c_int
        L.scene_render.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.scene_c4_setup.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int]
        L.scene_c4_vertex_count.restype = ctypes.c_int
        L.scene_count.restype = ctypes.c_int
        L.scene_name.restype = ctypes.c_char_p
        L.scene_name.argtypes = [ctypes.c_int]
        L.glGetError.restype = ctypes.c_uint
        if hasattr(L, "mtgl_harness_device"):
            L.mtgl_harness_device.restype = ctypes.c_void_p
            L.mtgl_harness_device.argtypes = [ctypes.c_void_p]
        pos, nrm, faces = load_suzanne()
        self._mesh = (pos, nrm, faces)
        L.scene_set_mesh(pos.ctypes.data, nrm.ctypes.data, faces.ctypes.data, pos.shape[0], faces.shape[0])
        self.ctx = None
        self.size = None

    # -- contexts ------------------------------------------------------------------------------
    def create(self, w: int, h: int):
        self.destroy()
        ctx = self.lib.mtgl_harness_create(w, h)
        if not ctx:
            raise RuntimeError(f"{self.kind}: gl_create_context({w}, {h}) failed"
                               + (" (no CUDA device? there is no CPU fallback)" if self.kind == "b200" else ""))
        self.ctx, self.size = ctx, (w, h)
        return ctx

    def destroy(self):
        if self.ctx:
            self.lib.mtgl_harness_destroy(self.ctx)
            self.ctx = None

    def read(self, color=True, depth=True, stencil=True):
        w, h = self.size
        c = np.empty((h, w), np.uint32) if color else None
        d = np.empty((h, w), np.float32) if depth else None
        s = np.empty((h, w), np.uint8) if stencil else None
        rc = self.lib.mtgl_harness_read(self.ctx, c.ctypes.data if color else None, d.ctypes.data if depth else None,
                                        s.ctypes.data if stencil else None)
        if rc != 0:
            raise RuntimeError(f"{self.kind}: framebuffer read failed")
        return c, d, s

    def device(self):
        return self.lib.mtgl_harness_device(self.ctx)

    # -- scenes --------------------------------------------------------------------------------
    def scene_names(self):
        return [self.lib.scene_name(i).decode() for i in range(self.lib.scene_count())]

    def render(self, name: str, w: int, h: int, variant: int = 0, prefill=True):
        """Fresh context, (optionally) defined framebuffer contents, one scene, planes back."""
        self.create(w, h)
        if prefill:
            # gl_create_context leaves the planes uninitialised (framebuffer.h:44-46): define them
            self.lib.glClearColor(ctypes.c_float(0.25), ctypes.c_float(0.5), ctypes.c_float(0.75), ctypes.c_float(1.0))
            self.lib.glClearStencil(0)
            self.lib.glClear(0x4000 | 0x0100 | 0x0400)
            self.lib.glClearColor(ctypes.c_float(0), ctypes.c_float(0), ctypes.c_float(0), ctypes.c_float(1))
        if self.lib.scene_render(name.encode(), w, h, variant) != 0:
            raise KeyError(name)
        out = self.read()
        err = self.lib.glGetError()
        self.destroy()
        return out + (err,)


def load_b200() -> SceneLibrary:
    return SceneLibrary(REPO_ROOT / "tests" / "scenes" / "_build" / "libscenes_b200.so", "b200")


def load_front_oracle() -> SceneLibrary:
    return SceneLibrary(REPO_ROOT / "oracle" / "_build" / "libfront_oracle.so", "front+oracle")


def _runs_here(path: Path) -> bool:
    """True if the library executes on this host (a -march=native build may not)."""
    code = ("import ctypes,sys; l=ctypes.CDLL(sys.argv[1]); l.mtgl_harness_create.restype=ctypes.c_void_p;"
            "c=l.mtgl_harness_create(64,64); l.scene_render.argtypes=[ctypes.c_char_p]+[ctypes.c_int]*3;"
            "l.scene_render(b'c2_cube',64,64,0)")
    try:
        return subprocess.run([sys.executable, "-c", code, str(path)], timeout=120,
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL).returncode == 0
    except Exception:
        return False


def load_reference(kind: str = "strict") -> SceneLibrary:
    """kind: 'strict' (parity oracle) or 'shipped' (timing baseline: the reference's own flags)."""
    base = REPO_ROOT / "oracle" / "_ref"
    if kind == "strict":
        return SceneLibrary(base / "libref_strict.so", "ref-strict")
    native = base / "libref_shipped.so"
    if native.exists() and os.environ.get("MTGL_REF_PORTABLE") != "1" and _runs_here(native):
        return SceneLibrary(native, "ref-shipped(-march=native)")
    return SceneLibrary(base / "libref_shipped_v3.so", "ref-shipped(-march=x86-64-v3)")
