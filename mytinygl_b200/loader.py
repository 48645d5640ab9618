"""ctypes loader of the product library.

``load_b200()`` loads tests/scenes/_build/libscenes_b200.so -> mytinygl_b200/lib/libMyTinyGL_b200.so (the CUDA back
end behind the gl* API; raises LibraryMissing when it was not built -- there is no fallback).

``SceneLibrary`` is the generic wrapper around any library that exports scene_set_mesh / scene_render / scene_c4_*
(tests/scenes/scenes.c), the four mtgl_harness_* functions and the public gl* API.  The checkers -- the CPU oracle and
the unmodified reference -- are loaded with the same class by tests/oracle_loader.py; nothing in this package knows
where they live.
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
        L.scene_c3_setup.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int]
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
