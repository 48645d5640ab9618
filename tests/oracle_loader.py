"""Loaders of the two CHECKERS (test infrastructure, never part of the product path):

* ``load_front_oracle()``  oracle/_build/libfront_oracle.so: gl* front end + the CPU restatement oracle/mtgl_oracle.c
* ``load_reference(kind)`` oracle/_ref/libref_{strict,shipped,shipped_v3}.so: the UNMODIFIED reference compiled from
                           /root/reference by `make ref` (strict = canonical IEEE build, parity oracle; shipped = the
                           reference's own flags, timing baseline)

Used by tests/, tools/, __graft_entry__.smoke() and the reference / cpu_baseline / parity legs of bench.py.
"""
from __future__ import annotations

import os
import subprocess
import sys
from pathlib import Path

from mytinygl_b200.loader import REPO_ROOT, SceneLibrary


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
