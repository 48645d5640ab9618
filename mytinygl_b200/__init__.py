"""mytinygl_b200 -- B200-native (sm_100a) rendering back end behind MyTinyGL's gl* API.

The product is the C library ``mytinygl_b200/lib/libMyTinyGL_b200.so`` (gl* front end in C++ plus the
hand-written CUDA back end behind the C ABI of ``include/mtgl_dev.h``).  This Python package is only
the thin ctypes loader that the tests and ``bench.py`` use to drive it; it contains no rendering
code and no CPU fallback: if the CUDA library is missing, loading fails loudly.
"""
from .loader import (  # noqa: F401
    REPO_ROOT,
    LibraryMissing,
    SceneLibrary,
    build,
    load_b200,
    load_suzanne,
)

__all__ = [
    "REPO_ROOT",
    "LibraryMissing",
    "SceneLibrary",
    "build",
    "load_b200",
    "load_suzanne",
]
