"""CPU tests: the shipped C-ABI library loads and exports every symbol the public headers declare
(no compute calls -- there is no GPU here), and refuses to create a context without a device."""
import ctypes
import re

import pytest

from mytinygl_b200 import REPO_ROOT

LIB = REPO_ROOT / "mytinygl_b200" / "lib" / "libMyTinyGL_b200.so"


def declared(header, pattern):
    text = (REPO_ROOT / "include" / header).read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(pattern, text)))


@pytest.fixture(scope="module")
def lib():
    assert LIB.exists(), "run `make product` (or __graft_entry__.build()) first"
    return ctypes.CDLL(str(LIB))


def test_exports_device_abi(lib):
    names = declared("mtgl_dev.h", r"\b(mtgl_dev_[a-z_0-9]+)\s*\(")
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), n
    lib.mtgl_dev_abi_version.restype = ctypes.c_int
    assert lib.mtgl_dev_abi_version() == 6


def test_exports_gl_api(lib):
    names = declared("GL/gl.h", r"\b(gl[A-Z][A-Za-z0-9]*)\s*\(")
    assert len(names) == 111          # the reference's public entry points (include/GL/gl.h:531-669)
    for n in names:
        assert hasattr(lib, n), n
    for n in declared("mtgl_context.h", r"\b((?:gl_|mtgl_)[a-z_]+|mtgl[A-Z][A-Za-z]+)\s*\("):
        assert hasattr(lib, n), n


def test_no_cpu_fallback(lib):
    """Without a CUDA device gl_create_context must fail (NULL), never render on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib.gl_create_context.restype = ctypes.c_void_p
    assert lib.gl_create_context(64, 64) is None
    out = ctypes.c_void_p()
    lib.mtgl_dev_create.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]
    assert lib.mtgl_dev_create(64, 64, -1, ctypes.byref(out)) == -1     # MTGL_E_NO_DEVICE


def test_product_does_not_link_the_oracle():
    """The oracle is test infrastructure: none of its objects may be inside the product library."""
    import subprocess
    syms = subprocess.run(["nm", "-D", "--defined-only", str(LIB)], capture_output=True, text=True).stdout
    assert "scene_render" not in syms          # scenes live in the test-only libscenes_b200.so
    strings = subprocess.run(["strings", str(LIB)], capture_output=True, text=True).stdout
    assert "mtgl_oracle" not in strings
