"""CPU test: the gl* front end's queryable state against the unmodified reference.

After each of a dozen scene recipes every token of include/GL/gl.h is pushed through glGetFloatv / glGetIntegerv /
glGetBooleanv / glGetDoublev / glIsEnabled, every (light, token) pair through glGetLightfv, every (face, token) pair
through glGetMaterialfv, and a range of names through glIsTexture / glIsBuffer / glIsList -- on the front end
(oracle/_build/libfront_oracle.so shares its front-end objects with the product library) and on the reference's strict
build (oracle/_ref/libref_strict.so).  The sticky error each query raises must agree, and every error-free query must
return the same bytes.  Reference: src/gl_api.c (glGet* 1426-1743, glIsEnabled 459-488, glGetLightfv / glGetMaterialfv).
"""
import ctypes
import re

import numpy as np
import pytest

from mytinygl_b200 import REPO_ROOT

SCENES = [("c1_suzanne", 320, 240, 3), ("c2_cube", 320, 240, 0), ("c2_texenv", 320, 240, 3), ("c3_fill", 128, 96, 2),
          ("c4_grid", 240, 160, 2 | (1 << 8)), ("fog", 160, 120, 2), ("blend", 160, 120, 5), ("stencil", 150, 100, 2),
          ("lighting", 160, 120, 4), ("scissor", 160, 120, 1), ("state_churn", 160, 120, 0), ("texture_misc", 160, 120, 7),
          ("displaylist", 160, 120, 1), ("vbo", 160, 120, 1), ("pixels", 160, 120, 2), ("wireframe", 200, 150, 3)]


def gl_tokens():
    text = (REPO_ROOT / "include" / "GL" / "gl.h").read_text()
    toks = {}
    for name, val in re.findall(r"#define\s+(GL_[A-Z0-9_]+)\s+(0x[0-9A-Fa-f]+|\d+)\b", text):
        toks.setdefault(int(val, 0), name)
    return sorted(toks.items())


def snapshot(lib, tokens):
    """Everything the query entry points return in the current context: {query: (error, bytes)}."""
    L = lib.lib
    L.glGetError.restype = ctypes.c_uint
    L.glIsEnabled.restype = ctypes.c_ubyte
    for f in (L.glIsTexture, L.glIsBuffer, L.glIsList):
        f.restype = ctypes.c_ubyte
        f.argtypes = [ctypes.c_uint]
    out = {}
    L.glGetError()
    getters = (("f", L.glGetFloatv, np.float32), ("i", L.glGetIntegerv, np.int32), ("b", L.glGetBooleanv, np.uint8), ("d", L.glGetDoublev, np.float64))
    for val, name in tokens:
        for tag, fn, dt in getters:
            buf = np.full(32, 0x5A, dtype=np.uint8).view(np.uint8)
            arr = np.frombuffer(np.full(32 * 8, 0x5A, dtype=np.uint8).tobytes(), dtype=dt).copy()
            fn(ctypes.c_uint(val), arr.ctypes.data_as(ctypes.c_void_p))
            out[(tag, name)] = (L.glGetError(), arr.tobytes())
            del buf
        r = L.glIsEnabled(ctypes.c_uint(val))
        out[("e", name)] = (L.glGetError(), bytes([r]))
    lights = [v for v, n in tokens if n.startswith("GL_LIGHT") and n[8:].isdigit()]
    for lv in lights + [0x4000 + 8, 0]:
        for val, name in tokens:
            arr = np.frombuffer(np.full(16 * 4, 0x5A, dtype=np.uint8).tobytes(), dtype=np.float32).copy()
            L.glGetLightfv(ctypes.c_uint(lv), ctypes.c_uint(val), arr.ctypes.data_as(ctypes.c_void_p))
            out[("l", lv, name)] = (L.glGetError(), arr.tobytes())
    for face in (0x0404, 0x0405, 0x0408, 0):
        for val, name in tokens:
            arr = np.frombuffer(np.full(16 * 4, 0x5A, dtype=np.uint8).tobytes(), dtype=np.float32).copy()
            L.glGetMaterialfv(ctypes.c_uint(face), ctypes.c_uint(val), arr.ctypes.data_as(ctypes.c_void_p))
            out[("m", face, name)] = (L.glGetError(), arr.tobytes())
    for i in list(range(0, 12)) + [255, 256, 257, 1024, 1025, 0xFFFFFFFF]:
        out[("t", i)] = (0, bytes([L.glIsTexture(i)]))
        out[("u", i)] = (0, bytes([L.glIsBuffer(i)]))
        out[("s", i)] = (0, bytes([L.glIsList(i)]))
    return out


@pytest.mark.parametrize("scene", SCENES, ids=lambda s: f"{s[0]}-v{s[3]}")
def test_queryable_state_matches_the_reference(front_oracle, ref_strict, scene):
    tokens = gl_tokens()
    assert len(tokens) > 300
    snaps = []
    for lib in (ref_strict, front_oracle):
        name, w, h, variant = scene
        lib.create(w, h)
        assert lib.lib.scene_render(name.encode(), w, h, variant) == 0
        lib.lib.glFinish()
        snaps.append(snapshot(lib, tokens))
        lib.destroy()
    want, got = snaps
    assert want.keys() == got.keys()
    compared = 0
    for k in want:
        assert want[k][0] == got[k][0], (k, "sticky error", hex(want[k][0]), hex(got[k][0]))
        if want[k][0] == 0:
            assert want[k][1] == got[k][1], (k, want[k][1][:32], got[k][1][:32])
            compared += 1
    assert compared > 400          # the error-free queries: the valid state tokens, lights, materials, names
