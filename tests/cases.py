"""Scene cases shared by the CPU (oracle vs reference) and GPU (CUDA vs oracle/reference) parity tests.

Each case is (scene, width, height, variant); see tests/scenes/scenes.c for what the variants select.
Sizes are chosen so that the CPU oracle finishes each case in well under a second.
"""

C4_SMALL = 3 | (2 << 8)          # 3x2 Suzannes
C4_SMALL_PHONG = C4_SMALL | (1 << 16)

CASES = (
    [("c1_suzanne", 800, 600, v) for v in range(5)]
    + [("c1_suzanne", 333, 211, 0)]                       # framebuffer not a multiple of the tile / vector width
    + [("c2_cube", 480, 270, 0), ("c2_cube", 1920, 1080, 0)]
    + [("c2_floor", 400, 300, v) for v in range(6)]
    + [("c2_texenv", 320, 240, v) for v in range(5)]
    + [("c3_fill", 256, 144, 4), ("c3_fill", 640, 360, 9), ("c3_fill", 320, 200, 6 | 256)]    # | 256: per-quad texture coordinates
    + [("c4_grid", 480, 270, C4_SMALL), ("c4_grid", 480, 270, C4_SMALL_PHONG), ("c4_grid", 960, 540, 6 | (4 << 8))]
    + [("clipping", 320, 240, v) for v in range(2)]
    + [("primitives", 320, 240, v) for v in (0, 1, 3, 5, 7, 8)]
    + [("zbuffer", 320, 240, v) for v in list(range(8)) + [9, 17]]
    + [("fog", 320, 240, v) for v in range(3)]
    + [("blend", 320, 240, v) for v in range(8)]
    + [("stencil", 300, 200, v) for v in range(4)]
    + [("lighting", 320, 240, v) for v in range(6)]
    + [("displaylist", 320, 240, v) for v in range(2)]
    + [("vbo", 320, 240, v) for v in range(2)]
    + [("validation", 320, 240, 0)]
    + [("scissor", 320, 240, v) for v in range(4)]
    + [("texture_misc", 320, 240, v) for v in range(16)]
    + [("state_churn", 320, 240, 0), ("state_churn", 517, 389, 0)]
    + [("lines", 320, 240, v) for v in (0, 1, 2, 6, 9, 16, 19, 32, 63)]
    + [("wireframe", 400, 300, v) for v in (0, 1, 2, 3, 4, 5, 8)]
    + [("depth_order", 320, 240, v) for v in (0, 1, 2, 3, 5, 6, 8, 9, 10, 11, 12, 16, 31)]
    + [("depth_order", 517, 389, 1), ("depth_order", 1280, 720, 8)]
    + [("cull", 480, 270, v) for v in range(6)]
    + [("vbo_large", 480, 270, 0)]
    + [("pixels", 320, 240, v) for v in (0, 1, 2, 5, 7)] + [("pixels", 517, 389, 3)]
    + [("displaylist_runs", 480, 270, v) for v in range(4)]
    + [("indexed", 480, 360, v) for v in range(7)]          # Suzanne through glDrawElements: u16 / u32 / client indices / strips with u8
    + [("lifetime", 400, 300, v) for v in range(3)]         # objects deleted / redefined while draws that use them are queued
    + [("c4_grid", 480, 270, C4_SMALL | (1 << 17)), ("c4_grid", 960, 540, 6 | (4 << 8) | (1 << 17))]     # one mesh, many draws
)

# Above 4096 pixels the float edge functions of large triangles round (products beyond 2^24, raster.c:299-302): the
# cases the CUDA path must reproduce with the reference's own expression instead of its exact integer forms.
LARGE_CASES = [
    ("c3_fill", 7680, 4320, 2),          # full-screen quads: in-order tiles, k_fill's inexact edge form
    ("c2_cube", 7680, 4320, 0),          # large deferred triangles: k_vis phase 2 / raster_blocks
    ("clipping", 7680, 4320, 0),         # clipped fans, ordered visibility
    ("blend", 7680, 4320, 3),
    ("c3_fill", 4100, 4100, 1),          # just above 4096, width not a multiple of the tile
]


def case_id(c):
    return f"{c[0]}-{c[1]}x{c[2]}-v{c[3]}"
