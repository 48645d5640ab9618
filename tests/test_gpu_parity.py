"""GPU tests (-m gpu): the CUDA path, driven through the gl* front end and the C ABI, against the
CPU oracle on the same scenes, and against the committed golden planes of the reference."""
import hashlib
import json

import numpy as np
import pytest

from cases import CASES, LARGE_CASES, case_id
from mytinygl_b200 import REPO_ROOT
from parity import assert_gate, compare_planes

pytestmark = pytest.mark.gpu
GOLDEN = json.loads((REPO_ROOT / "tests" / "golden" / "planes.json").read_text())


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("case", CASES + LARGE_CASES, ids=case_id)
def test_cuda_matches_oracle(b200, front_oracle, case):
    got = b200.render(*case)
    ref = front_oracle.render(*case)
    stats = compare_planes(ref, got)
    assert_gate(stats, case_id(case))
    assert got[3] == ref[3]                      # sticky GL error code
    # integer-domain results must also match the reference's committed hashes exactly; so must depth, which is a function
    # of the positions alone (no libm on the device between a vertex and its depth)
    assert digest(got[2]) == GOLDEN[case_id(case)]["stencil"]
    assert digest(got[1]) == GOLDEN[case_id(case)]["depth"]


# scenes with in-order states (blending, textured alpha test, colour masks, stencil + blend ...) on triangles of any size
FILL_SCENES = {"c3_fill", "c2_texenv", "blend", "stencil", "fog", "texture_misc", "state_churn", "scissor", "zbuffer", "depth_order",
               "clipping", "primitives", "pixels", "validation", "c2_floor", "c2_cube"}


@pytest.mark.parametrize("mode", ["always", "never"])
@pytest.mark.parametrize("case", [c for c in CASES + LARGE_CASES if c[0] in FILL_SCENES], ids=case_id)
def test_in_order_tile_kernels(b200, front_oracle, case, mode, monkeypatch):
    """In-order tiles have two kernels: k_fill (a thread owns pixels; takes the tiles of large triangles) and
    k_raster<false> (a warp owns a region; small triangles, lines, points).  MTGL_FILL=always sends every eligible tile
    through the first, never through the second: both must reproduce the reference on every in-order case, whatever the
    triangle size."""
    monkeypatch.setenv("MTGL_FILL", mode)
    got = b200.render(*case)
    ref = front_oracle.render(*case)
    assert_gate(compare_planes(ref, got), case_id(case))
    assert digest(got[2]) == GOLDEN[case_id(case)]["stencil"]
    assert digest(got[1]) == GOLDEN[case_id(case)]["depth"]


def test_exactness_report(b200, front_oracle):
    """Not a gate: counts how many cases are bit-identical (libm differences are the only expected source)."""
    exact = 0
    for case in CASES:
        s = compare_planes(front_oracle.render(*case), b200.render(*case))
        exact += int(s["color_diff_pixels"] == 0 and s["depth_diff_pixels"] == 0)
    print(f"bit-identical cases: {exact}/{len(CASES)}")
    assert exact >= len(CASES) // 2


@pytest.mark.parametrize("case,split", [(("c1_suzanne", 800, 600, 0), 320), (("c4_grid", 480, 270, 3 | (2 << 8)), 100),
                                        (("stencil", 300, 200, 0), 64), (("state_churn", 517, 389, 0), 200),
                                        (("cull", 480, 270, 0), 70), (("cull", 480, 270, 1), 135), (("cull", 480, 270, 2), 101),
                                        (("cull", 480, 270, 3), 64), (("cull", 480, 270, 4), 200)])
def test_cuda_bands_stitch_to_full_frame(b200, front_oracle, case, split):
    """mtgl_dev_set_band: two contexts owning complementary (not tile-aligned) bands reproduce the full frame."""
    import ctypes
    name, w, h, variant = case
    b200.lib.mtgl_dev_set_band.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    bands = []
    for y0, y1 in ((0, split), (split, h)):
        b200.create(w, h)
        assert b200.lib.mtgl_dev_set_band(b200.device(), y0, y1) == 0
        b200.lib.glClearColor(ctypes.c_float(0.25), ctypes.c_float(0.5), ctypes.c_float(0.75), ctypes.c_float(1.0))
        b200.lib.glClear(0x4000 | 0x0100 | 0x0400)
        b200.lib.glClearColor(ctypes.c_float(0), ctypes.c_float(0), ctypes.c_float(0), ctypes.c_float(1))
        assert b200.lib.scene_render(name.encode(), w, h, variant) == 0
        planes = b200.read()
        b200.destroy()
        bands.append([p[y0:y1] for p in planes])
    got = tuple(np.concatenate([bands[0][i], bands[1][i]]) for i in range(3))
    assert_gate(compare_planes(front_oracle.render(*case), got), case_id(case))


@pytest.mark.parametrize("case", [("c1_suzanne", 800, 600, 0), ("depth_order", 320, 240, 8), ("c3_fill", 256, 144, 4),
                                  ("c4_grid", 480, 270, 3 | (2 << 8))], ids=case_id)
def test_list_guess_miss_requeues(b200, front_oracle, case, monkeypatch):
    """Optimistic tile-list sizing (mtgl_dev.cu): when the guessed capacity is too small the fill and raster kernels
    must leave everything untouched and the host re-queues them; the frame is the same as with a fitting guess."""
    want = b200.render(*case)
    monkeypatch.setenv("MTGL_LIST_GUESS", "8")
    got = b200.render(*case)                     # a fresh context: its list buffer starts empty
    for a, b in zip(want[:3], got[:3]):
        assert np.array_equal(a, b)
    assert_gate(compare_planes(front_oracle.render(*case), got), case_id(case))


def test_chunk_culling_engages(b200):
    """The culling pass in front of set-up (k_cull.cu) must actually drop chunks: on a band from the first frame, on a
    whole frame from the second draw of an unchanged buffer (the `cull` scene draws twice; in variant 5 most of its grid is
    off screen).
    That the result does not depend on it is what test_cuda_matches_oracle / test_cuda_bands_stitch check."""
    import ctypes

    class Stats(ctypes.Structure):
        _fields_ = [("u64", ctypes.c_uint64 * 5), ("f", ctypes.c_float * 9), ("batches", ctypes.c_uint64), ("d", ctypes.c_double * 9),
                    ("chunks_culled", ctypes.c_uint64)]
    L = b200.lib
    L.mtgl_dev_get_stats.argtypes = [ctypes.c_void_p, ctypes.POINTER(Stats)]
    L.mtgl_dev_set_band.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    st = Stats()
    for band, variant in ((None, 5), ((0, 64), 0)):
        b200.create(480, 270)
        if band:
            assert L.mtgl_dev_set_band(b200.device(), *band) == 0
        assert L.scene_render(b"cull", 480, 270, variant) == 0
        L.glFinish()
        assert L.mtgl_dev_get_stats(b200.device(), ctypes.byref(st)) == 0
        b200.destroy()
        # 8x6 Suzannes = 46464 triangles = 182 chunks; the stats are those of the last batch (the second frame)
        assert st.chunks_culled > 40, (band, st.chunks_culled)


@pytest.mark.parametrize("shape", ["small", "large"])
@pytest.mark.parametrize("case", [("c1_suzanne", 800, 600, 0), ("c4_grid", 960, 540, 6 | (4 << 8)), ("c4_grid", 480, 270, 3 | (2 << 8) | (1 << 16)),
                                  ("depth_order", 517, 389, 1), ("cull", 480, 270, 1), ("c2_cube", 1920, 1080, 0),
                                  ("c3_fill", 640, 360, 9), ("blend", 320, 240, 3), ("stencil", 300, 200, 1), ("c2_texenv", 320, 240, 3)], ids=case_id)
def test_tile_kernel_shapes(b200, front_oracle, case, shape, monkeypatch):
    """The tile kernels come in a throughput shape (k_vis<256>, one k_shade / k_fill CTA per tile: grids of several waves)
    and a latency shape (k_vis<512>, four k_shade / k_fill CTAs per tile: the band of a multi-GPU frame).  Both must give
    the same frame at any grid size."""
    monkeypatch.setenv("MTGL_GRID_SHAPE", shape)
    got = b200.render(*case)
    assert_gate(compare_planes(front_oracle.render(*case), got), case_id(case))
    monkeypatch.setenv("MTGL_GRID_SHAPE", "large" if shape == "small" else "small")
    other = b200.render(*case)
    for a, b in zip(got[:3], other[:3]):
        assert np.array_equal(a, b)


def _present_child(conn, case, mode=0):
    """second process: renders `case` with its colour stores mirrored into the parent's plane"""
    import ctypes
    from mytinygl_b200 import load_b200
    lib = load_b200()
    L = lib.lib
    L.mtgl_dev_set_present_target.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    L.mtgl_dev_set_present_mode.argtypes = [ctypes.c_void_p, ctypes.c_int]
    name, w, h, variant = case
    lib.create(w, h)
    assert L.mtgl_dev_set_present_mode(lib.device(), mode) == 0
    handle = (ctypes.c_ubyte * 64)(*conn.recv())
    rc = L.mtgl_dev_set_present_target(lib.device(), handle)
    if rc == 0:
        L.glClearColor(ctypes.c_float(0.25), ctypes.c_float(0.5), ctypes.c_float(0.75), ctypes.c_float(1.0))
        L.glClear(0x4000 | 0x0100 | 0x0400)
        L.glClearColor(ctypes.c_float(0), ctypes.c_float(0), ctypes.c_float(0), ctypes.c_float(1))
        assert L.scene_render(name.encode(), w, h, variant) == 0
        L.glFlush()
        L.mtgl_dev_frame_barrier.argtypes = [ctypes.c_void_p, ctypes.c_uint]
        assert L.mtgl_dev_frame_barrier(lib.device(), 2) == 0      # "my frame has landed", behind the raster kernels
        color = lib.read()[0]
        conn.send((0, color))
    else:
        conn.send((rc, None))
    conn.recv()             # keep the mapping alive until the parent has read its plane
    lib.destroy()


@pytest.mark.parametrize("mode", [0, 1], ids=["stores", "copy"])
@pytest.mark.parametrize("case", [("c4_grid", 517, 389, 3 | (2 << 8)), ("state_churn", 517, 389, 0), ("lines", 320, 240, 19)], ids=case_id)
def test_present_target_mirrors_color(b200, case, mode):
    """mtgl_dev_export_color_plane / mtgl_dev_set_present_target: every colour store of a context also lands in the
    plane it was given.  Two processes on one GPU here (CUDA IPC needs two processes); bench.py --gpus N uses the same
    calls across GPUs over NVLink and checks the assembled frame against a single-GPU render."""
    import ctypes
    import multiprocessing as mp
    L = b200.lib
    L.mtgl_dev_export_color_plane.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    name, w, h, variant = case
    b200.create(w, h)
    L.mtgl_dev_set_present_mode.argtypes = [ctypes.c_void_p, ctypes.c_int]
    assert L.mtgl_dev_set_present_mode(b200.device(), mode) == 0       # 0: fused peer stores, 1: asynchronous band copy at the barrier
    L.glClearColor(ctypes.c_float(1), ctypes.c_float(0), ctypes.c_float(1), ctypes.c_float(0))
    L.glClear(0x4000)
    L.glFinish()
    handle = (ctypes.c_ubyte * 64)()
    assert L.mtgl_dev_export_color_plane(b200.device(), handle) == 0
    ctx = mp.get_context("spawn")
    parent, child = ctx.Pipe()
    proc = ctx.Process(target=_present_child, args=(child, case, mode))
    proc.start()
    try:
        parent.send(list(handle))
        # the presenter's side of the frame barrier: its stream continues once the other context's frame is complete
        L.mtgl_dev_frame_barrier.argtypes = [ctypes.c_void_p, ctypes.c_uint]
        assert L.mtgl_dev_frame_barrier(b200.device(), 2) == 0
        mine = b200.read()[0]                   # queued behind the barrier: no other synchronisation with the child
        rc, theirs = parent.recv()
        assert rc == 0, f"mtgl_dev_set_present_target failed ({rc})"
        assert np.array_equal(mine, theirs)
    finally:
        parent.send("done")
        proc.join(60)
        b200.destroy()


def test_pipelined_transfers_match_the_synchronous_path(b200):
    """mtglBufferDataPinned / mtglReadColorAsync (include/mtgl_context.h): frames queued back to back, each with its own
    vertex data uploaded into orphaned storage and its colour plane read back asynchronously, give the same pixels as
    glBufferData + draw + synchronous read -- including while earlier frames are still in flight."""
    import ctypes
    L = b200.lib
    w, h, variant = 480, 270, 3 | (2 << 8)
    L.scene_c4_host_data.restype = ctypes.c_void_p
    L.scene_c4_vbo.restype = ctypes.c_uint
    L.glBindBuffer.argtypes = [ctypes.c_uint, ctypes.c_uint]
    L.glBufferData.argtypes = [ctypes.c_uint, ctypes.c_long, ctypes.c_void_p, ctypes.c_uint]
    L.mtglBufferDataPinned.argtypes = [ctypes.c_uint, ctypes.c_long, ctypes.c_void_p, ctypes.c_uint]
    L.mtglReadColorAsync.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    b200.create(w, h)
    L.scene_c4_setup(w, h, variant)
    nv = L.scene_c4_vertex_count()
    base = np.ctypeslib.as_array(ctypes.cast(L.scene_c4_host_data(), ctypes.POINTER(ctypes.c_float)), shape=(nv, 8)).copy()
    frames = []
    for k in range(5):                       # a different mesh every frame: shifted and squeezed
        f = base.copy()
        f[:, 0] = f[:, 0] * (1.0 - 0.07 * k) + 0.11 * k
        f[:, 1] += 0.05 * k
        frames.append(np.ascontiguousarray(f))
    vbo = L.scene_c4_vbo()
    want = []
    for f in frames:                         # synchronous path
        L.glBindBuffer(0x8892, vbo)
        L.glBufferData(0x8892, f.nbytes, f.ctypes.data, 0x88E4)
        L.scene_c4_draw()
        want.append(b200.read()[0].copy())
    outs = [np.zeros((h, w), np.uint32) for _ in frames]
    for f, o in zip(frames, outs):           # pipelined path: nothing waits until glFinish
        L.glBindBuffer(0x8892, vbo)
        L.mtglBufferDataPinned(0x8892, f.nbytes, f.ctypes.data, 0x88E4)
        L.scene_c4_draw()
        L.mtglReadColorAsync(0, h, o.ctypes.data)
    L.glFinish()
    assert L.glGetError() == 0
    for k, (a, o) in enumerate(zip(want, outs)):
        assert np.array_equal(a, o), f"frame {k}"
    assert not np.array_equal(want[0], want[4])
    b200.destroy()


def test_frame_barrier_times_out_without_poisoning_the_context(b200, front_oracle, monkeypatch):
    """mtgl_dev_frame_barrier with a participant that never arrives: after MTGL_BARRIER_TIMEOUT_S the barrier kernel gives
    up, mtgl_dev_finish reports it (glFinish: GL_INVALID_OPERATION) and the context keeps rendering correctly -- the
    kernel used to trap, which left a sticky CUDA error behind on every rank."""
    import ctypes
    monkeypatch.setenv("MTGL_BARRIER_TIMEOUT_S", "0.2")
    L = b200.lib
    L.mtgl_dev_frame_barrier.argtypes = [ctypes.c_void_p, ctypes.c_uint]
    name, w, h, variant = "c1_suzanne", 320, 240, 0

    def twice(lib, between):
        """the scene, `between`, then the planes cleared and the scene again in the SAME context"""
        lib.create(w, h)
        assert lib.lib.scene_render(name.encode(), w, h, variant) == 0
        between(lib)
        lib.lib.glClearColor(ctypes.c_float(0.25), ctypes.c_float(0.5), ctypes.c_float(0.75), ctypes.c_float(1.0))
        lib.lib.glClearStencil(0)
        lib.lib.glClear(0x4000 | 0x0100 | 0x0400)
        assert lib.lib.scene_render(name.encode(), w, h, variant) == 0
        out = lib.read()
        assert lib.lib.glGetError() == 0
        lib.destroy()
        return out

    def stuck_barrier(lib):
        assert L.mtgl_dev_frame_barrier(lib.device(), 2) == 0     # two participants, one process: nobody else will come
        L.glFinish()
        assert L.glGetError() == 0x0502                           # GL_INVALID_OPERATION: the finish saw the timeout
        assert L.glGetError() == 0

    got = twice(b200, stuck_barrier)                              # the context is alive and exact afterwards
    want = twice(front_oracle, lambda lib: lib.lib.glFinish())
    for a, b, plane in zip(got[:3], want[:3], ("color", "depth", "stencil")):
        assert np.array_equal(a, b), plane


def test_orphaned_buffer_filled_on_a_side_stream(b200):
    """mtgl_context_buffer_orphan (include/mtgl_context.h): the buffer name gets fresh storage every frame, the application
    fills it on a stream of its own (here: a device-to-device copy, in bench.py a host slice + NCCL all-gather) while earlier
    frames are still in flight, and orders the context's stream behind the fill with an event.  Same pixels as glBufferData."""
    import ctypes
    import torch
    L = b200.lib
    w, h, variant = 480, 270, 3 | (2 << 8)
    L.scene_c4_host_data.restype = ctypes.c_void_p
    L.scene_c4_vbo.restype = ctypes.c_uint
    L.glBindBuffer.argtypes = [ctypes.c_uint, ctypes.c_uint]
    L.glBufferData.argtypes = [ctypes.c_uint, ctypes.c_long, ctypes.c_void_p, ctypes.c_uint]
    L.mtglReadColorAsync.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    L.gl_get_current_context.restype = ctypes.c_void_p
    L.mtgl_context_buffer_orphan.argtypes = [ctypes.c_void_p, ctypes.c_uint, ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_uint64)]
    L.mtgl_dev_stream.restype = ctypes.c_void_p
    L.mtgl_dev_stream.argtypes = [ctypes.c_void_p]
    b200.create(w, h)
    L.scene_c4_setup(w, h, variant)
    nv = L.scene_c4_vertex_count()
    base = np.ctypeslib.as_array(ctypes.cast(L.scene_c4_host_data(), ctypes.POINTER(ctypes.c_float)), shape=(nv, 8)).copy()
    frames = []
    for k in range(5):
        f = base.copy()
        f[:, 0] = f[:, 0] * (1.0 - 0.06 * k) + 0.09 * k
        f[:, 1] -= 0.04 * k
        frames.append(np.ascontiguousarray(f))
    vbo = L.scene_c4_vbo()
    want = []
    for f in frames:
        L.glBindBuffer(0x8892, vbo)
        L.glBufferData(0x8892, f.nbytes, f.ctypes.data, 0x88E4)
        L.scene_c4_draw()
        want.append(b200.read()[0].copy())

    class DevTensor:
        def __init__(self, ptr, nbytes):
            self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 3}
    staged = [torch.from_numpy(f.view(np.uint8).reshape(-1)).cuda() for f in frames]
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    lib_stream = torch.cuda.ExternalStream(L.mtgl_dev_stream(b200.device()))
    ctx = L.gl_get_current_context()
    outs = [np.zeros((h, w), np.uint32) for _ in frames]
    seen = set()
    for k, (f, src, o) in enumerate(zip(frames, staged, outs)):
        ptr, size = ctypes.c_void_p(), ctypes.c_uint64()
        hint = f.ctypes.data if k % 2 == 0 else None        # with and without the host copy of the contents to come
        assert L.mtgl_context_buffer_orphan(ctx, vbo, hint, ctypes.byref(ptr), ctypes.byref(size)) == 0 and size.value == f.nbytes
        seen.add(ptr.value)
        dst = torch.as_tensor(DevTensor(ptr.value, f.nbytes), device="cuda")
        ev = torch.cuda.Event()
        with torch.cuda.stream(side):
            dst.copy_(src, non_blocking=True)
            ev.record(side)
        lib_stream.wait_event(ev)
        L.scene_c4_draw()
        L.mtglReadColorAsync(0, h, o.ctypes.data)
    L.glFinish()
    assert L.glGetError() == 0
    assert len(seen) >= 1
    for k, (a, o) in enumerate(zip(want, outs)):
        assert np.array_equal(a, o), f"frame {k}"
    b200.destroy()


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 6])
def test_indexed_draws_share_vertices(b200, front_oracle, variant, monkeypatch):
    """glDrawElements (gl_api.c:1854-1941) re-emits a vertex per index; the back end runs the vertex stage once per buffer
    ELEMENT when the arrays hold fewer elements than the draw has indices (Suzanne: 507 + 1 default slot instead of 2904)
    and set-up looks vertices up by index.  Same frame as the per-index path, and the shared path must actually engage."""
    import ctypes

    class Stats(ctypes.Structure):
        _fields_ = [("vertices", ctypes.c_uint64), ("rest", ctypes.c_uint64 * 40)]
    L = b200.lib
    L.mtgl_dev_get_stats.argtypes = [ctypes.c_void_p, ctypes.POINTER(Stats)]
    case = ("indexed", 480, 360, variant)
    seen = {}
    for mode in ("shared", "per-index"):
        if mode == "per-index":
            monkeypatch.setenv("MTGL_NO_SHARED_VERTS", "1")
        b200.create(480, 360)
        L.glClearColor(ctypes.c_float(0.25), ctypes.c_float(0.5), ctypes.c_float(0.75), ctypes.c_float(1.0))
        L.glClear(0x4000 | 0x0100 | 0x0400)
        assert L.scene_render(b"indexed", 480, 360, variant) == 0
        planes = b200.read()
        st = Stats()
        assert L.mtgl_dev_get_stats(b200.device(), ctypes.byref(st)) == 0
        b200.destroy()
        seen[mode] = (planes, int(st.vertices))
    for a, b in zip(seen["shared"][0], seen["per-index"][0]):
        assert np.array_equal(a, b)
    draws = 3 if variant == 3 else 1
    assert seen["shared"][1] == draws * 508 and seen["per-index"][1] == draws * 2904, (seen["shared"][1], seen["per-index"][1])
    assert_gate(compare_planes(front_oracle.render(*case), b200.render(*case)), case_id(case))
