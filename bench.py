#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native MyTinyGL back end.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c4|c3|c5|c4i|c3s]
                    [--no-secondary] [--no-parity] [--no-cpu-baseline]

A "step" is one frame of the workload: glClear + the draw calls + whatever makes the result observable.
Headline workload: C4 of BASELINE.json -- 38x28 Suzannes (1 029 952 triangles, 3 089 856 vertices), 8 lights,
trilinear 64x64 texture, one glDrawArrays from a VBO, 3840x2160.

Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  value        covered fragments/s with the VBO and texture resident in HBM (device-timed, K frames)
  e2e          the same metric with the per-frame host->device upload of the 98.9 MB vertex buffer from pinned host
               memory and the device->host read-back of the colour plane inside the timed region, through the public
               gl* API + the C ABI, pipelined across frames (mtglBufferDataPinned / mtglReadColorAsync at N = 1; at
               N > 1 a sharded upload + NCCL all-gather into orphaned storage on a side stream and every rank's band
               read back into one shared page-locked frame); --serial-e2e keeps the strictly sequential form.  The region
               is K steps + glFinish, run three times after four untimed steps; the median run is reported
               (`runs_ms_per_step` lists all three)
  roofline     the step's dominant kernel (whichever the live per-group CUDA-event times say): its algorithmic bytes
               per launch / its CUDA-event duration, against MEASURED_PEAKS.json's HBM copy bandwidth
  cpu_baseline the unmodified reference (as-shipped flags, 1 thread) on this box's host CPU
  stages_ms    CUDA-event time per pipeline stage, mean over the timed steps
  parity       the headline frame rendered once more by the unmodified reference (strict IEEE build,
               oracle/_ref/libref_strict.so) on this box and compared with the CUDA frame: the checker leg
  secondary    the other two BASELINE workloads measured the same way in the same run, at every N:
               c3_fill_3840x2160 (64 full-screen quads, alpha + stencil + blend) and c5_grid_7680x4320 (C4 at 8K),
               each with ms_per_step, value, roofline, parity and -- for N > 1 -- gather_check

Multi-GPU (torchrun, one rank per GPU): sort-first bands of framebuffer rows.  Every rank receives the whole command
stream; a culling pass drops the 256-triangle chunks that cannot reach its band before set-up touches them, and it
bins / rasterises only its band.  The raster kernels store every finished tile both locally and -- over NVLink peer
memory -- into rank 0's colour plane (fused gather), or the band follows the frame as one asynchronous peer copy on a
side stream (--gather copy; 'auto' picks it when more than 64 MB per frame converge on the presenting GPU); a
frame-barrier kernel per rank ends the frame.  NCCL bootstraps the group (and all-gathers the sharded vertex-buffer
upload of the end-to-end step).  Total work is fixed: "scaling": "strong".
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from mytinygl_b200 import load_b200  # noqa: E402

GL_ARRAY_BUFFER = 0x8892
GL_STATIC_DRAW = 0x88E4

WORKLOADS = {
    # name: (counts key, width, height, c4 variant / c3 quads)
    "c4": ("c4_grid_3840x2160", 3840, 2160, 0),
    "c5": ("c5_grid_7680x4320", 7680, 4320, 0),
    "c3": ("c3_fill_3840x2160", 3840, 2160, 64),
    # diagnostic: C3 with per-quad texture coordinates (no two layers sample the same texels: k_fill's coincident-run reuse
    # covers coverage and barycentrics only)
    "c3s": ("c3_shifted_3840x2160", 3840, 2160, 64 | 256),
    # SURVEY.md 8(d) secondary layout: ONE Suzanne in the VBO, 1064 glDrawArrays under glPushMatrix / glTranslatef
    "c4i": ("c4_instanced_3840x2160", 3840, 2160, 1 << 17),
}
SCENE_OF = {"c4": "c4_grid", "c5": "c4_grid", "c4i": "c4_grid", "c3": "c3_fill", "c3s": "c3_fill"}


class Stats(ctypes.Structure):
    _fields_ = [("vertices", ctypes.c_uint64), ("triangles_in", ctypes.c_uint64), ("triangles_setup", ctypes.c_uint64),
                ("tile_refs", ctypes.c_uint64), ("kernel_launches", ctypes.c_uint64), ("last_batch_ms", ctypes.c_float),
                ("stage_ms", ctypes.c_float * 5), ("raster_ms", ctypes.c_float * 3),
                ("batches", ctypes.c_uint64), ("cum_batch_ms", ctypes.c_double), ("cum_stage_ms", ctypes.c_double * 5),
                ("cum_raster_ms", ctypes.c_double * 3), ("chunks_culled", ctypes.c_uint64)]


def counts_for(key):
    return json.loads((ROOT / "tests" / "golden" / "workload_counts.json").read_text())[key]


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def workload_config(workload, world):
    """The `config` object: what the workload is.  Both arms (this back end and --impl reference) print exactly this."""
    key, w, h, _ = WORKLOADS[workload]
    cnt = counts_for(key)
    return {"workload": key, "width": w, "height": h, "triangles": cnt["vertices"] // 3 if cnt["vertices"] else (128 if workload in ("c3", "c3s") else 0),
            "covered_fragments": cnt["covered"], "depth_passing_fragments": cnt["tested"], "shaded_fragments": cnt["shaded"],
            "partition": f"sort-first bands x{world}",
            "l2": "flushed between frames (256 MiB fill)" if workload in ("c3", "c3s") else "inputs larger than L2 (>330 MB streamed per frame)"}


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True).start()
        except OSError:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ reference arm
def load_reference(kind):
    """The unmodified reference compiled by `make ref` (oracle/_ref/).  Only the reference arm, the cpu_baseline leg and
    the parity leg of this file come here; the loader lives with the tests, not in the product package."""
    from oracle_loader import load_reference as _load
    return _load(kind)


def run_reference(args, workload, steps=None, warmup=None):
    """The reference's own CPU implementation, unmodified, as-shipped flags, one thread (it has no threading)."""
    key, w, h, variant = WORKLOADS[workload]
    cnt = counts_for(key)
    lib = load_reference("shipped")
    lib.create(w, h)
    steps = max(1, args.steps if steps is None else steps)
    warm = max(0, args.warmup if warmup is None else warmup)
    times = []
    if workload in ("c3", "c3s"):
        # one C3 frame takes ~40 s on a host core: a step is a bounded sample -- the first 8 of the 64 quads (the state
        # mix and the per-fragment work are the same for every quad), scaled to the frame by fragment count
        sample_quads = 8
        frac = sample_quads / 64.0
        lib.lib.scene_c3_setup(w, h, sample_quads | (variant & 256))
        for i in range(min(warm, 1) + min(steps, 3)):
            t0 = time.perf_counter()
            lib.lib.scene_c3_draw()
            lib.lib.glFinish()
            if i >= min(warm, 1):
                times.append((time.perf_counter() - t0) / frac)
        sample = f"{len(times)} steps of {sample_quads} of the 64 full-screen quads each, time scaled by 64/{sample_quads}"
    else:
        lib.lib.scene_c4_setup(w, h, variant)
        for i in range(warm + steps):
            t0 = time.perf_counter()
            lib.lib.scene_c4_draw()
            lib.lib.glFinish()
            if i >= warm:
                times.append(time.perf_counter() - t0)
        sample = f"{len(times)} full frames of the workload, VBO resident in host memory"
    lib.destroy()
    t = float(np.mean(times))
    value = cnt["covered"] / t
    return {
        "impl": "reference", "metric": "covered_fragments_per_s", "value": value, "unit": "fragments/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "frames_per_s": 1.0 / t,
        "config": workload_config(workload, args.gpus),
        "cpu_baseline": {"value": value, "unit": "fragments/s", "cores": 1, "kind": "reference", "sample": sample,
                         "build": lib.kind, "host_cpus": os.cpu_count()},
        "e2e": {"value": value, "unit": "fragments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }


# ------------------------------------------------------------------------------------------ B200 arm
def band_rows(height, rank, world):
    """Contiguous bands of 64-row tile rows, as even as possible."""
    tile_rows = (height + 63) // 64
    lo = (tile_rows * rank) // world
    hi = (tile_rows * (rank + 1)) // world
    return min(lo * 64, height), min(hi * 64, height)


def timed_e2e(step, finish, steps, dist, torch, local):
    """The end-to-end region: four untimed steps (the orphan pool and the pinned targets reach their steady state), then
    K pipelined steps + finish, three times over; the MEDIAN run is reported (one host hiccup -- an allocation, a page
    fault -- in a 40 ms region otherwise decides the figure).  Every rank takes the same run: the one whose slowest rank is
    the median.  Returns (seconds of the reported run, [seconds of each run, max over ranks])."""
    for k in range(4):
        step(k)
    finish()
    runs = []
    for _ in range(3):
        t0 = time.perf_counter()
        for k in range(steps):
            step(k)
        finish()
        t = time.perf_counter() - t0
        if dist:
            tt = torch.tensor([t], device=f"cuda:{local}")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t = float(tt.item())
        runs.append(t)
    return sorted(runs)[1], runs


def nbytes_ok(nbytes, parts):
    """True when a buffer of nbytes splits into `parts` equal 16-byte-aligned pieces."""
    return nbytes > 0 and nbytes % (parts * 16) == 0


def rebalanced(bounds, times, height, quantum=16):
    """New band boundaries from the ranks' last frame times: cost density taken as uniform inside each current band,
    boundaries moved (half way, for stability) to where the cumulative cost reaches k/N of the total, snapped to
    `quantum` rows.  Every rank computes the same result from the same gathered times."""
    n = len(times)
    dens = [t / max(b1 - b0, 1) for t, b0, b1 in zip(times, bounds[:-1], bounds[1:])]
    total = sum(times)
    out = [0]
    r, acc = 0, 0.0
    for k in range(1, n):
        want = total * k / n
        while r < n - 1 and acc + times[r] < want:
            acc += times[r]
            r += 1
        row = bounds[r] + (want - acc) / max(dens[r], 1e-12)
        row = 0.5 * (row + bounds[k])                       # damped
        row = int(round(row / quantum)) * quantum
        row = max(out[-1] + quantum, min(row, height - (n - k) * quantum))
        out.append(row)
    out.append(height)
    return out


class DevTensor:
    """Zero-copy torch view of a device allocation owned by the C library."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 3}


class Session:
    """One process = one GPU: rank bookkeeping, the process group, the product library."""

    def __init__(self, args):
        import torch
        self.torch = torch
        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist_mod
            self.dist = dist_mod
            self.dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.lib = load_b200()
        L = self.lib.lib
        L.mtgl_set_device(self.local)
        L.mtgl_dev_get_stats.argtypes = [ctypes.c_void_p, ctypes.POINTER(Stats)]
        L.mtgl_dev_timer_mark.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.mtgl_dev_timer_elapsed_ms.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_float)]
        L.mtgl_dev_set_band.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        L.mtgl_dev_plane_pointers.argtypes = [ctypes.c_void_p] + [ctypes.POINTER(ctypes.c_void_p)] * 3
        L.mtgl_dev_read_framebuffer.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int] + [ctypes.c_void_p] * 3
        L.mtgl_dev_export_color_plane.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.mtgl_dev_set_present_target.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.mtgl_dev_frame_barrier.argtypes = [ctypes.c_void_p, ctypes.c_uint]
        L.mtgl_context_buffer_pointer.argtypes = [ctypes.c_void_p, ctypes.c_uint, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_uint64)]
        L.gl_get_current_context.restype = ctypes.c_void_p
        L.mtgl_dev_stream.restype = ctypes.c_void_p
        L.mtgl_dev_stream.argtypes = [ctypes.c_void_p]
        L.glBindBuffer.argtypes = [ctypes.c_uint, ctypes.c_uint]
        L.glBufferData.argtypes = [ctypes.c_uint, ctypes.c_long, ctypes.c_void_p, ctypes.c_uint]
        L.glBufferSubData.argtypes = [ctypes.c_uint, ctypes.c_long, ctypes.c_long, ctypes.c_void_p]
        L.scene_c4_host_data.restype = ctypes.c_void_p
        L.scene_c4_vbo.restype = ctypes.c_uint
        L.scene_c4_vertex_count.restype = ctypes.c_int
        L.mtglBufferDataPinned.argtypes = [ctypes.c_uint, ctypes.c_long, ctypes.c_void_p, ctypes.c_uint]
        L.mtglReadColorAsync.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p]

    def close(self):
        if self.dist:
            self.dist.destroy_process_group()


def parity_block(sess, workload, assembled_color=None):
    """The checker leg (rank 0): the workload's frame rendered by the unmodified reference (strict IEEE build) on this
    box's host CPU, compared with the CUDA frame -- all three planes of a fresh single-GPU render, and for N > 1 also the
    colour plane assembled from the bands in rank 0's framebuffer."""
    from parity import compare_planes
    key, w, h, variant = WORKLOADS[workload]
    scene = SCENE_OF[workload]
    cnt = counts_for(key)
    t0 = time.perf_counter()
    ref = load_reference("strict").render(scene, w, h, variant)
    t_ref = time.perf_counter() - t0
    if os.environ.get("MTGL_BENCH_BAND"):
        return {"skipped": "band emulation"}
    got = sess.lib.render(scene, w, h, variant)
    s = compare_planes(ref, got)
    out = {"reference": "oracle/_ref/libref_strict.so (unmodified reference, -O2 -fno-fast-math -ffp-contract=off)",
           "reference_frame_s": round(t_ref, 2),
           "stencil_diff": s["stencil_diff_pixels"], "depth_max_ulp": s["depth_max_ulp"], "depth_diff_pixels": s["depth_diff_pixels"],
           "color_max_abs": s["color_max_abs"], "color_diff_pixels": s["color_diff_pixels"], "color_identical_frac": s["color_identical_frac"],
           "gl_error": [int(ref[3]), int(got[3])]}
    if workload in ("c3", "c3s"):      # stencil INCR_WRAP on every covered fragment (<= 128 per pixel): the plane sums to the covered-fragment count
        out["covered_fragments"] = int(got[2].astype(np.int64).sum())
        out["covered_fragments_expected"] = cnt["covered"]
    else:                      # pixels some fragment passed the depth test on
        out["covered_pixels"] = int((got[1] != np.float32(1.0)).sum())
        out["covered_pixels_reference"] = int((ref[1] != np.float32(1.0)).sum())
    if assembled_color is not None:
        a = assembled_color.reshape(h, w)
        ch = np.abs(a.view(np.uint8).reshape(h, w, 4).astype(np.int16) - ref[0].view(np.uint8).reshape(h, w, 4).astype(np.int16))
        out["assembled_color_max_abs"] = int(ch.max())
        out["assembled_color_identical_frac"] = float((a == ref[0]).mean())
    ok = (s["stencil_diff_pixels"] == 0 and s["depth_max_ulp"] <= 1 and s["color_max_abs"] <= 1 and s["color_identical_frac"] >= 0.999
          and out.get("covered_fragments", 0) == out.get("covered_fragments_expected", 0)
          and out.get("covered_pixels", 0) == out.get("covered_pixels_reference", 0)
          and out.get("assembled_color_max_abs", 0) <= 1)
    out["gate"] = "green" if ok else "RED"
    return out


def measure(sess, workload, primary):
    """One workload on this rank's GPU: device-timed frames, end-to-end frames, roofline of the dominant kernel."""
    torch, dist, args = sess.torch, sess.dist, sess.args
    rank, world, local = sess.rank, sess.world, sess.local
    lib = sess.lib
    L = lib.lib
    key, w, h, variant = WORKLOADS[workload]
    cnt = counts_for(key)
    lib.create(w, h)
    dev = lib.device()

    bounds = [band_rows(h, r, world)[0] for r in range(world)] + [h]     # band r = rows [bounds[r], bounds[r + 1])
    y0, y1 = bounds[rank], bounds[rank + 1]
    if world > 1:
        assert L.mtgl_dev_set_band(dev, y0, y1) == 0
    elif os.environ.get("MTGL_BENCH_BAND"):      # diagnostics: one GPU plays rank r of n ("r/n"); the line is not a bench value
        er, en = (int(x) for x in os.environ["MTGL_BENCH_BAND"].split("/"))
        y0, y1 = band_rows(h, er, en)
        assert L.mtgl_dev_set_band(dev, y0, y1) == 0

    is_c3 = workload in ("c3", "c3s")
    if is_c3:
        L.scene_c3_setup(w, h, variant)     # state + texture once; the frame is the clear and the 64 quads
    else:
        L.scene_c4_setup(w, h, variant)

    def frame():
        if is_c3:
            L.scene_c3_draw()
        else:
            L.scene_c4_draw()

    # colour plane as a torch tensor for the unfused gather variant
    cptr = ctypes.c_void_p()
    L.mtgl_dev_plane_pointers(dev, ctypes.byref(cptr), None, None)
    color_dev = torch.as_tensor(DevTensor(cptr.value, w * h * 4), device=f"cuda:{local}") if world > 1 else None

    # Fused gather (default for N > 1): rank 0 exports its colour plane over CUDA IPC, the other ranks map it and their
    # raster kernels store every colour of their band straight into it over NVLink (mtgl_dev_set_present_target);
    # what is left of the gather is the frame-barrier kernel.
    # "auto": the fused peer stores, unless the bands entering the presenting GPU exceed 64 MB per frame (8K frames on many
    # GPUs) -- its NVLink ingest is then the bottleneck and fused stores stall the SMs that issue them
    gather_mode = args.gather
    if gather_mode == "auto":
        gather_mode = "copy" if w * h * 4 * (world - 1) / max(world, 1) > 64e6 else "peer"
    peer = world > 1 and gather_mode in ("peer", "copy")
    if world > 1 and gather_mode == "copy":     # bands pushed by an asynchronous peer-to-peer copy behind each frame (MTGL_PRESENT_COPY)
        L.mtgl_dev_set_present_mode.argtypes = [ctypes.c_void_p, ctypes.c_int]
        assert L.mtgl_dev_set_present_mode(dev, 1) == 0
    if peer:
        hbuf = (ctypes.c_ubyte * 64)()
        if rank == 0:
            assert L.mtgl_dev_export_color_plane(dev, hbuf) == 0
        ht = torch.tensor(list(hbuf), dtype=torch.uint8, device=f"cuda:{local}")
        dist.broadcast(ht, 0)
        if rank != 0:
            hb = (ctypes.c_ubyte * 64)(*ht.cpu().tolist())
            rc = L.mtgl_dev_set_present_target(dev, hb)
            assert rc == 0, f"mtgl_dev_set_present_target failed ({rc})"
        dist.barrier()          # every rank has mapped the plane (and its barrier counter) before the first frame

    def gather():
        """colour rows of every band -> rank 0's framebuffer"""
        if world == 1:
            return
        if peer:        # the stores already went to rank 0's plane; wait until every rank's frame is complete:
            L.glFlush()     # one barrier kernel on the library's stream behind the raster kernels, counter in rank 0's HBM
            assert L.mtgl_dev_frame_barrier(dev, world) == 0
            L.glFinish()
            return
        L.glFinish()
        ops = []
        if rank == 0:
            for r in range(1, world):
                a, b = bounds[r], bounds[r + 1]
                if b > a:
                    ops.append(dist.P2POp(dist.irecv, color_dev[a * w * 4:b * w * 4], r))
        elif y1 > y0:
            ops.append(dist.P2POp(dist.isend, color_dev[y0 * w * 4:y1 * w * 4], 0))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()

    def sync_all():
        L.glFinish()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
            torch.cuda.synchronize()

    # L2 hygiene: C4/C5 stream > 126 MB per frame (98.9 MB of vertices, > 100 MB of triangle records, 74-300 MB of planes),
    # i.e. inputs larger than L2; the C3 working set is small, so flush L2 between its frames.
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{local}") if is_c3 else None

    warm = max(args.warmup, 3)
    for _ in range(warm):
        frame(); gather()
    sync_all()

    st = Stats()
    # Load-balanced bands (N > 1, frames that start with a full clear -- depth and stencil of rows that change owner are
    # not exchanged): the uniform split gives the ranks in the middle of C4's grid ~35 % more work than the ones that own
    # the margins.  Ten untimed feedback rounds move the boundaries (16-row quantum) towards equal device time per rank;
    # the best split seen is kept.  The assembled frame is checked against a single-GPU render after the timed loops.
    balance_log = None
    if world > 1 and not args.uniform_bands:
        best = (float("inf"), list(bounds))
        balance_log = []
        for it in range(10):
            if it:
                frame(); gather()
                sync_all()
            frame(); gather()
            sync_all()
            L.mtgl_dev_get_stats(dev, ctypes.byref(st))
            tt = torch.tensor([float(st.last_batch_ms)], device=f"cuda:{local}")
            allt = [torch.zeros_like(tt) for _ in range(world)]
            dist.all_gather(allt, tt)
            times = [float(x.item()) for x in allt]
            balance_log.append({"bounds": list(bounds), "max_ms": max(times)})
            if max(times) < best[0]:
                best = (max(times), list(bounds))
            bounds = rebalanced(bounds, times, h) if it < 9 else best[1]
            y0, y1 = bounds[rank], bounds[rank + 1]
            assert L.mtgl_dev_set_band(dev, y0, y1) == 0
        frame(); gather()
        sync_all()

    L.mtgl_dev_get_stats(dev, ctypes.byref(st))
    launches0 = st.kernel_launches
    sampler = ClockSampler(local)
    if rank == 0 and primary:
        sampler.start()

    # ---- timed region 1: inputs resident in HBM ----
    # K frames back to back.  C4/C5 are pipelined the way a render loop is: glFlush() hands each frame to the device and
    # returns (the host prepares frame i+1 while the GPU rasterises frame i), one glFinish() ends the region; the device
    # time is taken between two CUDA events on the library's stream.  For N > 1 every frame ends with the library's frame
    # barrier (mtgl_dev_frame_barrier: one tiny kernel per rank on the same stream, counter in the presenting GPU's HBM
    # over NVLink): frame i+1 starts on no rank before every band of frame i has landed in rank 0's plane, and there is
    # no host synchronisation and no NCCL call inside the loop.  C3 flushes L2 between frames (on the same stream,
    # outside the per-frame event pairs).
    pipelined = world == 1 or peer
    sync_all()
    L.mtgl_dev_get_stats(dev, ctypes.byref(st))
    cum0 = (st.batches, st.cum_batch_ms, np.array(list(st.cum_stage_ms)), np.array(list(st.cum_raster_ms)))
    dev_ms_total = 0.0
    t_wall0 = time.perf_counter()
    if pipelined and flush is not None and not (world > 1 and gather_mode == "copy"):
        # C3: the L2 flush runs on the library's own stream between the frames and every frame sits between its own pair
        # of CUDA events on that stream, so the host queues ahead of the device and no host time (Python, a page fault,
        # a descheduled thread) lands between the events; the flushes are outside the pairs.
        lib_stream = torch.cuda.ExternalStream(L.mtgl_dev_stream(dev), device=f"cuda:{local}")
        pairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for e0, e1 in pairs:
            L.glFlush()                 # whatever the previous step left queued is on the stream before the flush
            with torch.cuda.stream(lib_stream):
                flush.fill_(1)
            e0.record(lib_stream)
            frame()
            L.glFlush()
            if world > 1:
                L.mtgl_dev_frame_barrier(dev, world)
            e1.record(lib_stream)
        L.glFinish()
        sync_all()
        dev_ms_total = float(sum(e0.elapsed_time(e1) for e0, e1 in pairs))
    elif pipelined and flush is None:
        L.mtgl_dev_timer_mark(dev, 0)
        for _ in range(args.steps):
            frame()
            L.glFlush()
            if world > 1:
                L.mtgl_dev_frame_barrier(dev, world)
        L.mtgl_dev_timer_mark(dev, 1)
        L.glFinish()
        ms = ctypes.c_float()
        L.mtgl_dev_timer_elapsed_ms(dev, ctypes.byref(ms))
        dev_ms_total = ms.value
    else:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(args.steps):
            if flush is not None:
                flush.fill_(1); torch.cuda.synchronize()
                if dist:
                    dist.barrier(); torch.cuda.synchronize()     # the ranks start the frame together
            L.mtgl_dev_timer_mark(dev, 0)
            frame()
            gather()
            L.glFinish()
            L.mtgl_dev_timer_mark(dev, 1)
            ms = ctypes.c_float()
            L.mtgl_dev_timer_elapsed_ms(dev, ctypes.byref(ms))
            dev_ms_total += ms.value
        ev1.record()
    sync_all()
    wall_s = time.perf_counter() - t_wall0
    # device time of the K steps: CUDA events on the library's stream (C3: around each step, flushes excluded); the
    # unfused NCCL gather runs on torch's stream, so the bracketing torch events are used for it instead
    elapsed_ms = dev_ms_total if (world == 1 or peer) else ev0.elapsed_time(ev1)
    if dist:
        tt = torch.tensor([elapsed_ms], device=f"cuda:{local}")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        elapsed_ms = float(tt.item())
    # per-kernel-group CUDA-event times of exactly these K frames (the library keeps them per batch and sums them up)
    L.mtgl_dev_get_stats(dev, ctypes.byref(st))
    assert st.batches - cum0[0] == args.steps, (st.batches, cum0[0])
    stage = np.array(list(st.cum_stage_ms)) - cum0[2]
    rstage = np.array(list(st.cum_raster_ms)) - cum0[3]
    batch_ms = st.cum_batch_ms - cum0[1]
    if os.environ.get("MTGL_BENCH_DEBUG"):
        print(f"[rank {rank}] {workload} stages {np.round(stage / args.steps, 4).tolist()} raster {np.round(rstage / args.steps, 4).tolist()} dev_ms {dev_ms_total / args.steps:.4f} records {int(st.triangles_setup)} tile_refs {int(st.tile_refs)}", file=sys.stderr, flush=True)
    launches = int(st.kernel_launches - launches0)
    clocks = sampler.stop() if (rank == 0 and primary) else None

    # ---- timed region 2: end to end through the public API with host buffers (C4/C5: VBO re-upload + read-back) ----
    h2d = d2h = 0
    # N = 1: this rank uploads the whole VBO and reads the whole colour plane back.
    # N > 1 (peer gather): every rank uploads 1/N of the VBO from its pinned copy, an NCCL all-gather over NVLink
    # completes the buffer on every GPU (mtgl_dev_buffer_pointer), each rank renders its band into rank 0's plane
    # and rank 0 alone reads the whole frame back -- the bytes below are per step, summed over ranks.
    whole_on_rank0 = peer
    ry0, ry1 = (0, h) if (whole_on_rank0 and rank == 0) else ((y0, y1) if not whole_on_rank0 else (0, 0))
    pinned_out = torch.empty((ry1 - ry0) * w, dtype=torch.int32).pin_memory() if ry1 > ry0 else None
    vbo_dev = None
    if not is_c3:
        nbytes = int(L.scene_c4_vertex_count()) * 32          # what the VBO holds (c4i: one mesh, not the whole grid)
        pinned_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
        ctypes.memmove(pinned_in.data_ptr(), L.scene_c4_host_data(), nbytes)
        vbo = L.scene_c4_vbo()
        if peer and nbytes % world == 0:
            bp, bs = ctypes.c_void_p(), ctypes.c_uint64()
            assert L.mtgl_context_buffer_pointer(L.gl_get_current_context(), vbo, ctypes.byref(bp), ctypes.byref(bs)) == 0 and bs.value == nbytes
            vbo_dev = torch.as_tensor(DevTensor(bp.value, nbytes), device=f"cuda:{local}")
        h2d = nbytes if (world == 1 or vbo_dev is not None) else nbytes * world      # bytes per step, summed over ranks
    d2h = h * w * 4

    def e2e_step():
        if not is_c3:
            L.glBindBuffer(GL_ARRAY_BUFFER, vbo)
            if vbo_dev is not None:
                part = nbytes // world
                L.glBufferSubData(GL_ARRAY_BUFFER, rank * part, part, pinned_in.data_ptr() + rank * part)
                dist.all_gather_into_tensor(vbo_dev, vbo_dev[rank * part:(rank + 1) * part])
                torch.cuda.current_stream().synchronize()
            else:
                L.glBufferData(GL_ARRAY_BUFFER, nbytes, pinned_in.data_ptr(), GL_STATIC_DRAW)
        frame()
        if whole_on_rank0:
            gather()
            torch.cuda.current_stream().synchronize()
        else:
            L.glFinish()                # the frame is rendered before it is read
        if pinned_out is not None:      # read the colour rows into pinned host memory
            base = pinned_out.data_ptr() - ry0 * w * 4
            assert L.mtgl_dev_read_framebuffer(dev, ry0, ry1, base, None, None) == 0

    # N = 1: the pipelined form of the same step (include/mtgl_context.h): the upload is queued from pinned memory into
    # fresh storage (mtglBufferDataPinned), the frame is queued behind it, the read-back of its colour plane is queued
    # behind the frame (mtglReadColorAsync, two pinned targets in turn) -- so frame i+1's upload crosses PCIe while frame
    # i is rasterised and read back.  Every step still uploads its whole VBO and reads its whole frame back, and the timed
    # region ends when the last read-back has landed (glFinish).
    pipelined_e2e = world == 1 and not args.serial_e2e
    host_gather = peer and not args.serial_e2e
    host_gather_check = None
    if host_gather:
        # N > 1: the frame is assembled in HOST memory.  Rank 0 creates a page-locked frame that all ranks map (POSIX shared
        # memory registered with CUDA); every rank queues the read-back of its own band into it behind its frame
        # (mtglReadColorAsync) -- N PCIe links in parallel instead of rank 0 reading all 33 MB -- and goes on to the next
        # step: the sharded upload + NCCL all-gather of step i+1 overlaps the read-back of step i.  Two frames in turn.
        from multiprocessing import resource_tracker, shared_memory
        names = [None]
        if rank == 0:
            shm = shared_memory.SharedMemory(create=True, size=2 * h * w * 4)
            names = [shm.name]
        dist.broadcast_object_list(names, 0)
        if rank != 0:
            shm = shared_memory.SharedMemory(name=names[0])
            try:
                resource_tracker.unregister(shm._name, "shared_memory")      # rank 0 owns the segment
            except Exception:
                pass
        host = np.ndarray((2, h * w), dtype=np.uint32, buffer=shm.buf)
        rc = torch.cuda.cudart().cudaHostRegister(host.ctypes.data, 2 * h * w * 4, 0)
        assert int(rc) == 0, f"cudaHostRegister failed ({rc})"

        # The upload is pipelined too: every step the VBO name gets fresh storage (mtgl_context_buffer_orphan), this rank's
        # slice goes up from pinned memory and the NCCL all-gather completes the buffer on a side stream, and the library's
        # stream waits for that with an event -- so the fill of step i+1 overlaps the frame of step i on every GPU.
        L.mtgl_context_buffer_orphan.argtypes = [ctypes.c_void_p, ctypes.c_uint, ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_uint64)]
        # The buffer goes up in `pieces` chunks, each cut into one slice per rank: while the all-gather of chunk c runs over
        # NVLink (gather stream), this rank's slice of chunk c + 1 crosses PCIe (upload stream).
        up_stream = torch.cuda.Stream(device=f"cuda:{local}")
        gather_stream = torch.cuda.Stream(device=f"cuda:{local}")
        lib_stream = torch.cuda.ExternalStream(L.mtgl_dev_stream(dev), device=f"cuda:{local}")
        ctx = L.gl_get_current_context()
        pieces = int(os.environ.get("MTGL_BENCH_PIECES", "4"))
        if not nbytes_ok(nbytes if not is_c3 else 0, pieces * world):
            pieces = 1

        def e2e_step(k=0):
            if not is_c3:
                if vbo_dev is not None:
                    bp, bs = ctypes.c_void_p(), ctypes.c_uint64()
                    assert L.mtgl_context_buffer_orphan(ctx, vbo, pinned_in.data_ptr(), ctypes.byref(bp), ctypes.byref(bs)) == 0 and bs.value == nbytes
                    fresh = torch.as_tensor(DevTensor(bp.value, nbytes), device=f"cuda:{local}")
                    chunk = nbytes // pieces
                    sub = chunk // world
                    up_done = [torch.cuda.Event() for _ in range(pieces)]
                    landed = torch.cuda.Event()
                    with torch.cuda.stream(up_stream):
                        for c in range(pieces):
                            o = c * chunk + rank * sub
                            fresh[o:o + sub].copy_(pinned_in[o:o + sub], non_blocking=True)
                            up_done[c].record(up_stream)
                    with torch.cuda.stream(gather_stream):
                        for c in range(pieces):
                            o = c * chunk + rank * sub
                            gather_stream.wait_event(up_done[c])
                            dist.all_gather_into_tensor(fresh[c * chunk:(c + 1) * chunk], fresh[o:o + sub])
                        landed.record(gather_stream)
                    lib_stream.wait_event(landed)
                else:
                    L.glBindBuffer(GL_ARRAY_BUFFER, vbo)
                    L.glBufferData(GL_ARRAY_BUFFER, nbytes, pinned_in.data_ptr(), GL_STATIC_DRAW)
            frame()
            if y1 > y0:
                L.mtglReadColorAsync(y0, y1, host[k & 1].ctypes.data)

        e2e_s, e2e_runs = timed_e2e(lambda k: e2e_step(k), lambda: (L.glFinish(), sync_all()), args.steps, dist, torch, local)
        host[0].fill(0)                                         # the check below must see this frame, not an earlier one
        sync_all()
        e2e_step(0); L.glFinish(); sync_all()
        if rank == 0:
            host_frame = host[0].copy()
        sync_all()
        torch.cuda.cudart().cudaHostUnregister(host.ctypes.data)
        del host
        shm.close()
        if rank == 0:
            shm.unlink()
    elif pipelined_e2e:
        outs = [pinned_out, torch.empty_like(pinned_out).pin_memory()]

        def e2e_step(k=0):
            if not is_c3:
                L.glBindBuffer(GL_ARRAY_BUFFER, vbo)
                L.mtglBufferDataPinned(GL_ARRAY_BUFFER, nbytes, pinned_in.data_ptr(), GL_STATIC_DRAW)
            frame()
            L.mtglReadColorAsync(0, h, outs[k & 1].data_ptr())

        e2e_s, e2e_runs = timed_e2e(lambda k: e2e_step(k), lambda: (L.glFinish(), sync_all()), args.steps, dist, torch, local)
        if not is_c3 and os.environ.get("MTGL_BENCH_BAND") is None:      # the frames that came back are the frame
            want = np.empty(h * w, dtype=np.uint32)
            assert L.mtgl_dev_read_framebuffer(dev, 0, h, want.ctypes.data, None, None) == 0
            for o in outs:
                assert np.array_equal(o.numpy().view(np.uint32), want), "pipelined read-back differs from the plane"
    else:
        e2e_step()
        sync_all()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        sync_all()
        e2e_s = time.perf_counter() - t0
        if dist:
            tt = torch.tensor([e2e_s], device=f"cuda:{local}")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e2e_s = float(tt.item())
        e2e_runs = [e2e_s]
    e2e_value = cnt["covered"] * args.steps / e2e_s

    # ---- N > 1: the frame assembled in rank 0's plane must be bit-identical to a single-GPU render ----
    gather_check = None
    assembled = None
    if world > 1:
        frame(); gather(); sync_all()
        if rank == 0:
            got = np.empty(h * w, dtype=np.uint32)
            assert L.mtgl_dev_read_framebuffer(dev, 0, h, got.ctypes.data, None, None) == 0
            assert L.mtgl_dev_set_band(dev, 0, h) == 0
            frame(); L.glFinish()
            want = np.empty(h * w, dtype=np.uint32)
            assert L.mtgl_dev_read_framebuffer(dev, 0, h, want.ctypes.data, None, None) == 0
            gather_check = "bit-identical to a single-GPU render" if np.array_equal(got, want) else f"MISMATCH in {int((got != want).sum())} pixels"
            if host_gather:
                host_gather_check = "bit-identical to a single-GPU render" if np.array_equal(host_frame, want) else f"MISMATCH in {int((host_frame != want).sum())} pixels"
            assembled = got
        sync_all()

    lib.destroy()

    # ---- checker legs on this box's host CPU (rank 0): the reference as timing baseline and as parity oracle ----
    cpu = None
    if rank == 0 and world == 1 and primary and not args.no_cpu_baseline:
        try:
            cpu = run_reference(args, workload, steps=2 if not is_c3 else 1, warmup=1 if not is_c3 else 0)["cpu_baseline"]
        except Exception as e:  # pragma: no cover
            cpu = {"value": None, "unit": "fragments/s", "cores": 1, "kind": "reference", "sample": f"unavailable: {e}"}
    parity = None
    if rank == 0 and not args.no_parity and not (is_c3 and world > 1):    # the 64-quad reference frame takes ~50 s: at N = 1 only
        try:
            parity = parity_block(sess, workload, assembled)
        except Exception as e:  # pragma: no cover
            parity = {"gate": "unavailable", "error": repr(e)}
    elif is_c3 and world > 1:
        parity = {"gate": "see N=1", "note": "the 64-quad reference frame takes ~50 s of host time; checked in the N=1 run, gather_check covers N>1"}
    if dist:
        dist.barrier()
    if rank != 0:
        return None

    ms_per_step = elapsed_ms / args.steps
    value = cnt["covered"] * args.steps / (elapsed_ms * 1e-3)
    bw, bw_src = peaks()
    px = w * h
    if is_c3:       # stencil R+W on every covered fragment, blend read + colour write on shaded ones; clear colour+stencil
        frag_bytes = cnt["covered"] * 2 + cnt["shaded"] * 8
        clear_bytes = px * (4 + 1)
        vertex_bytes = 0
    else:           # depth read on every covered fragment, depth + colour write on passing ones; clear colour+depth
        frag_bytes = cnt["covered"] * 4 + cnt["tested"] * 8
        clear_bytes = px * (4 + 4)
        vertex_bytes = cnt["vertices"] * 32
    # the dominant kernel of the step, whichever the live per-group CUDA-event times say: the fused vertex + set-up
    # kernel (K1+K2), K4a (visibility: coverage + depth), K4b (shade) or the in-order kernels (k_fill / k_raster<false>).
    # Its algorithmic bytes are SURVEY.md 8(d)'s figures restricted to what that kernel owns (DESIGN.md "Rooflines"): the
    # enabled attribute arrays per input vertex for set-up, the reference's per-fragment framebuffer traffic for the
    # planes a raster kernel owns.
    groups = ["k_setup (K1+K2 vertex + set-up)", "k_vis (K4a visibility)", "k_shade (K4b)", "k_fill + k_raster<false> (in-order)"]
    group_ms = [stage[1] / args.steps] + [float(x) / args.steps for x in rstage]
    gi = int(np.argmax(group_ms))
    raster_ms = group_ms[gi]
    if is_c3:
        group_bytes = [0, 0, 0, frag_bytes + clear_bytes]
    else:
        group_bytes = [vertex_bytes, cnt["covered"] * 4 + cnt["tested"] * 4 + px * 4, cnt["tested"] * 4 + px * 4, 0]
        if group_bytes[gi] == 0:        # a state mix that sends C4/C5 through the in-order kernels
            group_bytes[gi] = frag_bytes + clear_bytes
    raster_bytes = group_bytes[gi] / world          # one launch per rank covers 1/N of the frame (set-up: the chunks that reach its band)
    achieved = raster_bytes / (raster_ms * 1e-3) / 1e9 if raster_ms > 0 else 0.0
    traffic, traffic_src = None, None
    try:
        tfile = ROOT / "profiles" / "r02_ncu_traffic.json"
        tj = json.loads(tfile.read_text()).get(workload, {}).get(groups[gi].split(" ")[0])
        if tj and world == 1:
            traffic = tj["dram_read_bytes"] + tj["dram_write_bytes"]
            traffic_src = f"profiles/r02_ncu_traffic.json ({tj.get('capture', 'ncu --set full')}, per launch)"
    except Exception:
        traffic = None
    frame_bytes = frag_bytes + clear_bytes + vertex_bytes
    out = {
        "metric": "covered_fragments_per_s", "value": value, "unit": "fragments/s", "n_gpus": world, "steps": args.steps,
        "warmup": warm, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "frames_per_s": 1e3 / ms_per_step, "triangles_per_s": cnt["vertices"] / 3 * 1e3 / ms_per_step,
        "config": workload_config(workload, world),
        "multi_gpu": None if world == 1 else {"gather": ("asynchronous NVLink peer copy per band + frame-barrier kernel on a side stream" if gather_mode == "copy" else
                                                         "fused NVLink peer-store gather + frame-barrier kernel") if peer else "NCCL send/recv gather",
                                              "band_rows": bounds, "band_balance": balance_log, "gather_check": gather_check,
                                              "host_gather_check": host_gather_check},
        "e2e": {"value": e2e_value, "unit": "fragments/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": e2e_s * 1e3 / args.steps,
                "runs_ms_per_step": [t * 1e3 / args.steps for t in e2e_runs], "reported": "median of the runs (K steps each)",
                "mode": ("pipelined across frames (mtglBufferDataPinned / mtglReadColorAsync)" if (world == 1 and not args.serial_e2e) else
                         "pipelined: sharded upload + NCCL all-gather into orphaned storage on a side stream, every rank reads its band back into one shared page-locked frame (mtglReadColorAsync)" if host_gather else
                         "upload, render, read-back in sequence")},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": {"kernel": groups[gi], "bound": "hbm", "achieved": achieved, "peak": bw, "unit": "GB/s",
                     "frac": achieved / bw, "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": bw_src,
                     "algorithmic_bytes_per_launch": raster_bytes, "kernel_ms": raster_ms,
                     "frame_bytes": frame_bytes, "frame_frac": frame_bytes / (bw * 1e9) / (ms_per_step * 1e-3)},
        "stages_ms": {k: float(v / args.steps) for k, v in zip(["vertex", "setup", "bin_count_scan", "bin_fill", "raster"], stage)},
        "raster_ms": {k: float(v / args.steps) for k, v in zip(["visibility", "shade", "in_order"], rstage)},
        "batch_ms": batch_ms / args.steps, "wall_ms_per_step": wall_s * 1e3 / args.steps,
        "chunks_culled_rank0": int(st.chunks_culled), "chunks": (cnt["vertices"] // 3 + 255) // 256,
        "cpu_baseline": cpu,
        "parity": parity,
    }
    return out


SECONDARY_KEYS = ("value", "unit", "ms_per_step", "frames_per_s", "config", "multi_gpu", "e2e", "gpu_launches", "roofline",
                  "stages_ms", "raster_ms", "parity")


def run_b200(args):
    sess = Session(args)
    try:
        out = measure(sess, args.workload, primary=True)
        secondary = {}
        if not args.no_secondary and args.workload == "c4":
            for wl in ("c3", "c5"):
                r = measure(sess, wl, primary=False)
                if r is not None:
                    secondary[WORKLOADS[wl][0]] = {k: r[k] for k in SECONDARY_KEYS}
        if out is not None:
            out["secondary"] = secondary or None
        return out
    finally:
        sess.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the C3 and C5 lines of the default (C4) run")
    ap.add_argument("--no-parity", action="store_true", help="skip the reference-rendered parity check of each workload")
    ap.add_argument("--serial-e2e", action="store_true", help="N = 1: upload, render and read back strictly in sequence (no pipelining across frames)")
    ap.add_argument("--uniform-bands", action="store_true", help="N > 1: keep the uniform split of tile rows (no load balancing)")
    ap.add_argument("--gather", default="auto", choices=["auto", "peer", "copy", "nccl"],
                    help="N > 1: 'peer' = raster kernels store into rank 0's plane over NVLink (fused), 'copy' = one asynchronous "
                         "peer-to-peer copy per band behind the frame, overlapped with the next frame's geometry, 'nccl' = send/recv after the frame")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        if rank == 0:
            print(json.dumps(run_reference(args, args.workload)), flush=True)
        return
    # stdout carries the one JSON line and nothing else: whatever the libraries underneath print there while the bench
    # runs (NCCL's version banner, for one) is sent to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        out = run_b200(args)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    if out is not None:
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
