"""Sort-first band partitioning on CPU: two gloo ranks, each rendering its band of rows with the
front end + CPU oracle behind the same C ABI call (mtgl_dev_set_band), colour/depth/stencil bands gathered on
rank 0 and compared bit-for-bit with the committed single-process golden planes of the reference."""
import hashlib
import json
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

CASES = [("c1_suzanne", 800, 600, 0), ("stencil", 300, 200, 0), ("c4_grid", 480, 270, 3 | (2 << 8)), ("blend", 320, 240, 0)]


def band_rows(height, rank, world, align=64):
    tile_rows = (height + align - 1) // align
    return min((tile_rows * rank) // world * align, height), min((tile_rows * (rank + 1)) // world * align, height)


def _worker(rank, world, port, out_path):
    import ctypes
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cases import case_id
    from oracle_loader import load_front_oracle
    lib = load_front_oracle()
    lib.lib.mtgl_dev_set_band.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    digests = {}
    for case in CASES:
        name, w, h, variant = case
        lib.create(w, h)
        # unaligned split on purpose for one case: the band boundary need not be a tile boundary
        y0, y1 = band_rows(h, rank, world, align=64 if name != "stencil" else 50)
        assert lib.lib.mtgl_dev_set_band(lib.device(), y0, y1) == 0
        lib.lib.glClearColor(ctypes.c_float(0.25), ctypes.c_float(0.5), ctypes.c_float(0.75), ctypes.c_float(1.0))
        lib.lib.glClear(0x4000 | 0x0100 | 0x0400)
        lib.lib.glClearColor(ctypes.c_float(0), ctypes.c_float(0), ctypes.c_float(0), ctypes.c_float(1))
        assert lib.lib.scene_render(name.encode(), w, h, variant) == 0
        col, dep, sten = lib.read()
        lib.destroy()
        planes = []
        for plane in (col.view(np.int32), dep.view(np.int32), sten.astype(np.int32)):
            mine = torch.from_numpy(np.ascontiguousarray(plane[y0:y1])).reshape(-1)
            sizes = [(band_rows(h, r, world, 64 if name != "stencil" else 50)) for r in range(world)]
            if rank == 0:       # bands may differ in size: point-to-point, exactly like bench.py's NCCL gather
                bufs = [mine] + [torch.empty((b - a) * w, dtype=torch.int32) for a, b in sizes[1:]]
                for r in range(1, world):
                    dist.recv(bufs[r], src=r)
                planes.append(torch.cat(bufs).numpy().reshape(h, w))
            else:
                dist.send(mine, dst=0)
        if rank == 0:
            digests[case_id(case)] = {
                "color": hashlib.sha256(planes[0].view(np.uint32).tobytes()).hexdigest(),
                "depth": hashlib.sha256(planes[1].view(np.float32).tobytes()).hexdigest(),
                "stencil": hashlib.sha256(planes[2].astype(np.uint8).tobytes()).hexdigest(),
            }
    if rank == 0:
        Path(out_path).write_text(json.dumps(digests))
    dist.destroy_process_group()


def test_two_rank_bands_match_single_process(tmp_path):
    from cases import case_id
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    out = tmp_path / "digests.json"
    mp.spawn(_worker, args=(2, port, str(out)), nprocs=2, join=True)
    got = json.loads(out.read_text())
    golden = json.loads((ROOT / "tests" / "golden" / "planes.json").read_text())
    for case in CASES:
        g = golden[case_id(case)]
        for plane in ("color", "depth", "stencil"):
            assert got[case_id(case)][plane] == g[plane], (case, plane)
