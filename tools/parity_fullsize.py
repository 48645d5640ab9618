"""Full-size parity report (GPU box): renders the BASELINE workloads with the CUDA path and with the CPU oracle
(and the shipped reference build when present) and prints the gate statistics as JSON lines."""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from mytinygl_b200 import load_b200  # noqa: E402
from oracle_loader import load_front_oracle  # noqa: E402
from parity import compare_planes  # noqa: E402

CASES = {"c4": ("c4_grid", 3840, 2160, 0), "c4_phong": ("c4_grid", 3840, 2160, 1 << 16), "c3_8quads": ("c3_fill", 3840, 2160, 8),
         "c3": ("c3_fill", 3840, 2160, 64), "c5": ("c4_grid", 7680, 4320, 0), "c1_4k": ("c1_suzanne", 3840, 2160, 0),
         "c2": ("c2_cube", 1920, 1080, 0)}
which = sys.argv[1:] or ["c4", "c3_8quads", "c1_4k", "c2"]
gpu, cpu = load_b200(), load_front_oracle()
for k in which:
    t0 = time.time(); got = gpu.render(*CASES[k]); t1 = time.time(); ref = cpu.render(*CASES[k]); t2 = time.time()
    s = compare_planes(ref, got)
    s.update(case=k, gpu_s=round(t1 - t0, 3), oracle_s=round(t2 - t1, 3))
    print(json.dumps(s), flush=True)
