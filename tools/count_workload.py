"""Exact work counters of the BASELINE workloads, measured once with the CPU oracle (front end + restatement,
pinned bit-exact to the reference) and committed as tests/golden/workload_counts.json.  bench.py turns them into
fragments/s and algorithmic bytes; it never runs the oracle for this.

  covered = fragments passing the inclusive inside test (raster.c:539-540)
  tested  = fragments past the stencil and depth tests
  shaded  = fragments reaching the colour write (survive the alpha test)
"""
import ctypes
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from oracle_loader import load_front_oracle  # noqa: E402

WORKLOADS = {
    "c1_suzanne_800x600": ("c1_suzanne", 800, 600, 0),
    "c2_cube_1920x1080": ("c2_cube", 1920, 1080, 0),
    "c3_fill_3840x2160": ("c3_fill", 3840, 2160, 64),
    "c3_shifted_3840x2160": ("c3_fill", 3840, 2160, 64 | 256),      # diagnostic: every quad with its own texture coordinates
    "c4_grid_3840x2160": ("c4_grid", 3840, 2160, 0),
    "c4_grid_phong_3840x2160": ("c4_grid", 3840, 2160, 1 << 16),
    "c5_grid_7680x4320": ("c4_grid", 7680, 4320, 0),
    "c4_instanced_3840x2160": ("c4_grid", 3840, 2160, 1 << 17),    # one 2904-vertex VBO drawn 1064 times under glTranslatef
}


def main():
    lib = load_front_oracle()
    lib.lib.mtgl_oracle_fragment_counts.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64)]
    out_path = ROOT / "tests" / "golden" / "workload_counts.json"
    out = json.loads(out_path.read_text()) if out_path.exists() else {}
    only = sys.argv[1:]
    for key, (name, w, h, variant) in WORKLOADS.items():
        if only and key not in only:
            continue
        t0 = time.time()
        lib.create(w, h)
        assert lib.lib.scene_render(name.encode(), w, h, variant) == 0
        lib.lib.glFinish()
        counts = (ctypes.c_uint64 * 3)()
        lib.lib.mtgl_oracle_fragment_counts(lib.device(), counts)
        verts = lib.lib.scene_c4_vertex_count() if name == "c4_grid" else 0
        if variant & (1 << 17):
            verts *= 38 * 28        # vertices through the vertex stage: the one mesh, once per draw
        lib.destroy()
        out[key] = {"scene": name, "width": w, "height": h, "variant": variant, "covered": int(counts[0]),
                    "tested": int(counts[1]), "shaded": int(counts[2]), "vertices": int(verts)}
        print(key, out[key], f"{time.time() - t0:.1f}s", flush=True)
        out_path.write_text(json.dumps(out, indent=1, sort_keys=True) + "\n")


if __name__ == "__main__":
    main()
