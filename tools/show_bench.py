"""One-line summary of bench.py JSON lines:  python tools/show_bench.py file.json [...]"""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], f"N={d['n_gpus']}", "ms/frame", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 3),
              "gather:", d["config"]["gather_check"], {k: round(v, 3) for k, v in d["stages_ms"].items()},
              {k: round(v, 3) for k, v in d["raster_ms"].items()}, "roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"], 4))
    except Exception as e:  # noqa: BLE001
        print(f, "-", e)
