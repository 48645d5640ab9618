"""One-line summary of bench.py JSON lines:  python tools/show_bench.py file.json [...]"""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], f"N={d['n_gpus']}", "ms/frame", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["ms_per_step"], 3),
              "gather:", (d.get("multi_gpu") or {}).get("gather_check"), {k: round(v, 3) for k, v in d["stages_ms"].items()},
              {k: round(v, 3) for k, v in d["raster_ms"].items()}, "roofline", d["roofline"]["kernel"], round(d["roofline"]["frac"], 4))
        for k, v in (d.get("secondary") or {}).items():
            print("   ", k, "ms/frame", round(v["ms_per_step"], 4), "e2e", round(v["e2e"]["ms_per_step"], 3), "gather:", (v.get("multi_gpu") or {}).get("gather_check"),
                  "host gather:", (v.get("multi_gpu") or {}).get("host_gather_check"), "parity:", (v.get("parity") or {}).get("gate"),
                  {a: round(b, 3) for a, b in v["stages_ms"].items()}, {a: round(b, 3) for a, b in v["raster_ms"].items()})
    except Exception as e:  # noqa: BLE001
        print(f, "-", e)
