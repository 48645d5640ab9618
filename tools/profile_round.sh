#!/bin/bash
# The round's N = 1 evidence, in two GPU-box passes (gpurun brings back at most 64 MiB per call):
#   gpurun --timeout 1200 -- 'bash tools/profile_round.sh r02 capture'
#       ncu launch lists and --set full captures of one steady-state C4 / C3 / C5 frame -> gpurun_out/<TAG>_prof_*.ncu-rep;
#       then, here:  python tools/ncu_traffic.py profiles/<TAG>_ncu_traffic.json c4=... c3=... c5=...
#       (bench.py reads that file as roofline.traffic)
#   gpurun --timeout 1200 -- 'bash tools/profile_round.sh r02 bench'
#       parity tests, the default bench line (never under a profiler), the reference arm, the secondary layouts
# Copy what is to be judged from gpurun_out/ into profiles/.
TAG=${1:-r02}
WHAT=${2:-bench}
O=gpurun_out; mkdir -p $O
B="python bench.py --no-cpu-baseline --no-parity --no-secondary"

if [ "$WHAT" = capture ]; then
  for wl in c4 c3 c5; do
    # per-launch durations of a whole short run (cold-cache, serialised: shares, not absolutes)
    ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches_$wl.csv \
        $B --workload $wl --steps 5 --warmup 3 > $O/${TAG}_launches_$wl.log 2>&1
  done
  # one steady-state frame with all sections and source correlation (C3 / C5: the kernels their frames are made of)
  ncu --set full --clock-control none --import-source on -k regex:'k_setup|k_vis|k_shade|k_bin|k_fill|k_raster|k_chunk|k_upload' -s 24 -c 9 \
      -o $O/${TAG}_prof_c4 -f $B --workload c4 --steps 2 --warmup 3 > $O/${TAG}_ncu_c4.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:'k_fill|k_raster|k_bin' -s 12 -c 4 \
      -o $O/${TAG}_prof_c3 -f $B --workload c3 --steps 2 --warmup 3 > $O/${TAG}_ncu_c3.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:'k_setup|k_vis|k_shade|k_bin' -s 15 -c 5 \
      -o $O/${TAG}_prof_c5 -f $B --workload c5 --steps 2 --warmup 3 > $O/${TAG}_ncu_c5.log 2>&1
  for wl in c4 c3 c5; do tail -n 1 $O/${TAG}_ncu_$wl.log; done
  du -sh $O; ls -la $O/*.ncu-rep
  exit 0
fi

python -m pytest tests -m gpu -q > $O/${TAG}_tests.log 2>&1; tail -2 $O/${TAG}_tests.log
python bench.py > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench_n1.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_reference_arm.json 2> $O/${TAG}_reference_arm.err
$B --workload c4i > $O/${TAG}_bench_c4i.json 2>> $O/${TAG}_bench_n1.err
$B --workload c3s --steps 5 > $O/${TAG}_bench_c3s.json 2>> $O/${TAG}_bench_n1.err
python - <<P
import json
d = json.loads(open("$O/${TAG}_bench_n1.json").read().strip().splitlines()[-1])
def line(k, v):
    print(k, round(v["ms_per_step"], 4), "e2e", round(v["e2e"]["ms_per_step"], 3), "frac", round(v["roofline"]["frac"], 4), v["roofline"]["kernel"],
          "traffic", v["roofline"]["traffic"], {a: round(b, 4) for a, b in v["stages_ms"].items()}, {a: round(b, 4) for a, b in v["raster_ms"].items()},
          (v.get("parity") or {}).get("gate"))
line("c4", d)
for k, v in (d.get("secondary") or {}).items():
    line(k, v)
print("clocks", d["clocks"], "cpu", d["cpu_baseline"])
for w in ("c4i", "c3s"):
    e = json.loads(open("$O/${TAG}_bench_%s.json" % w).read().strip().splitlines()[-1])
    print(w, round(e["ms_per_step"], 4))
print(open("$O/${TAG}_reference_arm.json").read()[:300])
P
