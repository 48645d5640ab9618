#!/bin/bash
# One GPU-box pass: parity tests, bench lines, ncu launch list and (optionally) one --set full capture.
#   gpurun --timeout 900 -- 'bash tools/gpu_check.sh TAG [ncu] [c3] [c5]'
TAG=${1:-run}; shift
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/${TAG}_tests.log 2>&1; tail -2 $O/${TAG}_tests.log
python bench.py --steps 20 --warmup 3 > $O/${TAG}_bench_c4.json 2> $O/${TAG}_bench_c4.err
for a in "$@"; do
  case $a in
    c3) python bench.py --workload c3 --steps 5 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_c3.json 2>>$O/${TAG}_bench_c4.err;;
    c5) python bench.py --workload c5 --steps 20 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_c5.json 2>>$O/${TAG}_bench_c4.err;;
    ncu)
      ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
          python bench.py --steps 5 --warmup 3 --no-cpu-baseline > $O/${TAG}_launches.log 2>&1
      ncu --set full --clock-control none --import-source on -k regex:'k_setup|k_vis|k_shade|k_bin_small' -s 12 -c 4 \
          -o $O/${TAG}_prof_c4 -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${TAG}_ncu.log 2>&1;;
  esac
done
python - <<P
import json
for w in ("c4","c3","c5"):
    try:
        d=json.loads(open("$O/${TAG}_bench_%s.json"%w).read().strip().splitlines()[-1])
        print(w, round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["stages_ms"].items()}, {k:round(v,3) for k,v in d["raster_ms"].items()}, "e2e", round(d["e2e"]["ms_per_step"],2))
    except Exception as e: print(w, "-", e)
P
