#!/bin/bash
# bench.py at N = 8, 4, 2 (and 1) on one multi-GPU box:  gpurun --gpus 8 -- 'bash tools/scale_run.sh TAG'
TAG=${1:-scale}; O=gpurun_out; mkdir -p $O
for n in ${NLIST:-8 4 2 1}; do
  if [ $n = 1 ]; then timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $O/${TAG}_n1.json 2> $O/${TAG}_n1.err
  else MTGL_BENCH_DEBUG=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --steps 20 --warmup 3 > $O/${TAG}_n$n.json 2> $O/${TAG}_n$n.err; fi
  python tools/show_bench.py $O/${TAG}_n$n.json
  grep -o "\[rank [0-9]\] stages[^d]*dev_ms [0-9.]*" $O/${TAG}_n$n.err | sort | head -8
done
