#!/bin/bash
# One GPU plays rank r of n (MTGL_BENCH_BAND=r/n): per-rank stage times of the sort-first split without an n-GPU box.
for b in ${@:-0/8 3/8 7/8 1/2 1/4}; do MTGL_BENCH_BAND=$b MTGL_BENCH_DEBUG=1 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 >/dev/null | grep "rank 0" | sed "s|^|band $b |"; done
