#!/bin/bash
# quick GPU check: selected tests + tile bench (+ per-kernel table of one traced chain tile)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q ${1:-} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
for n in 12 24; do
  STC_TILE_TIMING=1 STC_CF_TIMING=1 timeout 600 python tools/bench_tile.py --n $n --reps ${2:-1} > gpurun_out/tile_n$n.json 2> gpurun_out/tile_n$n.err; echo "rc=$?"
  cat gpurun_out/tile_n$n.json; grep "tile_run\|remove_clouds" gpurun_out/tile_n$n.err | tail -24
  timeout 600 python tools/bench_tile.py --n $n --reps 0 --trace gpurun_out/tile_trace_q_n$n.csv > gpurun_out/tile_q_n$n.json 2> gpurun_out/tile_q_n$n.err
  grep -o '"chain_ms": \[[^]]*\]' gpurun_out/tile_q_n$n.json; grep -A45 "kernel table" gpurun_out/tile_q_n$n.err
done
