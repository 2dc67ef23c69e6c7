#!/bin/bash
# A/B of the DSen2 convolution occupancy: chain tile with per-kernel table, one and two CTAs per SM
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -x -q -k "superresolve" 2>&1 | tail -3
for occ in 1 2; do
  for n in 12 24; do
    STC_SR_OCC=$occ timeout 600 python tools/bench_tile.py --n $n --reps 0 --trace gpurun_out/tile_trace_occ${occ}_n$n.csv > gpurun_out/tile_occ${occ}_n$n.json 2> gpurun_out/tile_occ${occ}_n$n.err
    echo "occ $occ n $n"; grep "conv_other\|sum " gpurun_out/tile_occ${occ}_n$n.err
  done
done
