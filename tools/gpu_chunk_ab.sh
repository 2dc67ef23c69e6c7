#!/bin/bash
# A/B of the forward's sub-batch size on the 36 subtile patches of a chain tile
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cloud_fill.py tests/test_tile_chain.py -m gpu -x -q 2>&1 | tail -2
for ch in 32 36 18 12; do
  echo "STC_CHUNK=$ch"
  STC_CHUNK=$ch STC_TILE_TIMING=1 timeout 600 python tools/bench_tile.py --n 12 --reps 0 2> gpurun_out/chunk_$ch.err | grep -o '"chain_ms": \[[^]]*\]'
  grep "subtile gather" gpurun_out/chunk_$ch.err | tail -2
done
