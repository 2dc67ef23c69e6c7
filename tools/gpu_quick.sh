#!/bin/bash
# quick GPU check: selected tests + tile bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q ${1:-} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
STC_TILE_TIMING=1 STC_CF_TIMING=1 timeout 600 python tools/bench_tile.py --n 12 --reps 1 > gpurun_out/tile_n12.json 2> gpurun_out/tile_n12.err; echo "rc=$?"
cat gpurun_out/tile_n12.json; grep "tile_run\|remove_clouds" gpurun_out/tile_n12.err | tail -24
STC_TILE_TIMING=1 STC_CF_TIMING=1 timeout 600 python tools/bench_tile.py --n 24 --reps 1 > gpurun_out/tile_n24.json 2> gpurun_out/tile_n24.err; echo "rc=$?"
cat gpurun_out/tile_n24.json; grep "tile_run\|remove_clouds" gpurun_out/tile_n24.err | tail -24
