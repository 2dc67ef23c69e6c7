#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== chain tests"
timeout 900 python -m pytest tests/test_tile_chain.py -m gpu -x -q > gpurun_out/pytest_chain.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/pytest_chain.log
echo "=== all gpu tests"
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "=== tile bench"
STC_TILE_TIMING=1 STC_CF_TIMING=1 timeout 600 python tools/bench_tile.py --n 12 --reps 2 > gpurun_out/tile_n12.json 2> gpurun_out/tile_n12.err; echo "rc=$?"
cat gpurun_out/tile_n12.json; grep "tile_run" gpurun_out/tile_n12.err | tail -40
echo "=== bench"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
