#!/bin/bash
# Whole-tile chain: per-stage wall times (Python mirrors) and an ncu launch list of the same run.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-12}
STC_TILE_TIMING=1 STC_CF_TIMING=1 timeout 600 python tools/bench_tile.py --n $N --reps 3 > gpurun_out/tile_n$N.json 2> gpurun_out/tile_n$N.err; echo "rc=$?"
cat gpurun_out/tile_n$N.json; tail -60 gpurun_out/tile_n$N.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/tile_launches_n$N.csv \
   python tools/bench_tile.py --n $N --reps 1 > gpurun_out/tile_ncu_n$N.log 2>&1; echo "ncu rc=$?"
python tools/summarize_dram.py gpurun_out/tile_launches_n$N.csv > gpurun_out/tile_launches_n$N.md 2>&1; head -50 gpurun_out/tile_launches_n$N.md
