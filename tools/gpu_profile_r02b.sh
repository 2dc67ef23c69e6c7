#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:conv3x3_umma2_kernel -s 6 -c 6 -o gpurun_out/r02_prof_convs \
    python bench.py --steps 1 --warmup 1 --batch 32 --no-tile-chain --no-cpu-baseline > gpurun_out/r02_prof_convs.log 2>&1; echo "full convs rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'smooth_fused_kernel|k_cloud_refs|k_mosaic_ref|k_gram' -c 6 -o gpurun_out/r02_prof_chain2 \
    python tools/bench_tile.py --n 24 --reps 0 > gpurun_out/r02_prof_chain2.log 2>&1; echo "full chain2 rc=$?"
ls -la gpurun_out/*.ncu-rep
