#!/bin/bash
# Round-2 evidence: (1) ncu launch list (time + DRAM bytes) of the default bench command, (2) the same for the whole-tile chain
# (12 and 24 dates), (3) one `--set full` capture of the ConvGRU gates conv, gru_apply2 and the fused smoothing kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
# (1) bench.py, batch 32 so the launch list stays short (one sub-batch per step): steps 2, warm-up 1 as in the recipe
timeout 1200 ncu --metrics $M --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --batch 32 --no-tile-chain --no-cpu-baseline > gpurun_out/r02_bench_under_ncu.log 2>&1; echo "ncu bench rc=$?"
python tools/summarize_dram.py gpurun_out/r02_launches_bench.csv 6541 > gpurun_out/r02_launches_bench.md 2>&1; head -24 gpurun_out/r02_launches_bench.md
# (2) the chain
for N in 12 24; do
  timeout 1200 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_launches_tile_n$N.csv \
      python tools/bench_tile.py --n $N --reps 0 > gpurun_out/r02_tile_under_ncu_n$N.log 2>&1; echo "ncu tile n=$N rc=$?"
  python tools/summarize_dram.py gpurun_out/r02_launches_tile_n$N.csv 6541 > gpurun_out/r02_launches_tile_n$N.md 2>&1; head -40 gpurun_out/r02_launches_tile_n$N.md
done
# (3) full captures
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'conv3x3_umma2_kernel<64, 4, 16' -s 8 -c 2 -o gpurun_out/r02_prof_gates \
    python bench.py --steps 1 --warmup 1 --batch 32 --no-tile-chain --no-cpu-baseline > gpurun_out/r02_prof_gates.log 2>&1; echo "full gates rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'gru_apply2_kernel' -s 4 -c 2 -o gpurun_out/r02_prof_apply2 \
    python bench.py --steps 1 --warmup 1 --batch 32 --no-tile-chain --no-cpu-baseline > gpurun_out/r02_prof_apply2.log 2>&1; echo "full apply2 rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'smooth_fused_kernel|k_rowdist|k_cloud_refs|k_sel_hist' -c 8 -o gpurun_out/r02_prof_chain \
    python tools/bench_tile.py --n 24 --reps 0 > gpurun_out/r02_prof_chain.log 2>&1; echo "full chain rc=$?"
ls -la gpurun_out/*.ncu-rep
