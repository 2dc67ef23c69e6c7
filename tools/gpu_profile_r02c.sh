#!/bin/bash
# Round-2 evidence after the second pass over the chain: ncu launch lists (time + DRAM bytes) of one whole tile through
# stc_tile_run_host (12 and 24 dates) and one `--set full` capture of the kernels that pass rebuilt.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
for N in 12 24; do
  timeout 1200 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02c_launches_tile_n$N.csv \
      python tools/bench_tile.py --n $N --reps 0 > gpurun_out/r02c_tile_under_ncu_n$N.log 2>&1; echo "ncu tile n=$N rc=$?"
  python tools/summarize_dram.py gpurun_out/r02c_launches_tile_n$N.csv 6541 > gpurun_out/r02c_launches_tile_n$N.md 2>&1; head -45 gpurun_out/r02c_launches_tile_n$N.md
done
timeout 1200 ncu --set full --clock-control none --import-source on \
    -k regex:'k_gram|k_sel_hist|k_window2d|k_dilate_col_scan|k_static_refs|k_shadow_candidates|k_nnls|k_mosaic_ref|conv3x3_umma2_kernel<32' -c 40 -o gpurun_out/r02c_prof_chain \
    python tools/bench_tile.py --n 24 --reps 0 > gpurun_out/r02c_prof_chain.log 2>&1; echo "full chain rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -3
