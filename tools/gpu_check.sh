#!/bin/bash
# One GPU-box session: parity tests, smoke, bench, ncu launch list + full captures of the dominant kernels.
# Everything lands in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== tests"
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "golden [ab]|out vs|umma vs|passed|failed|FAILED" gpurun_out/pytest_gpu.log | tail -12
echo "=== smoke"
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
echo "=== bench"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
echo "=== bench --impl reference"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "rc=$?"; cat gpurun_out/bench_ref.json
echo "=== ncu launch list (single stream so kernels are attributed cleanly)"
STC_SINGLE_STREAM=1 STC_CONV_PRIO=0 timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 1 --batch 32 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "rc=$?"
echo "=== ncu full capture: GRU conv kernels and the gating kernels"
STC_SINGLE_STREAM=1 STC_CONV_PRIO=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'conv3x3_umma2|gru_apply2' -s 4 -c 4 -o gpurun_out/prof_main \
   python bench.py --steps 1 --warmup 1 --batch 32 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"
ls -la gpurun_out | head -30
