#!/bin/bash
# One GPU-box session: parity tests, smoke, bench, ncu launch list + one full capture of the
# dominant kernel.  Everything lands in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== tests"
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "^impl|golden|err|umma vs|passed|failed|FAILED" gpurun_out/pytest_gpu.log | tail -40
echo "=== smoke"
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
echo "=== bench (umma)"
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_umma.json 2> gpurun_out/bench_umma.err; echo "rc=$?"; cat gpurun_out/bench_umma.json; tail -3 gpurun_out/bench_umma.err
echo "=== bench (simt)"
STC_CONV_IMPL=1 timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_simt.json 2> gpurun_out/bench_simt.err; echo "rc=$?"; cat gpurun_out/bench_simt.json; tail -3 gpurun_out/bench_simt.err
echo "=== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 1 --batch 32 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "rc=$?"
echo "=== ncu full capture of the conv kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3x3_umma -s 20 -c 6 -o gpurun_out/prof_conv \
   python bench.py --steps 1 --warmup 1 --batch 32 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"
ls -la gpurun_out
