#!/bin/bash
# One GPU-box session: parity tests, smoke, bench, ncu launch list + one full capture of the
# dominant kernel.  Everything lands in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== tests"
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "golden|err|umma vs|passed|failed|FAILED|mismatch" gpurun_out/pytest_gpu.log | tail -30
echo "=== smoke"
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
echo "=== bench (tcgen05, dual stream)"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_umma.json 2> gpurun_out/bench_umma.err; echo "rc=$?"; cat gpurun_out/bench_umma.json; tail -3 gpurun_out/bench_umma.err
echo "=== bench (tcgen05, single stream)"
STC_SINGLE_STREAM=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_single.json 2> gpurun_out/bench_single.err; echo "rc=$?"; cat gpurun_out/bench_single.json; tail -3 gpurun_out/bench_single.err
if [ -n "$STC_RUN_EXP" ]; then
echo "=== experiment: aligned A taps (timing only, results invalid)"
STC_SINGLE_STREAM=1 STC_EXP_ALIGN=1 timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_exp_align.json 2> gpurun_out/bench_exp_align.err; echo "rc=$?"; cat gpurun_out/bench_exp_align.json
fi
echo "=== ncu launch list (single stream so kernels are attributed cleanly)"
STC_SINGLE_STREAM=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 1 --batch 32 --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "rc=$?"
echo "=== ncu full capture of the conv kernel"
STC_SINGLE_STREAM=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3x3_umma -s 20 -c 6 -o gpurun_out/prof_conv \
   python bench.py --steps 1 --warmup 1 --batch 32 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"
ls -la gpurun_out
