#!/bin/bash
# One GPU-box session: parity tests (SIMT verification kernel first, then the tcgen05 path),
# smoke, a short bench and the ncu launch list.  Everything lands in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
echo "=== SIMT impl tests" 
STC_CONV_IMPL=1 timeout 600 python -m pytest tests -m gpu -q -s -k "not umma and not tcgen05" > gpurun_out/pytest_simt.log 2>&1; echo "simt rc=$?"
tail -5 gpurun_out/pytest_simt.log
echo "=== UMMA impl tests"
timeout 900 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_umma.log 2>&1; echo "umma rc=$?"
tail -15 gpurun_out/pytest_umma.log
echo "=== smoke"
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
echo "=== bench (simt)"
STC_CONV_IMPL=1 timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_simt.json 2> gpurun_out/bench_simt.err; echo "rc=$?"; cat gpurun_out/bench_simt.json; tail -3 gpurun_out/bench_simt.err
echo "=== bench (umma)"
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_umma.json 2> gpurun_out/bench_umma.err; echo "rc=$?"; cat gpurun_out/bench_umma.json; tail -3 gpurun_out/bench_umma.err
