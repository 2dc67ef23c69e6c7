#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "=== full gpu suite"
timeout 1200 python -m pytest tests -m gpu -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "out vs|golden [ab]|passed|failed|umma vs|FAILED" gpurun_out/pytest_gpu.log | head -20
: > gpurun_out/conv_ab.jsonl
fmt() { tail -1 gpurun_out/conv_ab.jsonl | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(' step %.2f ms (%.0f tiles/s) conv-sum %.2f ms | '%(d['ms_per_step'],d['tiles_per_s'],d['conv_ms_per_step'])+' '.join('%s %.0f'%(k,v['avg_us']) for k,v in d.items() if isinstance(v,dict) and 'avg_us' in v), '| checksum', d['checksum'])"; }
run2() { echo "--- $*"; env "$@" timeout 300 python tools/conv_ab.py --batch 256 --steps 4 --tag "dual $*" >> gpurun_out/conv_ab.jsonl 2>> gpurun_out/conv_ab.err; echo "rc=$?"; fmt; }
run2 STC_SINGLE_STREAM=1
run2 STC_SLOTS=2
run2 STC_SLOTS=3
run2 STC_SLOTS=4
run2 STC_SLOTS=4 STC_CHUNK=16
run2 STC_SLOTS=4 STC_CHUNK=64
run2 STC_SLOTS=3 STC_CONV_PRIO=0
run2 STC_SLOTS=4 STC_CONV_SMEM_KB=131
echo "=== ncu launch list (single stream)"
STC_SINGLE_STREAM=1 STC_CONV_PRIO=0 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv \
   python tools/conv_ab.py --batch 32 --steps 1 > gpurun_out/ncu_list.log 2>&1; echo "rc=$?"
