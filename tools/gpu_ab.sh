#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q > gpurun_out/pytest_model.log 2>&1; echo pytest rc=$?; tail -1 gpurun_out/pytest_model.log
: > gpurun_out/conv_ab.jsonl
fmt() { tail -1 gpurun_out/conv_ab.jsonl | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(' step %.2f ms (%.0f tiles/s) conv-sum %.2f ms | checksum %s'%(d['ms_per_step'],d['tiles_per_s'],d['conv_ms_per_step'],d['checksum']))"; }
run2() { echo "--- $*"; env "$@" timeout 300 python tools/conv_ab.py --batch 256 --steps 6 --tag "$*" >> gpurun_out/conv_ab.jsonl 2>> gpurun_out/conv_ab.err; echo "rc=$?"; fmt; }
run2 STC_SLOTS=4
run2 STC_SINGLE_STREAM=1 STC_CONV_PRIO=0
run2 STC_SLOTS=4 STC_CONV_SMS=100 STC_ELEM_SMEM_KB=66
run2 STC_SLOTS=4 STC_CONV_SMS=116 STC_ELEM_SMEM_KB=66
run2 STC_SLOTS=4 STC_CONV_SMS=84 STC_ELEM_SMEM_KB=66
run2 STC_SLOTS=4 STC_CONV_SMS=100 STC_ELEM_SMEM_KB=66 STC_CONV_PRIO=0
run2 STC_SLOTS=2 STC_CONV_SMS=100 STC_ELEM_SMEM_KB=66
run2 STC_SLOTS=4 STC_CONV_SMS=100 STC_ELEM_SMEM_KB=0
