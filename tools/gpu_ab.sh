#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/conv_ab.jsonl
run2() { n=$1; shift; echo "--- $n: $*"; env "$@" timeout 300 python tools/conv_ab.py --batch 256 --steps 6 --tag "$*" --trace gpurun_out/trace_$n.csv >> gpurun_out/conv_ab.jsonl 2>> gpurun_out/conv_ab.err; echo "rc=$?"; tail -1 gpurun_out/conv_ab.jsonl | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(' step %.2f ms'%d['ms_per_step'])"; python tools/trace_summary.py gpurun_out/trace_$n.csv | grep -E "span|apply"; }
run2 b256 STC_SINGLE_STREAM=1 STC_CONV_PRIO=0
run2 b128 STC_SINGLE_STREAM=1 STC_CONV_PRIO=0 STC_GRU_BLOCK=128
run2 b64 STC_SINGLE_STREAM=1 STC_CONV_PRIO=0 STC_GRU_BLOCK=64
run2 d256 STC_SLOTS=4
run2 d128 STC_SLOTS=4 STC_GRU_BLOCK=128
run2 d64 STC_SLOTS=4 STC_GRU_BLOCK=64
