#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/conv_ab.jsonl
fmt() { tail -1 gpurun_out/conv_ab.jsonl | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(' step %.2f ms (%.0f tiles/s) conv-sum %.2f ms | '%(d['ms_per_step'],d['tiles_per_s'],d['conv_ms_per_step'])+' '.join('%s %.0f'%(k,v['avg_us']) for k,v in d.items() if isinstance(v,dict) and 'avg_us' in v), '| checksum', d['checksum'])"; }
run2() { n=$1; shift; echo "--- $n: $*"; env "$@" timeout 300 python tools/conv_ab.py --batch 256 --steps 4 --tag "$*" --trace gpurun_out/trace_$n.csv >> gpurun_out/conv_ab.jsonl 2>> gpurun_out/conv_ab.err; echo "rc=$?"; fmt; python tools/trace_summary.py gpurun_out/trace_$n.csv; }
run2 single STC_SINGLE_STREAM=1 STC_CONV_PRIO=0
run2 s2p0 STC_SLOTS=2 STC_CONV_PRIO=0
run2 s2p1 STC_SLOTS=2 STC_CONV_PRIO=1
run2 s4p1 STC_SLOTS=4 STC_CONV_PRIO=1
run2 s4p0 STC_SLOTS=4 STC_CONV_PRIO=0
