#!/bin/bash
# A/B session on one GPU box: tools/conv_ab.py under different run-time switches, each with a
# kernel timeline.  Usage: gpurun -- 'bash tools/gpu_ab.sh'   (edit the list of variants below)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/conv_ab.jsonl
run2() { n=$1; shift; echo "--- $n: $*"; env "$@" timeout 300 python tools/conv_ab.py --batch 256 --steps 6 --tag "$*" --trace gpurun_out/trace_$n.csv >> gpurun_out/conv_ab.jsonl 2>> gpurun_out/conv_ab.err; echo "rc=$?"; tail -1 gpurun_out/conv_ab.jsonl | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(' step %.2f ms (%.0f tiles/s), conv launches sum %.2f ms, checksum %s'%(d['ms_per_step'],d['tiles_per_s'],d['conv_ms_per_step'],d['checksum']))"; python tools/trace_summary.py gpurun_out/trace_$n.csv | head -8; }
run2 single STC_SINGLE_STREAM=1 STC_CONV_PRIO=0
run2 default STC_SLOTS=4
run2 slots2 STC_SLOTS=2
run2 noprio STC_SLOTS=4 STC_CONV_PRIO=0
run2 v1 STC_CONV_V=1
