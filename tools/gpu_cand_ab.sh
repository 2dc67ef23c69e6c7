#!/bin/bash
# A/B: one vs two CTAs per SM for the GRU candidate convolution (bench value + conv times from the traced step)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for occ in 1 2 1 2; do
  STC_CAND_OCC=$occ timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-tile-chain > gpurun_out/bench_cand$occ.json 2> gpurun_out/bench_cand$occ.err
  python - <<PY
import json, csv
d = json.loads(open("gpurun_out/bench_cand$occ.json").read().strip().splitlines()[-1])
t = {}
for r in csv.DictReader(open("gpurun_out/bench_trace_rank0.csv")):
    e = t.setdefault(r["label"], [0, 0.0]); e[0] += 1; e[1] += float(r["end_ms"]) - float(r["start_ms"])
print("occ $occ value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 2), "checksum", d.get("checksum"), "conv_cand", t.get("conv_cand"), "conv_gates", t.get("conv_gates"))
PY
done
STC_CAND_OCC=2 timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -2
