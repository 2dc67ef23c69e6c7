#!/bin/bash
# model step vs STC_CHUNK (tiles per sub-batch) and the write-combined staging experiment
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for c in 32 16 8; do
  STC_CHUNK=$c timeout 600 python bench.py --steps 10 --warmup 3 --no-tile-chain --no-cpu-baseline > gpurun_out/bench_chunk$c.json 2> gpurun_out/bench_chunk$c.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_chunk$c.json").read().strip().splitlines()[-1])
print("chunk $c: ms/step %.2f value %.0f e2e(u16) %.0f ms %.1f e2e_f32 %.0f gates frac %.3f (in region %.3f) hbm frac %.3f launches %d" % (
    d["ms_per_step"], d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e_f32"]["value"], d["roofline"]["frac"],
    d["roofline"].get("frac_in_timed_region", 0), d["roofline_hbm"]["frac"], d["gpu_launches"]))
PY
done
STC_BENCH_WC=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-tile-chain --no-cpu-baseline > gpurun_out/bench_wc.json 2> gpurun_out/bench_wc.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_wc.json").read().strip().splitlines()[-1])
print("write-combined u16 staging: e2e %.0f tiles/s, %.1f ms, %.1f GB/s" % (d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["h2d_gbs_per_rank"]))
PY
