#!/bin/bash
# two-rank sanity run of the default bench (tile chain included): worker threads of rank 1 must use GPU 1
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_n2.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_n2.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, "e2e", d["e2e"]["value"])
tc = d.get("tile_chain") or {}
for k in ("n12", "n24"):
    if k in tc: print(k, tc[k]["ms_per_tile"], tc[k]["tiles_per_s"], tc[k].get("tree_cover_mean"))
PY
