#!/bin/bash
# N-GPU runs of both bench configurations (N = $1)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --config region --steps 1 --warmup 1 > gpurun_out/region_n$N.json 2> gpurun_out/region_n$N.err; echo "region rc=$?"
tail -2 gpurun_out/region_n$N.err; tail -1 gpurun_out/region_n$N.json | cut -c1-1500
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"
tail -2 gpurun_out/bench_n$N.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_n$N.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"].get("h2d_gbs_per_rank"), "e2e_f32", d["e2e_f32"]["value"])
tc = d.get("tile_chain") or {}
for k in ("n12", "n24"):
    if k in tc: print(k, tc[k]["ms_per_tile"], tc[k]["tiles_per_s"])
PY
nvidia-smi topo -m 2>/dev/null | head -14
