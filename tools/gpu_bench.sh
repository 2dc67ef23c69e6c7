#!/bin/bash
# bench.py (both arms) on the GPU box; stdout lines -> gpurun_out/
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python bench.py --steps ${1:-10} --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -5 gpurun_out/bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"], d.get("cpu_baseline"))
print("roofline", d["roofline"]["frac"], d["roofline_hbm"]["frac"] if d.get("roofline_hbm") else None)
tc = d.get("tile_chain") or {}
for k in ("n12", "n24"):
    if k in tc:
        print(k, {a: b for a, b in tc[k].items() if a != "kernels"})
        print("  sum_kernel_ms", tc[k]["kernels"]["sum_kernel_ms"])
        for r in tc[k]["kernels"]["top"]:
            print("  ", r)
print("chain cpu", tc.get("cpu_baseline"))
PY
