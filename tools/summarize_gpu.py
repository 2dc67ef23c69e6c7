"""Summarise gpurun_out/ (pytest log, bench JSON lines, ncu launch list)."""
import collections, csv, json, os, re, sys
D = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gpurun_out")
log = open(os.path.join(D, "pytest_gpu.log")).read() if os.path.exists(os.path.join(D, "pytest_gpu.log")) else ""
for l in log.splitlines():
    if re.search(r"passed|failed|^FAILED|mismatching|umma vs|golden|superres|out vs", l):
        print(l[:160])
for f in ("bench_umma", "bench_simt"):
    try:
        d = json.load(open(os.path.join(D, f + ".json")))
        print(f, "value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms/step", round(d["ms_per_step"], 2),
              "roofline", round(d["roofline"]["achieved"], 1), round(d["roofline"]["frac"], 4), "launches", d["gpu_launches"],
              d["roofline"].get("note","")[:60], d["clocks"], "cpu", d.get("cpu_baseline", {}).get("value"))
    except Exception as e:
        print(f, "n/a", e)
p = os.path.join(D, "launches.csv")
if os.path.exists(p):
    lines = [l for l in open(p) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = row["Kernel Name"].split("(")[0] + " grid=" + row["Grid Size"]
        v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]
        v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
        agg.setdefault(k, []).append(v)
    tot = sum(sum(v) for v in agg.values())
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print("%-78s n=%3d avg=%8.1f us tot=%9.1f %5.1f%%" % (k[:78], len(v), sum(v) / len(v), sum(v), 100 * sum(v) / tot))
    print("total us", round(tot, 1))
