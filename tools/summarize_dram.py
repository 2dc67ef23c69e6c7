"""Per-kernel table from an ncu CSV holding gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum:
launches, total time, DRAM bytes and achieved DRAM GB/s (bytes / duration).  Usage: summarize_dram.py file.csv [peak_GBps]"""
import csv, collections, re, sys


def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    per = collections.OrderedDict()
    for row in csv.DictReader(lines):
        name = re.sub(r"\(.*", "", row["Kernel Name"]).split("::")[-1]
        key = (row["ID"], name)
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        m = row["Metric Name"]
        if m == "gpu__time_duration.sum":
            v = v * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(unit.replace("second", "s") if unit in ("second",) else unit, 1e-9)
        else:
            v = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
        per.setdefault(key, {})[m] = v
    return per


def main():
    per = load(sys.argv[1])
    peak = float(sys.argv[2]) if len(sys.argv) > 2 else None
    agg = collections.OrderedDict()
    for (_, name), m in per.items():
        a = agg.setdefault(name, [0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += m.get("gpu__time_duration.sum", 0); a[2] += m.get("dram__bytes_read.sum", 0); a[3] += m.get("dram__bytes_write.sum", 0)
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | total ms | share | DRAM read MB | DRAM write MB | DRAM GB/s%s |" % (" | of peak" if peak else ""))
    print("|---|---|---|---|---|---|---|%s" % ("---|" if peak else ""))
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        gbs = (a[2] + a[3]) / a[1] / 1e9 if a[1] else 0
        line = "| %s | %d | %.3f | %.1f%% | %.1f | %.1f | %.0f |" % (name[:60], a[0], a[1] * 1e3, 100 * a[1] / tot, a[2] / 1e6, a[3] / 1e6, gbs)
        if peak:
            line += " %.0f%% |" % (100 * gbs / peak)
        print(line)


if __name__ == "__main__":
    main()
