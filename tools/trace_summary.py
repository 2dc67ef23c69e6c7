"""Summarise a kernel timeline written by stc_trace (tools/conv_ab.py --trace): busy time per kernel class,
time with both classes active, idle time, and per-label durations.  Start stamps are stream-head times, so
`dur` is an upper bound on a kernel's own run time."""
import csv
import sys
from collections import defaultdict


def union(iv):
    iv = sorted(iv); out = []
    for a, b in iv:
        if out and a <= out[-1][1]:
            out[-1][1] = max(out[-1][1], b)
        else:
            out.append([a, b])
    return out


def length(iv):
    return sum(b - a for a, b in iv)


def intersect(x, y):
    i = j = 0; out = []
    while i < len(x) and j < len(y):
        a, b = max(x[i][0], y[j][0]), min(x[i][1], y[j][1])
        if a < b:
            out.append([a, b])
        if x[i][1] < y[j][1]:
            i += 1
        else:
            j += 1
    return out


def main(path):
    rows = list(csv.DictReader(open(path)))
    conv, elem, per = [], [], defaultdict(list)
    for r in rows:
        a, b = float(r["start_ms"]), float(r["end_ms"])
        (conv if r["label"].startswith("conv") else elem).append((a, b))
        per[r["label"]].append(b - a)
    t0 = min(float(r["start_ms"]) for r in rows); t1 = max(float(r["end_ms"]) for r in rows)
    uc, ue = union(conv), union(elem)
    both = intersect(uc, ue)
    anyb = union(conv + elem)
    print("span %.2f ms | conv busy %.2f | elementwise busy %.2f | both %.2f | neither %.2f" %
          (t1 - t0, length(uc), length(ue), length(both), (t1 - t0) - length(anyb)))
    print("sum of durations: conv %.2f, elementwise %.2f" % (sum(b - a for a, b in conv), sum(b - a for a, b in elem)))
    for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
        print("  %-12s n=%4d  sum %7.2f ms  mean %6.1f us  max %6.1f us" % (k, len(v), sum(v), 1e3 * sum(v) / len(v), 1e3 * max(v)))


if __name__ == "__main__":
    main(sys.argv[1])
