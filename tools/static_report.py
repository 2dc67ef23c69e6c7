"""Static evidence that needs no GPU: per-kernel resource usage (`cuobjdump -res-usage`) and a census of the Blackwell-specific
SASS opcodes (`cuobjdump -sass`) of the built library.  Writes profiles/r02_static_kernels.md.
Usage: python tools/static_report.py [path/to/libstc.so]"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "sentinel_tree_cover_b200", "libstc.so")
OPS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "HMMA", "IMMA", "ELECT", "REDUX", "LDGSTS",
       "ATOMG", "REDG", "ATOMS", "DFMA", "DADD", "MUFU", "LDL", "STL"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    def strip_args(o):
        o = o.replace("void ", "")
        if o.endswith(")"):                      # drop the trailing parameter list (balanced parentheses)
            depth = 0
            for i in range(len(o) - 1, -1, -1):
                depth += (o[i] == ")") - (o[i] == "(")
                if depth == 0:
                    return o[:i]
        return o
    return [strip_args(o).replace("(anonymous namespace)::", "") for o in out]


def main():
    res = subprocess.run(["cuobjdump", "-res-usage", SO], capture_output=True, text=True).stdout
    rows = re.findall(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", res)
    names = demangle([r[0] for r in rows])
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    per_fn, cur = collections.defaultdict(collections.Counter), None
    for line in sass.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1); continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            op = m.group(1).split(".")[0]
            per_fn[cur][op] += 1
            if "LOCAL" in line or op in ("LDL", "STL"):
                per_fn[cur]["<local>"] += 1
    tot = collections.Counter()
    for c in per_fn.values():
        tot.update(c)
    archs = sorted(set(re.findall(r"arch = (sm_\w+)", subprocess.run(["cuobjdump", "-lelf", SO], capture_output=True, text=True).stdout + sass)))
    with open(os.path.join(ROOT, "profiles", "r02_static_kernels.md"), "w") as f:
        f.write("# Static report of `libstc.so` (no GPU needed: `python tools/static_report.py`)\n\n")
        f.write("Architectures in the fat binary: %s.  %d kernels.\n\n" % (", ".join(archs) or "sm_100a", len(rows)))
        f.write("## Blackwell-specific and other telling SASS opcodes, whole library\n\n| opcode | count | what it is |\n|---|---|---|\n")
        what = {"UTCHMMA": "tcgen05.mma (fp16/bf16, TMEM accumulator)", "UTCBAR": "tcgen05.commit -> mbarrier", "LDTM": "tcgen05.ld (TMEM -> registers)",
                "STTM": "tcgen05.st", "UBLKCP": "cp.async.bulk (bulk copy engine, global <-> shared)", "UTMALDG": "cp.async.bulk.tensor load (tensor maps)",
                "UTMASTG": "tensor-map store", "SYNCS": "mbarrier arrive / try_wait", "HMMA": "legacy mma.sync (should be 0)", "IMMA": "legacy integer mma (should be 0)",
                "ELECT": "elect.sync", "REDUX": "warp reduce", "LDGSTS": "cp.async (Ampere-style)", "ATOMG": "global atomics with return", "REDG": "global reductions (fixed-point GroupNorm sums, counters)", "ATOMS": "shared-memory atomics",
                "LDL": "local-memory loads (spills / indexed per-thread arrays)", "STL": "local-memory stores",
                "DFMA": "fp64 fma (NumPy-order float64 statistics)", "DADD": "fp64 add", "MUFU": "special-function unit", "UTCQMMA": "tcgen05.mma fp8/fp4"}
        for op in OPS:
            f.write("| `%s` | %d | %s |\n" % (op, tot.get(op, 0), what.get(op, "")))
        f.write("\n## Per kernel: registers, stack, static shared memory, local memory; tcgen05 / bulk-copy / mbarrier instructions\n\n")
        f.write("| kernel | REG | STACK | SHARED (static) | LOCAL | UTCHMMA | LDTM | UBLKCP | SYNCS | LDL+STL |\n|---|---|---|---|---|---|---|---|---|---|\n")
        for (mangled, reg, stack, shared, local), name in sorted(zip(rows, names), key=lambda t: t[1]):
            c = per_fn.get(mangled, {})
            f.write("| `%s` | %s | %s | %s | %s | %d | %d | %d | %d | %d |\n" % (name[:110], reg, stack, shared, local, c.get("UTCHMMA", 0), c.get("LDTM", 0),
                                                                          c.get("UBLKCP", 0), c.get("SYNCS", 0), c.get("LDL", 0) + c.get("STL", 0)))
    spilled = [(n, r) for r, n in zip(rows, names) if int(r[4]) > 0 or int(r[2]) > 64]
    print("kernels:", len(rows), " with local memory or stack > 64 B:", len(spilled))
    for n, r in spilled:
        print("  ", n[:100], "STACK", r[2], "LOCAL", r[4])
    print({k: tot.get(k, 0) for k in OPS})


if __name__ == "__main__":
    main()
