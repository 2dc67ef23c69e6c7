"""Derivation and proof of the median-of-12 selection network in csrc/stc_indices.cuh (median12_net):
Batcher's odd-even mergesort for 16 inputs, comparators touching inputs 12..15 dropped (they are +inf), then pruned
backwards from outputs 5 and 6; verified with the 0-1 principle on all 4096 binary inputs."""
import itertools


def batcher(n):
    pairs = []

    def merge(lo, n, r):
        m = r * 2
        if m < n:
            merge(lo, n, m); merge(lo + r, n, m)
            for i in range(lo + r, lo + n - r, m):
                pairs.append((i, i + r))
        else:
            pairs.append((lo, lo + r))

    def sort(lo, n):
        if n > 1:
            m = n // 2
            sort(lo, m); sort(lo + m, m); merge(lo, n, 1)
    sort(0, n)
    return pairs


def prune(net, outs):
    need, keep = set(outs), []
    for a, b in reversed(net):
        if a in need or b in need:
            keep.append((a, b)); need.add(a); need.add(b)
    return list(reversed(keep))


def run(net, v):
    v = list(v)
    for a, b in net:
        if v[a] > v[b]:
            v[a], v[b] = v[b], v[a]
    return v


if __name__ == "__main__":
    net = [(a, b) for a, b in batcher(16) if a < 12 and b < 12]
    sel = prune(net, {5, 6})
    dedup = [c for i, c in enumerate(sel) if i == 0 or c != sel[i - 1]]
    assert all(run(dedup, bits)[5:7] == sorted(bits)[5:7] for bits in itertools.product([0, 1], repeat=12))
    print(len(dedup), "comparators:", dedup)
