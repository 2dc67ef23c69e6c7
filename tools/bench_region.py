"""CLI of the region-scale run (sentinel_tree_cover_b200/region_bench.py; also reachable as `bench.py --config region`).
  python tools/bench_region.py [--rows 190 --cols 190] [--periodic] [--verify]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_region.py ..."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=190)
    ap.add_argument("--cols", type=int, default=190)
    ap.add_argument("--patch", type=int, default=168)
    ap.add_argument("--stride", type=int, default=58)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--periodic", action="store_true", help="round-1 input: one 232-px cube addressed modulo its size")
    ap.add_argument("--verify", action="store_true", help="rank 0 checks its first canvas rows against the NumPy oracle blend")
    a = ap.parse_args()
    from sentinel_tree_cover_b200 import region_bench

    def verify(first_rows, canvas_u8, stride):
        """rank 0's first patch rows blended by the NumPy oracle: the canvas rows they fully determine must be identical"""
        import numpy as np
        from oracle import region_ref as RR
        nv = first_rows.shape[0]
        want = RR.blend_region(first_rows, stride)
        rows_ok = (nv - 2) * stride if nv > 2 else 0
        return {"rows_checked": rows_ok, "bit_exact_vs_oracle": bool(rows_ok == 0 or np.array_equal(canvas_u8[:rows_ok], want[:rows_ok]))}
    region_bench.run(a.rows, a.cols, a.patch, a.stride, a.batch, periodic=a.periodic, verify_fn=verify if a.verify else None)
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
