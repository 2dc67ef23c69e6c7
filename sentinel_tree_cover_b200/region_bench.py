"""Region-scale run (BASELINE configs[3]): R x C overlapping 168-px patches every 58 px (190 x 190 =
36,100 patches over an 11,130 px canvas for 1 x 1 degree), patch rows sharded over the GPUs of one box, Gaussian
overlap-blend mosaic -> uint8 canvas on rank 0.  Reached as `python bench.py --config region [--gpus N]` (and under
torch.distributed.run for N > 1) or through tools/bench_region.py.

Input: every rank holds ITS band of the canvas, [12, band rows, 11130, 13] float32, resident in HBM (77 GB for the whole
canvas at N = 1, ~10 GB per rank at N = 8) -- synthetic content (a seeded 232-px cube repeated periodically, written by the
GPU), real memory: the window gather streams it from HBM, nothing is L2-resident.  `periodic=True` keeps round 1's variant
(one 33 MB cube addressed modulo its size) for boxes without the memory.

One JSON line on rank 0 in bench.py's schema: `value` = tiles/s over gather + forward + halo exchange + blend + collection of
the canvas bands (wall clock between barriers, max over ranks; strong scaling: the region is fixed)."""
import json
import os
import time

import numpy as np


def run(rows=190, cols=190, patch=168, stride=58, batch=256, periodic=False, verify_fn=None, steps=1, warmup=1, emit=True):
    import torch
    from .api import StcSession
    from .weights import random_predict_weights
    from .shard import broadcast_weights, shard_range
    from . import region
    from . import synth as P
    rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w = random_predict_weights(0) if rank == 0 else None
    if world > 1:
        w = broadcast_weights(w, dist, device=torch.device("cuda", local))
    sess = StcSession(local, predict_weights=w)
    R, C = rows, cols
    S = patch - 14
    rr = region.RegionRunner(sess, R, C, patch, stride, rank, world, batch=batch)
    spans = [shard_range(R, r, world) for r in range(world)]
    base = torch.from_numpy(np.ascontiguousarray(P.synth_monthly(1, 232, 4000)[0])).cuda()         # [12, 232, 232, 13]
    first = region.halo_rows(rr.ra, S, stride, 7)
    n_halo = rr.ra - first
    if periodic:
        canvas, Hband, Wband, wrap, band_y0 = base, 232, 232, True, 0
    else:
        band_y0 = rr.ra * stride
        Hband = max((rr.rb - rr.ra - 1) * stride + patch, 0) if rr.rb > rr.ra else 1
        Wband = rr.Wc
        ys = (torch.arange(Hband, device="cuda") + band_y0) % 232
        xs = torch.arange(Wband, device="cuda") % 232
        canvas = torch.empty((12, Hband, Wband, 13), dtype=torch.float32, device="cuda")
        for t in range(12):                                   # one month at a time: the temporary is a row band, not the canvas
            canvas[t] = base[t][ys][:, xs]
        wrap = False
    preds = torch.zeros((n_halo + (rr.rb - rr.ra), C, S, S), dtype=torch.float32, device="cuda")
    own_ptr = preds.data_ptr() + n_halo * C * S * S * 4

    def barrier():
        sess.sync(); torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    def one_pass():
        t0 = time.perf_counter()
        rr.predict_rows(canvas.data_ptr(), 12, Hband, Wband, 13, wrap, band_y0, own_ptr)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        if world > 1:
            tail = torch.zeros((2, C, S, S), dtype=torch.float32, device="cuda")
            k = min(2, rr.rb - rr.ra)
            if k:
                tail[2 - k:] = preds[preds.shape[0] - k:]
            halo = region.exchange_halo(tail, dist, rank, world, spans=spans, need=(first, rr.ra))
            if n_halo:
                preds[:n_halo] = halo
        torch.cuda.synchronize(); t2 = time.perf_counter()
        band, (y0, y1) = rr.blend(preds.data_ptr(), first, preds.shape[0])
        t3 = time.perf_counter()
        if world > 1:
            # bands differ in height (the last rank takes the remainder): pad to the tallest, one NCCL all_gather of uint8
            sp = [region.owned_canvas_rows(a, b, R, patch, stride) for a, b in spans]
            hmax = max(b - a for a, b in sp)
            mine = torch.zeros((hmax, rr.Wc), dtype=torch.uint8, device="cuda")
            mine[:band.shape[0]] = torch.from_numpy(band).cuda()
            allb = torch.empty((world, hmax, rr.Wc), dtype=torch.uint8, device="cuda")
            dist.all_gather_into_tensor(allb, mine)
            out = np.concatenate([allb[r, :sp[r][1] - sp[r][0]].cpu().numpy() for r in range(world)]) if rank == 0 else None
        else:
            out = band
        t4 = time.perf_counter()
        return out, [t4 - t0, t1 - t0, t2 - t1, t3 - t2, t4 - t3]

    # warm-up: one batch through the same path (allocations, first launches)
    warm = region.RegionRunner(sess, 1, min(C, batch), patch, stride, 0, 1, batch=batch)
    tmp = torch.empty((1, min(C, batch), S, S), dtype=torch.float32, device="cuda")
    for _ in range(max(1, warmup)):
        warm.predict_rows(base.data_ptr(), 12, 232, 232, 13, True, 0, tmp.data_ptr())
    barrier()
    l0 = sess.launch_count()
    all_t = []
    canvas_u8 = None
    for _ in range(max(1, steps)):
        barrier()
        canvas_u8, tt = one_pass()
        all_t.append(tt)
    launches = sess.launch_count() - l0
    times = torch.tensor(all_t, dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    line = None
    if rank == 0:
        tt = times.mean(0).cpu().tolist()
        line = {"metric": "tiles/sec (12-step 168x168x13 S1+S2 patches)", "value": R * C / tt[0], "unit": "tiles/s", "n_gpus": world,
                "steps": max(1, steps), "warmup": max(1, warmup), "ms_per_step": 1e3 * tt[0], "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f16", "data": "synthetic",
                "config": {"workload": "configs[3]: 1x1 degree region, %d x %d patches of %d px every %d px, sharded by patch rows, "
                                       "Gaussian overlap-blend mosaic to a uint8 canvas" % (R, C, patch, stride),
                           "tiles": R * C, "canvas": [rr.Hc, rr.Wc], "batch_per_gpu": batch,
                           "input": ("periodic 232-px cube (33 MB, L2-resident)" if periodic else
                                     "per-rank canvas band [12, rows, %d, 13] float32 resident in HBM (%.1f GB on rank 0)"
                                     % (rr.Wc, canvas.numel() * 4 / 1e9)),
                           "weights": "random-init, released architecture", "parallelism": "patch rows sharded, one NCCL weight broadcast, "
                           "one all_gather of 2 halo rows, one all_gather of the uint8 bands"},
                "phases_s": {"gather+forward": tt[1], "halo exchange": tt[2], "blend": tt[3], "canvas gather": tt[4]},
                "gpu_launches": int(launches), "canvas_checksum": int(canvas_u8.astype(np.int64).sum()),
                "nodata_px": int((canvas_u8 == 255).sum()), "guard_px": int((canvas_u8 == 254).sum()),
                "e2e": {"value": R * C / tt[0], "unit": "tiles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": int(canvas_u8.nbytes),
                        "note": "the canvas band is resident (its upload is a one-off per region); the uint8 mosaic comes back to the host"}}
        if verify_fn is not None:                          # the caller's checker (tools/bench_region.py --verify brings the oracle)
            nv = min(6, rr.rb - rr.ra)
            line["verify"] = verify_fn(preds[n_halo:n_halo + nv].cpu().numpy(), canvas_u8, stride)
        if emit:
            print(json.dumps(line), flush=True)
    sess.close()
    if dist is not None:
        dist.barrier()
    return line
