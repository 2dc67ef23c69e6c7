"""Multi-GPU plumbing: tiles are independent, so the path shards with NO data-path
collective.  One process per GPU (torch.distributed, NCCL on
GPUs / gloo in CPU tests); the only collective is a one-time weight broadcast from
rank 0, plus the barrier / max-reduce bench.py uses for timing.
"""
import numpy as np


def shard_range(n_items, rank, world):
    """Contiguous block partition [lo, hi) of n_items for `rank` (row-bands of the patch grid)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def flatten_weights(w):
    names = sorted(w)
    flat = np.concatenate([np.asarray(w[k], np.float32).ravel() for k in names]) if names else np.zeros(0, np.float32)
    meta = [(k, tuple(np.asarray(w[k]).shape)) for k in names]
    return flat, meta


def unflatten_weights(flat, meta):
    out, o = {}, 0
    for k, shp in meta:
        n = int(np.prod(shp)) if len(shp) else 1
        out[k] = np.ascontiguousarray(flat[o:o + n].reshape(shp), np.float32)
        o += n
    return out


def broadcast_weights(w, dist, device=None, src=0):
    """Rank `src` holds `w` (dict); every rank returns the same dict.  Uses
    broadcast_object_list for the (name, shape) table and one tensor broadcast for the
    ~5.3 MB payload (NCCL over NVLink when tensors live on the GPU)."""
    import torch
    rank = dist.get_rank()
    if rank == src:
        flat, meta = flatten_weights(w)
    else:
        flat, meta = None, None
    box = [meta]
    dist.broadcast_object_list(box, src=src)
    meta = box[0]
    n = sum(int(np.prod(s)) if len(s) else 1 for _, s in meta)
    t = torch.from_numpy(flat.copy()) if rank == src else torch.empty(n, dtype=torch.float32)
    if device is not None:
        t = t.to(device)
    dist.broadcast(t, src=src)
    return unflatten_weights(t.cpu().numpy(), meta)
