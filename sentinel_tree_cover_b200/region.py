"""Region-scale driver: a region is an R x C grid of overlapping patches (168 px
windows every 58 px: 190 x 190 patches over an 11,130 px canvas for 1 x 1 degree).  Rows of the patch grid are
block-partitioned over the ranks (shard.shard_range); every rank cuts its patches out of its device-resident canvas
band, runs the batched forward, and blends the canvas rows it owns.  The only data-path exchange is the halo: a canvas
row is covered by up to three patch rows, so a rank needs the LAST TWO patch rows of the previous rank (their
probabilities, 2 x C x S x S floats) -- one all_gather of those blocks.  The blend adds the covering patches of a
pixel in grid order, so the mosaic is bit-identical for any number of ranks.

Host side only (integers, torch.distributed plumbing); arithmetic is in csrc/stc_region.cu and the model kernels."""
import ctypes as _C

import numpy as np

from . import api as _api
from .shard import shard_range


def _dev(p, offset=0):
    """Device pointer (int or ctypes.c_void_p) + byte offset -> ctypes.c_void_p."""
    base = p.value if isinstance(p, _C.c_void_p) else int(p)
    return _C.c_void_p(base + int(offset))


def canvas_size(n, patch, stride):
    return (n - 1) * stride + patch


def owned_canvas_rows(ra, rb, R, patch, stride):
    """Canvas rows blended by the rank that owns patch rows [ra, rb): one stride band per patch row, the last rank
    takes the remainder."""
    end = canvas_size(R, patch, stride)
    if ra >= rb:                      # more ranks than patch rows: the trailing ranks own nothing
        y = end if ra >= R else ra * stride
        return y, y
    return ra * stride, (rb * stride if rb < R else end)


def halo_rows(ra, S, stride, margin):
    """First patch row whose output can touch canvas row ra*stride (rows below come from the previous rank)."""
    a = ra * stride - margin - S
    first = (a // stride) + 1
    return max(first, 0)


class RegionRunner:
    """One rank's share of a region.  `canvas_dev` is a device pointer to the rank's canvas band [T, Hband, Wc, Cc]
    float32 whose first row is canvas row `band_y0` (wrap=True: a periodic base cube instead, origin (0, 0))."""

    def __init__(self, sess, R, C, patch=168, stride=58, rank=0, world=1, batch=256, sigma=36):
        self.sess, self.R, self.C, self.P, self.stride = sess, R, C, patch, stride
        self.S, self.margin = patch - 14, 7
        self.rank, self.world, self.batch, self.sigma = rank, world, batch, sigma
        self.ra, self.rb = shard_range(R, rank, world)
        self.Hc, self.Wc = canvas_size(R, patch, stride), canvas_size(C, patch, stride)

    # ---- forward over the rank's patch rows -------------------------------------------------
    def predict_rows(self, canvas_dev, T, Hband, Wband, Cc, wrap, band_y0, preds_dev):
        """Fills preds_dev [(rb-ra), C, S, S] float32 (device pointer)."""
        sess, P, S = self.sess, self.P, self.S
        n = (self.rb - self.ra) * self.C
        if n == 0:
            return 0
        idx = np.arange(n)
        ys = ((self.ra + idx // self.C) * self.stride - band_y0).astype(np.int32)
        xs = ((idx % self.C) * self.stride).astype(np.int32)
        if not wrap and (ys.min() < 0 or xs.min() < 0 or ys.max() + P > Hband or xs.max() + P > Wband):
            raise ValueError("region: a patch window leaves the canvas band")
        # window origins of ALL batches go to the device once, so the batches are enqueued back to back and the host runs
        # ahead of the GPU (a host hiccup between batches would otherwise idle the device)
        coords = np.ascontiguousarray(np.concatenate([ys, xs]))
        d_xy = sess.malloc(coords.nbytes)
        try:
            sess.h2d(d_xy, coords)
            mn, mnp = _api._f64(sess.min_all)
            mx, mxp = _api._f64(sess.max_all)
            # one call: per batch a gather on a side stream (overlapping the previous batch's forward) + the forward
            sess._check(sess.lib.stc_region_predict_dev(sess.h, _dev(canvas_dev), T, Hband, Wband, Cc, int(bool(wrap)),
                                                        _dev(d_xy), _dev(d_xy, n * 4), n, self.batch, P, mnp, mxp, _dev(preds_dev)))
        finally:
            sess.free(d_xy)
        return n

    # ---- blend of the owned canvas rows -------------------------------------------------------
    def blend(self, preds_dev, r_first, rows_have):
        y0, y1 = owned_canvas_rows(self.ra, self.rb, self.R, self.P, self.stride)
        out = np.empty((y1 - y0, self.Wc), np.uint8)
        if y1 <= y0:
            return out, (y0, y1)
        gauss = np.ascontiguousarray(_api.fspecial_gauss(self.S, self.sigma), np.float32)
        self.sess._check(self.sess.lib.stc_region_blend_dev(self.sess.h, _dev(preds_dev), int(r_first), int(rows_have), self.R, self.C, self.S,
                                                            self.stride, self.margin, _api._dptr(gauss), y0, y1, self.Wc, _api._dptr(out)))
        return out, (y0, y1)


def exchange_halo(own_tail, dist, rank, world, spans=None, need=None):
    """own_tail: torch tensor [2, C, S, S] = the rank's LAST two patch rows (row rb-2 in slot 0, rb-1 in slot 1; a rank that
    owns a single row puts it in slot 1, a rank that owns none sends zeros).  One all_gather (NCCL on GPUs, gloo in the CPU
    tests).  With `spans` = [(ra, rb)] of every rank and `need` = (first, ra), returns a tensor [ra - first, C, S, S] holding
    the patch rows first .. ra-1 picked by GLOBAL row index from whichever rank owns them (the previous rank may own fewer
    than two rows when there are more ranks than half the patch rows); raises if a needed row is in nobody's tail.
    Without `spans` (old call): the previous rank's tail as is (None on rank 0)."""
    import torch
    if world == 1:
        return None
    got = [torch.empty_like(own_tail) for _ in range(world)]
    dist.all_gather(got, own_tail)
    if spans is None:
        return got[rank - 1] if rank > 0 else None
    first, ra = need
    rows = []
    for g in range(first, ra):
        owner = next((r for r, (a, b) in enumerate(spans) if a <= g < b), None)
        if owner is None or g < spans[owner][1] - 2:
            raise RuntimeError("region halo: patch row %d is not among the last two rows of its owner (rank %r)" % (g, owner))
        rows.append(got[owner][2 - (spans[owner][1] - g)])
    return torch.stack(rows) if rows else own_tail[:0]
