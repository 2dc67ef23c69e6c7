"""Region-scale driver: grid bookkeeping, halo exchange (world_size 2 over gloo, CPU),
device gather and Gaussian overlap blend (GPU, bit-exact against oracle/region_ref.py)."""
import os
import subprocess
import sys
import textwrap
import numpy as np
import pytest
from oracle import region_ref as RR
from oracle import preproc_ref as P
from sentinel_tree_cover_b200 import region
from sentinel_tree_cover_b200.shard import shard_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_row_ownership_and_halo_cover_every_canvas_row():
    for (R, patch, stride) in ((190, 168, 58), (7, 44, 16), (3, 44, 16), (1, 44, 16)):
        S, margin = patch - 14, 7
        for world in (1, 2, 3, 8):
            rows = []
            for rank in range(world):
                ra, rb = shard_range(R, rank, world)
                y0, y1 = region.owned_canvas_rows(ra, rb, R, patch, stride)
                rows.append((y0, y1))
                first = region.halo_rows(ra, S, stride, margin)
                assert ra - first <= 2                                     # at most two rows from the previous rank
                for y in range(y0, y1):                                    # every covering patch row is available
                    cover = [r for r in range(R) if r * stride + margin <= y < r * stride + margin + S]
                    assert all(first <= r < rb for r in cover), (R, world, rank, y, cover, first, ra, rb)
            assert rows[0][0] == 0 and rows[-1][1] == region.canvas_size(R, patch, stride)
            assert all(rows[i][1] == rows[i + 1][0] for i in range(world - 1))


def test_blend_oracle_small_known_case():
    """One patch: the canvas is the patch's x100 truncation inside the footprint, 255 outside; two identical
    overlapping patches blend to the same values."""
    S, stride = 30, 16
    p = np.linspace(0.2, 0.9, S * S, dtype=np.float32).reshape(S, S)
    out = RR.blend_region(p[None, None], stride)
    assert out.shape == (44, 44) and (out[:7] == 255).all() and (out[:, -7:] == 255).all()
    inner = out[7:-7, 7:-7].astype(np.int32)
    assert np.abs(inner - np.floor(p * 100)).max() <= 1          # w*v/w may round one step off the direct truncation
    two = RR.blend_region(np.stack([p[None], p[None]]), stride)        # R=2, C=1, shifted copies
    assert two.shape == (60, 44) and (two[7:23, 7:-7] == out[7:23, 7:-7]).all()


_WORKER = textwrap.dedent('''
    import os, sys
    sys.path.insert(0, %(root)r)
    import numpy as np, torch, torch.distributed as dist
    from oracle import region_ref as RR
    from sentinel_tree_cover_b200 import region
    from sentinel_tree_cover_b200.shard import shard_range
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    R, C, patch, stride = %(R)d, 3, 44, 16
    S, margin = patch - 14, 7
    full = np.random.default_rng(5).uniform(0.05, 0.95, (R, C, S, S)).astype(np.float32)
    ra, rb = shard_range(R, rank, world)
    own = full[ra:rb]
    tail = np.zeros((2, C, S, S), np.float32)
    k = min(2, rb - ra)
    if k: tail[2 - k:] = own[-k:]
    first = region.halo_rows(ra, S, stride, margin)
    spans = [shard_range(R, r, world) for r in range(world)]
    halo = region.exchange_halo(torch.from_numpy(tail), dist, rank, world, spans=spans, need=(first, ra))
    have = own if halo is None or ra == first else np.concatenate([halo.numpy(), own])
    # blend the owned rows from `have` only: embed into a zero grid (rows outside [first, rb) never touch owned rows)
    grid = np.zeros_like(full); grid[first:rb] = have
    y0, y1 = region.owned_canvas_rows(ra, rb, R, patch, stride)
    mine = RR.blend_region(grid, stride, rows=(y0, y1))
    want = RR.blend_region(full, stride, rows=(y0, y1))
    assert np.array_equal(mine, want), (rank, int((mine != want).sum()))
    bands = [None] * world
    dist.all_gather_object(bands, mine)
    if rank == 0:
        assert np.array_equal(np.concatenate(bands), RR.blend_region(full, stride))
        print("REGION-OK")
    dist.destroy_process_group()
''')


@pytest.mark.parametrize("world,R", [(2, 7), (3, 4)], ids=["world2", "world3_single_row_ranks"])
def test_halo_exchange_gloo(tmp_path, world, R):
    """world 3 over 4 patch rows: shards [0,2) [2,3) [3,4) -- the last rank needs rows 1 and 2, which sit in the tails of TWO
    different ranks (the previous rank owns a single row): the halo is assembled by global row index."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % {"root": ROOT, "R": R})
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
                        "--master-port", str(29571 + world), str(script)], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and "REGION-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.gpu
def test_region_gather_matches_oracle(sess):
    r = np.random.default_rng(3)
    base = r.uniform(0, 1, (12, 96, 80, 13)).astype(np.float32)
    d = sess.malloc(base.nbytes); sess.h2d(d, base)
    ys = np.array([0, 5, 90, 60], np.int32); xs = np.array([0, 7, 70, 75], np.int32)
    xy = np.concatenate([ys, xs]); dxy = sess.malloc(xy.nbytes); sess.h2d(dxy, xy)
    out = np.empty((4, 12, 44, 44, 13), np.float32)
    do = sess.malloc(out.nbytes)
    sess._check(sess.lib.stc_region_gather_dev(sess.h, d, 12, 96, 80, 13, 1, dxy, region._dev(dxy, 16), 4, 44, do))
    sess.d2h(out, do); sess.sync()
    for b in range(4):
        want = np.stack([base[t][np.ix_(np.arange(ys[b], ys[b] + 44) % 96, np.arange(xs[b], xs[b] + 44) % 80)] for t in range(12)])
        assert np.array_equal(out[b], want)
    # without wrap: in-range windows are plain crops; the Python driver refuses windows that leave the band
    ys2 = np.array([0, 52], np.int32); xs2 = np.array([36, 0], np.int32)
    xy2 = np.concatenate([ys2, xs2]); sess.h2d(dxy, xy2)
    sess._check(sess.lib.stc_region_gather_dev(sess.h, d, 12, 96, 80, 13, 0, dxy, region._dev(dxy, 8), 2, 44, do))
    sess.d2h(out, do); sess.sync()
    for b in range(2):
        assert np.array_equal(out[b], base[:, ys2[b]:ys2[b] + 44, xs2[b]:xs2[b] + 44])
    rr = region.RegionRunner(sess, 3, 3, 44, 16)
    with pytest.raises(ValueError):
        rr.predict_rows(d, 12, 60, 80, 13, False, 0, do)                 # rows 2*16 + 44 = 76 > 60
    sess.free(dxy)
    sess.free(d); sess.free(do)


@pytest.mark.gpu
@pytest.mark.parametrize("world", [1, 3])
def test_region_blend_bit_exact(sess, world):
    R, C, patch, stride = 7, 4, 44, 16
    S = patch - 14
    full = np.random.default_rng(8).uniform(0.0, 1.0, (R, C, S, S)).astype(np.float32)
    full[2, 1, :10, :10] = 0.05                                    # below the 15 % floor
    want = RR.blend_region(full, stride)
    bands = []
    for rank in range(world):
        rr = region.RegionRunner(sess, R, C, patch, stride, rank, world)
        first = region.halo_rows(rr.ra, S, stride, 7)
        have = np.ascontiguousarray(full[first:rr.rb])
        d = sess.malloc(have.nbytes); sess.h2d(d, have)
        out, (y0, y1) = rr.blend(d, first, have.shape[0])
        sess.free(d)
        assert np.array_equal(out, want[y0:y1]), (rank, int((out != want[y0:y1]).sum()))
        bands.append(out)
    assert np.array_equal(np.concatenate(bands), want)


@pytest.mark.gpu
def test_region_end_to_end_small(sess, predict_weights):
    """4 x 3 grid of 44-px patches cut from a periodic base cube: gather -> forward -> blend on the device."""
    from oracle.model_ref import PredictRef
    from sentinel_tree_cover_b200.api import MIN_ALL, MAX_ALL
    R, C, patch, stride = 4, 3, 44, 16
    S = patch - 14
    base = P.synth_monthly(1, 64, 9)[0]                              # [12, 64, 64, 13], used periodically
    d = sess.malloc(base.nbytes); sess.h2d(d, np.ascontiguousarray(base))
    rr = region.RegionRunner(sess, R, C, patch, stride, batch=5)
    preds = np.empty((R, C, S, S), np.float32)
    dp = sess.malloc(preds.nbytes)
    assert rr.predict_rows(d, 12, 64, 64, 13, True, 0, dp) == R * C
    sess.d2h(preds, dp); sess.sync()
    out, (y0, y1) = rr.blend(dp, 0, R)
    assert (y0, y1) == (0, RR.canvas_size(R, patch, stride)) and np.array_equal(out, RR.blend_region(preds, stride))
    model = PredictRef(predict_weights)
    for (r_, c_) in ((0, 0), (3, 2)):
        m = RR.synth_canvas_patch(base, r_ * stride, c_ * stride, patch)[None]
        ref = model.forward(P.normalize_subtile(P.assemble(m), MIN_ALL, MAX_ALL))[0]
        assert np.abs(preds[r_, c_] - ref).max() < 1e-3
    sess.free(d); sess.free(dp)
