"""TEST INFRASTRUCTURE -- the CPU oracle ports chained like the body of the reference's main loop
(/root/reference/src/download_and_predict_job.py:1995-2020): cloud / shadow masks -> cloud removal -> super-resolve ->
smoothing + composites -> per-subtile forward.  Used ONLY as bench.py's `cpu_baseline` leg of the `tile_chain` object (a
bounded sample of a tile timed on the host cores) and by tests; nothing in the product path imports it.
Stages follow: cloud_ref.identify_clouds_shadows (cloud_removal.py:1215-1677), cloudfill_ref.remove_cloud_and_shadows
(:888-973), model_ref.SuperresolveRef (superresolve_large_tile :95-147), preproc_ref.smooth_stack (smooth_large_tile
:1057-1096), assemble + normalize_subtile + PredictRef.forward per 172-px window (process_subtiles :1345-1483)."""
import random
import time
import numpy as np


def run_sample(n_dates, px, seed, predict_weights, superresolve_weights, min_all, max_all):
    """One `px` x `px` cut-out with `n_dates` dates through every stage; returns ({stage: seconds}, n_windows)."""
    import torch
    from oracle import cloud_ref, cloudfill_ref, preproc_ref as P
    from oracle.model_ref import PredictRef, SuperresolveRef
    from sentinel_tree_cover_b200 import regrid
    from sentinel_tree_cover_b200.synth import synth_cloudy_cube
    img, dem = synth_cloudy_cube(n_dates, px, px, seed)
    dates = (np.arange(n_dates) * (330 // n_dates) + 10).astype(np.int64)
    t = {}
    t0 = time.perf_counter()
    clouds, fcps = cloud_ref.identify_clouds_shadows(img, dem)
    t["cloud_masks"] = time.perf_counter() - t0; t0 = time.perf_counter()
    random.seed(seed)
    tiles, areas, _ = cloudfill_ref.remove_cloud_and_shadows(np.copy(img), np.copy(clouds), np.copy(fcps))
    t["cloud_removal"] = time.perf_counter() - t0; t0 = time.perf_counter()
    sr = SuperresolveRef(superresolve_weights)
    with torch.no_grad():
        for d in range(n_dates):                         # one date per call keeps the working set in cache
            x = np.pad(tiles[d:d + 1], ((0, 0), (4, 4), (4, 4), (0, 0)), "reflect")
            tiles[d, ..., 4:] = np.asarray(sr.forward(x, x[..., 4:]))[0, 4:-4, 4:-4]
    t["superresolve"] = time.perf_counter() - t0; t0 = time.perf_counter()
    G, _ = regrid.regrid_matrix(dates)
    monthly = P.smooth_stack(tiles, G)                   # [12, px, px, 14]
    med = np.median(np.concatenate([tiles, P.make_indices(tiles)], -1), axis=0)
    quarterly = np.stack([np.median(monthly[3 * q:3 * q + 3], axis=0) for q in range(4)])
    t["smooth_composites"] = time.perf_counter() - t0; t0 = time.perf_counter()
    model = PredictRef(predict_weights)
    frames = np.concatenate([quarterly, med[None]], 0)    # [5, px, px, 14]
    o = px - 172                                          # four 172-px windows (a 618-px tile has 36: same 1/9 as the pixels at px = 206)
    nwin = 0
    for (x0, y0) in ((0, 0), (0, o), (o, 0), (o, o)):
        sub = np.zeros((5, 172, 172, 17), np.float32)
        cut = frames[:, x0:x0 + 172, y0:y0 + 172]
        sub[..., :10] = cut[..., :10]; sub[..., 10] = dem[x0:x0 + 172, y0:y0 + 172] / 90.0; sub[..., 13:] = cut[..., 10:14]
        model.forward(P.normalize_subtile(sub, min_all, max_all)[None])
        nwin += 1
    t["forward"] = time.perf_counter() - t0
    return t, nwin
