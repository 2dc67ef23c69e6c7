"""One whole tile through the mirrored per-tile drivers, raw arrays -> uint8 tree-cover tile:
process_tile (decode, dB, upsample, cloud masks / removal) -> superresolve_large_tile -> process_subtiles
(smoothing, 36 subtiles, one batched forward, post-filters, .npy files) -> load_mosaic_predictions.
Synthetic raw tile (oracle.tile_ref, 618 x 618 px, n dates); prints per-stage wall times as one JSON line.
Usage (GPU box): python tools/bench_tile.py [--n 12] [--reps 2]"""
import argparse, json, os, random, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=12)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--trace", default=None, help="CSV path: one more chain tile with CUDA events around every kernel launch")
    args = ap.parse_args()
    from oracle import tile_ref                       # synthetic raw tile only
    from sentinel_tree_cover_b200 import api, tile
    gold = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
    sess = api.StcSession(0, predict_weights=os.path.join(gold, "weights_predict_172.npz"),
                          superresolve_weights=os.path.join(gold, "weights_superresolve.npz"))
    store = tile_ref.FakeStore(tile_ref.synth_raw_tile(91, n=args.n, h=309, w=309))
    res = []
    for rep in range(args.reps):
        root = tempfile.mkdtemp() + "/"
        random.seed(4)
        t = [time.perf_counter()]
        s2, dates, interp, s1, dem, cloudshad, snow = tile.process_tile(1, 2, None, "/nonexistent/", [0, 0, 1, 1], make_shadow=True,
                                                                        sess=sess, loader=store.load, exists=store.exists)
        t.append(time.perf_counter())
        s2 = api.superresolve_large_tile(np.ascontiguousarray(s2), sess)
        t.append(time.perf_counter())
        tile.process_subtiles(1, 2, s2, dates, interp, s1, dem, sess, [0, 0, 1, 1], 158, None, local_path=root, length=4)
        t.append(time.perf_counter())
        out = api.load_mosaic_predictions(root + "1/2/processed/", 1, sess)
        t.append(time.perf_counter())
        d = np.diff(t) * 1e3
        res.append({"process_tile_ms": round(d[0], 1), "superresolve_ms": round(d[1], 1), "process_subtiles_ms": round(d[2], 1),
                    "mosaic_ms": round(d[3], 1), "total_ms": round(float(d.sum()), 1), "dates_kept": int(len(dates)),
                    "out_shape": list(out.shape), "out_dtype": str(out.dtype), "tree_cover_mean": round(float(out[out <= 100].mean()), 2)})
    # the same tile through the device-resident chain (ONE C call: stc_tile_run_host)
    raw = store.raw
    pin = {k: sess.pinned_empty(raw[k].shape, raw[k].dtype) for k in ("s2_10", "s2_20", "s1", "dem")}
    for k in pin:
        pin[k][...] = raw[k]
    chain = []
    for rep in range(args.reps + 1):
        random.seed(4)
        t0 = time.perf_counter()
        out2, kept = sess.run_tile(pin["s2_10"], pin["s2_20"], pin["s1"], pin["dem"], raw["s2_dates"])
        chain.append(round((time.perf_counter() - t0) * 1e3, 1))
    if not res:
        out = out2                         # --reps 0: the chain alone (profiling runs)
    if args.trace:
        sess.trace(1)
        random.seed(4)
        sess.run_tile(pin["s2_10"], pin["s2_20"], pin["s1"], pin["dem"], raw["s2_dates"])
        sess.trace(0, args.trace)
        tot = {}
        import csv
        for r in csv.DictReader(open(args.trace)):
            e = tot.setdefault(r["label"], [0, 0.0]); e[0] += 1; e[1] += float(r["end_ms"]) - float(r["start_ms"])
        print("kernel table (launches, ms, us/launch), sum %.2f ms" % sum(v[1] for v in tot.values()), file=sys.stderr)
        for name, (k, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:40]:
            print("  %-28s %4d %8.3f %8.1f" % (name, k, ms, ms / k * 1e3), file=sys.stderr)
    print(json.dumps({"tile": "618x618, %d dates, synthetic raw (uint16 S2/S1, f32 DEM)" % args.n, "runs": res,
                      "chain_ms": chain, "chain_equals_mirrors": bool(np.array_equal(out2, out)),
                      "chain_vs_mirrors_differing_px": int((out2 != out).sum()), "chain_vs_mirrors_max_abs": int(np.abs(out2.astype(int) - out.astype(int)).max()),
                      "chain_dates_kept": int(len(kept)),
                      "pool": sess.pool_info()}), flush=True)


if __name__ == "__main__":
    main()
