"""Generate tests/golden/recreate.npz: outputs of the REFERENCE's `recreate_resegmented_tifs` / `mosaic_subtiles`
(/root/reference/src/resegment_tiles_wide.py:1169-1547, through oracle/refshim.py, this container only) on seeded folders of
subtile predictions: normal subtiles, left / right / up / down border strips, no-data subtiles and no-data pixels.
`skimage.transform.resize` is the shim's restatement of scikit-image >= 0.19 (anti-aliased when shrinking).
Usage: python tools/make_golden_recreate.py"""
import os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

# (name, SIZE, shape = (Y, X), normal subtile size, normal x offsets, normal y offsets, strips, seed)
# strips: (kind, x, y, file shape)   kind l -> <x>/left<y>.npy, r -> right<x>/<y>.npy, u -> <x>/up<y>.npy, d -> <x>/down<y>.npy
CASES = [
    ("small", 400, (310, 324), 168, [0, 156], [0, 142],
     [("l", 0, 0, (206, 200)), ("l", 0, 104, (206, 200)), ("r", 224, 0, (206, 200)), ("r", 224, 104, (206, 200)),
      ("u", 0, 0, (120, 324)), ("d", 0, 250, (120, 324))], 11),
    ("no_strips", 400, (310, 324), 168, [0, 156], [0, 142], [], 12),
    ("tile618", 670, (618, 618), 216, [0, 134, 268, 402], [0, 134, 268, 402],
     [("l", 0, 0, (206, 670)), ("l", 0, 138, (206, 670)), ("l", 0, 276, (206, 670)), ("l", 0, 412, (206, 670)),
      ("r", 283, 0, (206, 670)), ("r", 283, 138, (206, 670)), ("r", 283, 276, (206, 670)), ("r", 283, 412, (206, 670))], 13),
]

SAMPLE = ((5, 7), (2, 3))          # strides of the stored samples: 618-px case, small cases


def write_case(folder, case):
    """Seeded prediction files in the layout the reference's process_subtiles loops write (probabilities 0..1, 255 = no data)."""
    name, size, shape, sub, xs, ys, strips, seed = case
    r = np.random.default_rng(seed)
    k = 0
    for x in xs:
        os.makedirs(os.path.join(folder, str(x)), exist_ok=True)
        for y in ys:
            p = np.clip(0.4 + 0.3 * np.sin(np.arange(sub)[:, None] / 17.0 + x / 50.0) * np.cos(np.arange(sub)[None] / 23.0 + y / 40.0)
                        + 0.05 * r.standard_normal((sub, sub)), 0, 1).astype(np.float32)
            if k == 1:
                p[:] = 255.                                     # a subtile without any data
            if k == 2:
                p[10:40, 20:70] = 255.                          # no-data pixels inside a subtile
            np.save(os.path.join(folder, str(x), "%d.npy" % y), p)
            k += 1
    for j, (kind, x, y, fshape) in enumerate(strips):
        p = np.clip(0.5 + 0.25 * np.cos(np.arange(fshape[0])[:, None] / 19.0 + j) * np.sin(np.arange(fshape[1])[None] / 29.0)
                    + 0.05 * r.standard_normal(fshape), 0, 1).astype(np.float32)
        if j == 1:
            p[5:25, -30:] = 255.; p[5:25, :30] = 255.           # no-data pixels on both halves of a strip
        if j == 3 and len(strips) > 6:
            p[:] = 255.                                         # a strip without any data
        d = os.path.join(folder, ("right%d" % x) if kind == "r" else str(x))
        os.makedirs(d, exist_ok=True)
        np.save(os.path.join(d, {"l": "left%d", "r": "%d", "u": "up%d", "d": "down%d"}[kind] % y + ".npy"), p)


def main():
    from oracle import refshim
    m = refshim.ref("resegment_tiles_wide")
    out = {}
    for case in CASES:
        name, size, shape = case[0], case[1], case[2]
        folder = tempfile.mkdtemp() + "/"
        write_case(folder, case)
        m.SIZE = size
        preds, sums = m.recreate_resegmented_tifs(folder, shape)
        print(name, preds.shape, preds.dtype, "no-data px:", int((preds == 255).sum()), "mean:", float(preds[preds < 255].mean()), flush=True)
        st = SAMPLE[0] if preds.size > 200000 else SAMPLE[1]      # a strided sample + checksums of the whole arrays
        out[name + "_preds_sample"] = preds[::st[0], ::st[1]].copy(); out[name + "_sums_sample"] = sums[::st[0], ::st[1]].copy()
        out[name + "_check"] = np.array([preds.sum(), np.nansum(sums), float((preds == 255).sum())])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "recreate.npz"), **out)


if __name__ == "__main__":
    main()
