"""Generate tests/golden/process_subtiles.npz by running the REFERENCE process_subtiles
(src/download_and_predict_job.py:1125-1486) through oracle/refshim.py on a seeded synthetic ARD cube.
TensorFlow is absent, so `predict_subtile` -- the one call into the TF session -- is replaced by the
oracle restatement of the frozen graph (oracle.model_ref.PredictRef with the released weights); every
other line of the reference function runs unmodified.  File / S3 outputs are stubbed out.
Usage: python tools/make_golden_subtiles.py"""
import os, sys, tempfile, types
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refshim, subtiles_ref

CASE = dict(seed=71, n=9, H=340, W=330)
# a date with > 10 % NaN pixels (dropped: interpolate_na_vals turns them into zeros, id_missing_px then removes the date,
# src/download_and_predict_job.py:1148, 1031-1036) and a date with a small NaN block (median-filled)
CASE_NAN = dict(seed=72, n=8, H=250, W=244)


def add_nans(s2):
    s2 = np.copy(s2)
    H, W = s2.shape[1:3]
    s2[3, : H // 2, : W // 2, :] = np.nan            # 25 % of date 3
    s2[5, 20:30, 40:60, 2:5] = np.nan                # 0.3 % of date 5, three bands
    return s2


def main():
    job = refshim.ref("download_and_predict_job")
    tmp = tempfile.mkdtemp() + "/"
    os.chdir(tmp)
    from sentinel_tree_cover_b200.api import MIN_ALL, MAX_ALL
    subtiles_ref.patch_reference(job, tmp, MIN_ALL, MAX_ALL)
    s2, dates, interp, s1, dem = subtiles_ref.synth_ard(**CASE)
    os.makedirs(f"{tmp}3/4/", exist_ok=True)
    job.process_subtiles(3, 4, np.copy(s2), np.copy(dates), np.copy(interp), np.copy(s1), np.copy(dem), None, [0, 0, 1, 1], 158, None)
    out = {"case": np.array([CASE["seed"], CASE["n"], CASE["H"], CASE["W"]], np.int32)}
    path = f"{tmp}3/4/processed/"
    names = []
    for fy in sorted(os.listdir(path), key=int):
        for f in sorted(os.listdir(path + fy), key=lambda s: int(s[:-4])):
            names.append((int(fy), int(f[:-4])))
            out["pred_%s_%s" % (fy, f[:-4])] = np.load(path + fy + "/" + f)
    out["names"] = np.array(names, np.int32)
    print(len(names), "subtiles", {k: (v.shape, v.dtype) for k, v in list(out.items())[2:4]})
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "process_subtiles.npz"), **out)
    # ---- NaN case ----
    tmp2 = tempfile.mkdtemp() + "/"
    os.chdir(tmp2)
    subtiles_ref.patch_reference(job, tmp2, MIN_ALL, MAX_ALL)
    s2, dates, interp, s1, dem = subtiles_ref.synth_ard(**CASE_NAN)
    s2 = add_nans(s2)
    os.makedirs(f"{tmp2}3/4/", exist_ok=True)
    job.process_subtiles(3, 4, np.copy(s2), np.copy(dates), np.copy(interp), np.copy(s1), np.copy(dem), None, [0, 0, 1, 1], 158, None)
    out = {"case": np.array([CASE_NAN["seed"], CASE_NAN["n"], CASE_NAN["H"], CASE_NAN["W"]], np.int32)}
    path = f"{tmp2}3/4/processed/"
    names = []
    for fy in sorted(os.listdir(path), key=int):
        for f in sorted(os.listdir(path + fy), key=lambda s: int(s[:-4])):
            names.append((int(fy), int(f[:-4])))
            out["pred_%s_%s" % (fy, f[:-4])] = np.load(path + fy + "/" + f).astype(np.float16 if False else np.float32)
    out["names"] = np.array(names, np.int32)
    print("NaN case:", len(names), "subtiles")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "process_subtiles_nan.npz"), **out)


if __name__ == "__main__":
    main()
