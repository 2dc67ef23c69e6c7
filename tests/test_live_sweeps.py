"""Where /root/reference is mounted (this container, not the GPU box): the cloud-stage oracles against the reference's OWN
functions run live through oracle/refshim.py on seeds and shapes the committed goldens do not hold -- the goldens pin 5 + 3
mask cases and 2 + 3 fill cases; these sweeps look for an input on which a restatement drifts.  Bit-identical or fail."""
import contextlib
import io
import os
import random
import numpy as np
import pytest
from oracle import refshim, cloud_ref, cloudfill_ref

pytestmark = pytest.mark.skipif(not refshim.available(), reason="reference not mounted")

MASK_CASES = [(7, 50, 66, 101), (10, 90, 70, 102), (4, 36, 36, 103), (14, 64, 64, 104), (6, 120, 100, 105), (11, 72, 88, 106),
              (8, 33, 47, 107), (16, 56, 60, 108)]
FILL_CASES = [(7, 70, 66, 201, 5), (10, 90, 80, 202, 6), (4, 48, 52, 203, 7), (12, 64, 64, 204, 8), (6, 100, 100, 205, 9),
              (9, 72, 56, 206, 10)]


@pytest.mark.parametrize("case", MASK_CASES, ids=lambda c: "T%dx%dx%d" % c[:3])
def test_cloud_mask_oracle_equals_live_reference(case):
    """identify_clouds_shadows, src/preprocessing/cloud_removal.py:1215-1677."""
    T, H, W, seed = case
    cr = refshim.ref("preprocessing.cloud_removal")
    img, dem = cloud_ref.synth_cloudy_cube(T, H, W, seed)
    with contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
        clouds, fcps = cr.identify_clouds_shadows(np.copy(img), np.copy(dem), None)
        oc, of = cloud_ref.identify_clouds_shadows(np.copy(img), np.copy(dem))[:2]
    assert np.array_equal(np.asarray(clouds) > 0, np.asarray(oc) > 0) and np.array_equal(np.asarray(fcps) > 0, np.asarray(of) > 0)
    assert 0 < np.mean(clouds) < 1


@pytest.mark.parametrize("case", FILL_CASES, ids=lambda c: "T%dx%dx%d" % c[:3])
def test_cloud_fill_oracle_equals_live_reference(case, tmp_path, monkeypatch):
    """remove_cloud_and_shadows, :888-973 (+ make_aligned_mosaic, align_interp_array_randomforest): filled values, feather
    weights, removal list and the position of Python's generator after the call."""
    T, H, W, seed, rseed = case
    cr = refshim.ref("preprocessing.cloud_removal")
    monkeypatch.chdir(tmp_path)                      # the reference dumps debug .npy files into the CWD
    img, dem = cloud_ref.synth_cloudy_cube(T, H, W, seed)
    with np.errstate(all="ignore"):
        clouds, fcps = cloud_ref.identify_clouds_shadows(img, dem)[:2]
    random.seed(rseed)
    with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()), np.errstate(all="ignore"):
        tiles, areas, to_remove = cr.remove_cloud_and_shadows(np.copy(img), np.copy(clouds), np.copy(clouds), list(range(T)),
                                                              np.copy(fcps), np.zeros((H, W, 2), np.float32))
    nxt = random.random()
    random.seed(rseed)
    with np.errstate(all="ignore"):
        got = cloudfill_ref.remove_cloud_and_shadows(np.copy(img), np.copy(clouds), np.copy(fcps))
    assert np.array_equal(tiles, got[0]) and np.array_equal(areas, got[1]) and list(to_remove) == list(got[2])
    assert random.random() == nxt
    assert (tiles != img).any()
