"""GPU: the DSen2 super-resolution against outputs of the RELEASED superresolve_graph.pb executed by OpenCV's DNN module
(tests/golden/superresolve_cv.npz, tools/make_golden_cv.py) -- a third-party executor of the reference's own graph.
Written after the GPU budget of round 2 had ended (its CPU dry run with the fp16-operand oracle standing in for the GPU gives
9.0e-4 / 2.2e-4 against the 2e-3 tolerance); the file sorts last so that its first GPU run cannot cut the suite short."""
import os
import numpy as np
import pytest
from conftest import golden

pytestmark = pytest.mark.gpu


def test_superresolve_vs_opencv_execution_of_the_released_graph(sess):
    """Against outputs of the released superresolve_graph.pb run by OpenCV's DNN module (tests/golden/superresolve_cv.npz,
    tools/make_golden_cv.py): a third-party executor of the reference's own graph.  Same tolerance as the graph golden."""
    import importlib.util
    cv = golden("superresolve_cv.npz")
    g = golden("superresolve.npz")
    y = sess.superresolve(g["x"], g["x"][..., 4:])
    assert np.abs(y - cv["y_small"]).max() < 2e-3
    spec = importlib.util.spec_from_file_location("mk_cv", os.path.join(os.path.dirname(__file__), "..", "tools", "make_golden_cv.py"))
    mk = importlib.util.module_from_spec(spec); spec.loader.exec_module(mk)
    x = mk.window_input(int(cv["seed_window"]))
    y = sess.superresolve(x, x[..., 4:])
    err = np.abs(y - cv["y_window"]).max()
    print("superresolve vs OpenCV, 118-px window", err)
    assert err < 2e-3
