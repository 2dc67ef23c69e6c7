"""CPU: the oracle's model restatement against the golden vectors produced by the
mechanical GraphDef interpreter (tools/make_golden.py), and -- where the reference tree
is present -- against the interpreter itself on another released graph."""
import os
import numpy as np
import pytest
from conftest import golden
from oracle.model_ref import PredictRef, SuperresolveRef
from oracle import preproc_ref as P


def test_predict_restatement_matches_graph_golden(predict_weights):
    g = golden("model_172.npz")
    m = PredictRef(predict_weights)
    for k in ("a", "b"):
        x = P.synth_model_input(1, 172, int(g["seed_" + k]))
        y = m.forward(x, np.full(1, int(g["length_" + k])))[0]
        assert y.shape == (158, 158)
        assert np.abs(y - g["y_" + k]).max() < 2e-5      # float32 op-order noise only


def test_superresolve_restatement_matches_graph_golden(sr_weights):
    g = golden("superresolve.npz")
    y = SuperresolveRef(sr_weights).forward(g["x"], g["x"][..., 4:])
    assert np.abs(y - g["y"]).max() < 5e-6


def test_model_structural_invariants(predict_weights):
    m = PredictRef(predict_weights)
    x = P.synth_model_input(2, 44, 1)
    y = m.forward(x)
    assert y.shape == (2, 30, 30) and (y > 0).all() and (y < 1).all()
    # batch independence
    y0 = m.forward(x[:1])
    assert np.abs(y0[0] - y[0]).max() < 1e-6
    # length gating: frames beyond `length` are never read (pb:.../while/Select_1)
    x2 = x.copy(); x2[:, 2:4] = 0.123
    assert np.abs(m.forward(x, np.full(2, 2)) - m.forward(x2, np.full(2, 2))).max() < 1e-7


@pytest.mark.skipif(not os.path.isdir("/root/reference/models-release"), reason="reference tree absent")
def test_restatement_vs_interpreter_on_76_graph():
    from oracle.tfgraph_interp import GraphInterpreter
    from sentinel_tree_cover_b200.weights import load_predict_pb
    pb = "/root/reference/models-release/master-ckpt-frozen/predict_graph-76.pb"
    x = P.synth_model_input(3, 76, 2)
    L = np.array([4, 3, 1], np.int64)
    y = GraphInterpreter(pb).run("conv2d/Sigmoid", {"Placeholder": x, "PlaceholderWithDefault": L})[..., 0]
    y2 = PredictRef(load_predict_pb(pb)).forward(x, L)
    assert np.abs(y - y2).max() < 2e-5
