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


def _mk_cv():
    import importlib.util
    spec = importlib.util.spec_from_file_location("mk_cv", os.path.join(os.path.dirname(__file__), "..", "tools", "make_golden_cv.py"))
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)
    return m


def test_superresolve_oracle_pinned_to_opencv_execution_of_the_released_graph(sr_weights):
    """tests/golden/superresolve_cv.npz = the released superresolve_graph.pb run by OpenCV's DNN module (a third-party
    executor sharing no code with the oracle; tools/make_golden_cv.py).  The interpreter-made golden and the restatement
    must agree with it to float32 op-order noise, on the golden's input and on a 118-px padded window."""
    cv = golden("superresolve_cv.npz")
    g = golden("superresolve.npz")
    assert np.abs(g["y"] - cv["y_small"]).max() < 3e-5                      # interpreter vs OpenCV (measured 9.9e-6)
    ref = SuperresolveRef(sr_weights)
    assert np.abs(ref.forward(g["x"], g["x"][..., 4:]) - cv["y_small"]).max() < 3e-5
    x = _mk_cv().window_input(int(cv["seed_window"]))
    y = ref.forward(x, x[..., 4:])
    assert y.shape == cv["y_window"].shape and np.abs(y - cv["y_window"]).max() < 3e-5
    assert np.abs(cv["y_window"] - x[..., 4:]).max() > 1e-3                  # the network does change the bilinear input


def test_opencv_runs_the_released_superresolve_graph_live():
    """Where the reference tree and cv2 are present: re-run OpenCV on the released graph and compare with the stored file."""
    pb = "/root/reference/models-release/supres-40k-swir/superresolve_graph.pb"
    cv2 = pytest.importorskip("cv2")
    if not os.path.exists(pb):
        pytest.skip("reference not mounted")
    mk = _mk_cv()
    cv = golden("superresolve_cv.npz")
    y = mk.run_opencv(pb, mk.window_input(int(cv["seed_window"])))
    assert np.abs(y - cv["y_window"]).max() < 1e-5


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


@pytest.mark.skipif(not os.path.isdir("/root/reference/models-release"), reason="reference tree absent")
@pytest.mark.parametrize("size", [76, 124, 172, 220])
def test_interpreter_covers_every_node_of_every_released_graph(size):
    """All four predict_graph-*.pb (four different weight sets, SURVEY headline fact 2) at their native size: the
    interpreter reaches the fetch without an unimplemented op, every node the fetch statically depends on was executed
    except (a) the DropBlock training branch (dead: tf.cond on is_training=False, pb:is_training) and (b) the
    Enter/Merge/NextIteration plumbing of the two while frames, which _run_frame drives directly; and the hand
    restatement with the weights of THE SAME .pb agrees with it."""
    from oracle.tfgraph_interp import GraphInterpreter
    from sentinel_tree_cover_b200.weights import load_predict_pb
    pb = "/root/reference/models-release/master-ckpt-frozen/predict_graph-%d.pb" % size
    gi = GraphInterpreter(pb)
    x = P.synth_model_input(1, size, 10 + size)
    y = gi.run("conv2d/Sigmoid", {"Placeholder": x})[..., 0]
    assert y.shape == (1, size - 14, size - 14)
    anc = gi.ancestors("conv2d/Sigmoid")
    assert len(gi.nodes) == 1466 and {"Placeholder", "PlaceholderWithDefault", "is_training"} <= anc
    unvisited = anc - set(gi.visited)
    rest = [n for n in unvisited if "drop_block2d" not in n and gi.nodes[n]["op"] not in ("Enter", "Merge", "NextIteration")]
    assert rest == [], rest[:10]
    assert not any(op == "RandomUniform" for op in gi.visited.values())        # the stochastic branch never ran
    ops = set(gi.visited.values())
    assert {"Conv2D", "MirrorPad", "ReverseSequence", "Select", "ResizeNearestNeighbor", "MaxPool", "Tanh", "Sigmoid", "Exit"} <= ops
    assert sum(1 for op in gi.visited.values() if op == "Conv2D") == 28        # SURVEY 8c op inventory
    y2 = PredictRef(load_predict_pb(pb)).forward(x)
    assert np.abs(y - y2).max() < 2e-5


@pytest.mark.skipif(not os.path.isdir("/root/reference/models-release"), reason="reference tree absent")
def test_interpreter_covers_every_node_of_the_superresolve_graph():
    from oracle.tfgraph_interp import GraphInterpreter
    pb = "/root/reference/models-release/supres-40k-swir/superresolve_graph.pb"
    gi = GraphInterpreter(pb)
    r = np.random.default_rng(3)
    x = r.uniform(0, 0.5, (2, 40, 36, 10)).astype(np.float32)
    gi.run("Add_2", {"Placeholder": x, "Placeholder_1": x[..., 4:]})
    anc = gi.ancestors("Add_2")
    assert anc <= set(gi.visited), sorted(anc - set(gi.visited))[:10]
    assert sum(1 for op in gi.visited.values() if op == "Conv2D") == 6


TF_GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("size", [76, 124, 172, 220])
def test_restatement_matches_real_tensorflow_golden_when_present(size, predict_weights):
    """tools/make_golden_tf.py writes tests/golden/model_tf_<size>.npz from a real tf.compat.v1 Session.run on a machine
    that has TensorFlow (none here: the model oracle is otherwise pinned to the GraphDef interpreter only).  When such a
    file has been committed, the restatement must reproduce TensorFlow's own output."""
    path = os.path.join(TF_GOLD, "model_tf_%d.npz" % size)
    if not os.path.exists(path):
        pytest.skip("no TensorFlow golden committed (run tools/make_golden_tf.py where TF >= 2.13 is installed)")
    g = np.load(path)
    from sentinel_tree_cover_b200.weights import load_npz
    w = {k[2:]: g[k] for k in g.files if k.startswith("w/")} or predict_weights
    x = P.synth_model_input(int(g["batch"]), size, int(g["seed"]))
    y = PredictRef(w).forward(x, g["length"])
    assert np.abs(y - g["y"]).max() < 2e-5
