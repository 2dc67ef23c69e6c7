"""Model weights: frozen GraphDef (.pb) -> canonical {name: float32 ndarray}.

The reference binds its models at src/download_and_predict_job.py:1785-1826
(tf.import_graph_def of predict_graph-<SIZE+14>.pb and superresolve_graph.pb).
Here the Const tensors are pulled straight out of the protobuf (pbread.py) and
renamed to a short canonical scheme used by both the CUDA context and the
test checker.  A canonical dict can be saved to / loaded from .npz so tests and the
benchmark run on machines that do not have the reference tree.

Canonical names, predict graph:
  gru.{fw,bw}.gates_w [3,3,49,64]   gru.{d}.{r,u}_gamma/_beta [32]
  gru.{d}.cand_w [3,3,49,32]        gru.{d}.cand_sse_w [32]   gru.{d}.y_gamma/_beta [32]
  {blk}.w [3,3,Cin,Cout] {blk}.gamma {blk}.beta [Cout] {blk}.sse_w [Cout] {blk}.sse_b [1]
     blk in conv_median conv_concat conv1 conv2 up2 up2_out up3 out
  head.w [64] head.b [1]
Super-resolve graph: sr.{in,r01,r02,r11,r12,out}.w / .b
"""
import numpy as np
from .pbread import read_consts

BLOCKS = ["conv_median", "conv_concat", "conv1", "conv2", "up2", "up2_out", "up3", "out"]
BLOCK_SHAPES = {  # (Cin, Cout, SAME?)
    "conv_median": (17, 64, True), "conv_concat": (128, 64, True), "conv1": (64, 128, False),
    "conv2": (128, 256, False), "up2": (256, 128, True), "up2_out": (256, 128, True),
    "up3": (128, 64, True), "out": (128, 64, False)}
SR_LAYERS = [("in", "in_conv/conv2d", 10, 32), ("r01", "01_conv/conv2d_1", 32, 32),
             ("r02", "02_conv/conv2d_2", 32, 32), ("r11", "11_conv/conv2d_3", 32, 32),
             ("r12", "12_conv/conv2d_4", 32, 32), ("out", "out_conv/conv2d_5", 32, 6)]


def _find(consts, suffix, contains=None):
    hits = [k for k in consts if k.endswith(suffix) and (contains is None or contains in k)]
    if len(hits) != 1:
        raise KeyError("expected exactly one Const matching %r/%r, got %r" % (contains, suffix, hits))
    return consts[hits[0]]


def load_predict_pb(path):
    c = read_consts(path)
    w = {}
    for d in ("fw", "bw"):
        pre = "down_16/bidirectional_rnn/%s/" % d
        w["gru.%s.gates_w" % d] = c[pre + "conv_gru_cell/gates/kernel"]
        w["gru.%s.cand_w" % d] = c[pre + "conv_gru_cell/candidate/kernel"]
        w["gru.%s.cand_sse_w" % d] = c[pre + "conv_gru_cell/candidate/kernel_1"].reshape(32)
        for g in ("r", "u"):
            w["gru.%s.%s_gamma" % (d, g)] = _find(c, "gamma_gates_" + g, pre)
            w["gru.%s.%s_beta" % (d, g)] = _find(c, "beta_gates_" + g, pre)
        w["gru.%s.y_gamma" % d] = _find(c, "gamma_candidate_y", pre)
        w["gru.%s.y_beta" % d] = _find(c, "beta_candidate_y", pre)
    for b in BLOCKS:
        ks = [k for k in c if k.startswith(b + "_conv/") and k.endswith("/kernel") and "/mask/" not in k]
        assert len(ks) == 1, ks
        w[b + ".w"] = c[ks[0]]
        w[b + ".gamma"] = c["%s_norm/gamma_%s" % (b, b)]
        w[b + ".beta"] = c["%s_norm/beta_%s" % (b, b)]
        w[b + ".sse_w"] = c["csse_%s_conv/kernel" % b].reshape(-1)
        w[b + ".sse_b"] = c["csse_%s_conv/bias" % b].reshape(1)
    w["head.w"] = c["conv2d/kernel"].reshape(64)
    w["head.b"] = c["conv2d/bias"].reshape(1)
    # structural assumptions the kernels rely on (Swish beta = 1, zoneout 0.75/0.25)
    for k, v in c.items():
        if k == "beta" or (k.startswith("beta_") and k[5:].isdigit()):
            assert float(v) == 1.0, (k, v)
    for d in ("fw", "bw"):
        p = "down_16/bidirectional_rnn/%s/%s/while/" % (d, d)
        assert float(c[p + "mul/y"]) == 0.75 and float(c[p + "mul_1/y"]) == 0.25
        assert float(c[p + "mul_2/y"]) == 0.75 and float(c[p + "mul_3/y"]) == 0.25
    return {k: np.ascontiguousarray(v, np.float32) for k, v in w.items()}


def load_superresolve_pb(path):
    c = read_consts(path)
    w = {}
    for short, scope, cin, cout in SR_LAYERS:
        w["sr.%s.w" % short] = c[scope + "/kernel"]
        w["sr.%s.b" % short] = c[scope + "/bias"]
        assert w["sr.%s.w" % short].shape == (3, 3, cin, cout)
    assert abs(float(c["Const"]) - 0.1) < 1e-7 and abs(float(c["Const_1"]) - 0.1) < 1e-7
    return {k: np.ascontiguousarray(v, np.float32) for k, v in w.items()}


def random_predict_weights(seed=0):
    """Random-init weights of the released architecture (He-normal kernels,
    gamma=1, beta=0) -- used by bench.py where the .pb files do not exist."""
    r = np.random.default_rng(seed)
    w = {}
    def he(shape):
        fan_in = shape[0] * shape[1] * shape[2]
        return (r.standard_normal(shape) * np.sqrt(2.0 / fan_in)).astype(np.float32)
    for d in ("fw", "bw"):
        w["gru.%s.gates_w" % d] = he((3, 3, 49, 64)) * 0.7
        w["gru.%s.cand_w" % d] = he((3, 3, 49, 32)) * 0.7
        w["gru.%s.cand_sse_w" % d] = (r.standard_normal(32) * 0.2).astype(np.float32)
        for g in ("r", "u", "y"):
            w["gru.%s.%s_gamma" % (d, g)] = (1 + 0.1 * r.standard_normal(32)).astype(np.float32)
            w["gru.%s.%s_beta" % (d, g)] = (0.1 * r.standard_normal(32)).astype(np.float32)
    for b in BLOCKS:
        cin, cout, _ = BLOCK_SHAPES[b]
        w[b + ".w"] = he((3, 3, cin, cout))
        w[b + ".gamma"] = (1 + 0.1 * r.standard_normal(cout)).astype(np.float32)
        w[b + ".beta"] = (0.1 * r.standard_normal(cout)).astype(np.float32)
        w[b + ".sse_w"] = (r.standard_normal(cout) * 0.1).astype(np.float32)
        w[b + ".sse_b"] = (r.standard_normal(1) * 0.1).astype(np.float32)
    w["head.w"] = (r.standard_normal(64) * 0.2).astype(np.float32)
    w["head.b"] = np.zeros(1, np.float32)
    return w


def random_superresolve_weights(seed=0):
    r = np.random.default_rng(seed)
    w = {}
    for short, _, cin, cout in SR_LAYERS:
        w["sr.%s.w" % short] = (r.standard_normal((3, 3, cin, cout)) * np.sqrt(2.0 / (9 * cin)) * 0.5).astype(np.float32)
        w["sr.%s.b" % short] = (r.standard_normal(cout) * 0.01).astype(np.float32)
    return w


def save_npz(path, w):
    np.savez(path, **w)


def load_npz(path):
    with np.load(path) as z:
        return {k: np.ascontiguousarray(z[k], np.float32) for k in z.files}
