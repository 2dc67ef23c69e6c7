"""TEST INFRASTRUCTURE (oracle) -- mechanical CPU interpreter for the reference's
frozen TensorFlow GraphDefs.  Never imported by the product path.

The reference runs its models with `sess.run(op, feed_dict=...)`
(src/download_and_predict_job.py:115-117 for superresolve_graph.pb,
:355-357 for predict_graph-*.pb).  TensorFlow is not installable in this
environment, so the model arithmetic is anchored on the GraphDef itself: this
module executes the node list op-by-op (TF1 dataflow semantics, including
Switch/Merge conditionals and Enter/Merge/NextIteration/Exit while-loop frames,
TensorArray scatter/read) so that no hand transcription of the architecture is
involved.  oracle/model_ref.py (the hand restatement that travels to the GPU
box) is validated against this interpreter.

Parity status: "parity unpinned" against real TensorFlow outputs (none exist in
the reference tree).  Op semantics follow the TF
op documentation: Conv2D = NHWC cross-correlation with HWIO kernels, MirrorPad
REFLECT excludes the border sample, tf.nn.moments is the biased variance, etc.
"""
import numpy as np
import torch
import torch.nn.functional as F

import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from sentinel_tree_cover_b200.pbread import read_graph  # noqa: E402


class _Dead:
    def __repr__(self):
        return "DEAD"


DEAD = _Dead()


class TensorArray:
    def __init__(self):
        self.items = None


def _split_name(ref):
    if ref.startswith("^"):
        return None, 0
    if ":" in ref:
        n, i = ref.rsplit(":", 1)
        return n, int(i)
    return ref, 0


def _strided_slice(x, begin, end, strides, a):
    bm, em = a.get("begin_mask", 0), a.get("end_mask", 0)
    elm, nam, sam = a.get("ellipsis_mask", 0), a.get("new_axis_mask", 0), a.get("shrink_axis_mask", 0)
    idx = []
    for i in range(len(begin)):
        if elm >> i & 1:
            idx.append(Ellipsis)
        elif nam >> i & 1:
            idx.append(None)
        elif sam >> i & 1:
            idx.append(int(begin[i]))
        else:
            b = None if bm >> i & 1 else int(begin[i])
            e = None if em >> i & 1 else int(end[i])
            idx.append(slice(b, e, int(strides[i])))
    return x[tuple(idx)]


class GraphInterpreter:
    def __init__(self, pb_path, dtype=np.float32, threads=None):
        self.nodes = {n["name"]: n for n in read_graph(pb_path)}
        self.dtype = dtype
        self.memo = {}
        self.feeds = {}
        self.visited = {}          # node name -> op type, every node _compute ran (all runs, all loop iterations)
        if threads:
            torch.set_num_threads(threads)
        # while-loop frames, identified by the name prefix up to ".../while/"
        self.frames = {}
        for n in self.nodes.values():
            if n["op"] == "Exit":
                prefix = n["name"][: n["name"].rfind("/while/") + len("/while/")]
                self.frames.setdefault(prefix, None)

    # ------------------------------------------------------------------
    def run(self, fetch, feeds):
        self.memo = {}
        self.frames = {k: None for k in self.frames}   # while-loop results are per run
        self.feeds = {k: np.asarray(v) for k, v in feeds.items()}
        return self._eval(fetch, None)

    def ancestors(self, fetch):
        """Every node the fetch depends on through data edges (static reachability, control edges excluded)."""
        seen, stack = set(), [_split_name(fetch)[0]]
        while stack:
            n = stack.pop()
            if n in seen:
                continue
            seen.add(n)
            for r in self.nodes[n]["inputs"]:
                m, _ = _split_name(r)
                if m is not None and m not in seen:
                    stack.append(m)
        return seen

    def _frame_of(self, name):
        for p in self.frames:
            if name.startswith(p):
                return p
        return None

    def _eval(self, ref, it):
        """Evaluate tensor `ref`; `it` = (frame_prefix, per-iteration memo) or None."""
        name, oi = _split_name(ref)
        fr = self._frame_of(name)
        if self.nodes[name]["op"] == "Const":
            fr = None
        memo = it[1] if (it is not None and fr == it[0]) else self.memo
        ctx = it if (it is not None and fr == it[0]) else None
        # (nodes merely *named* under a while/ scope but fed only from outside, e.g.
        # variable reads behind an Enter, are evaluated in the global memo)
        if name not in memo:
            memo[name] = self._compute(self.nodes[name], ctx)
        v = memo[name]
        if isinstance(v, tuple):
            return v[oi]
        assert oi == 0, (name, oi)
        return v

    # ------------------------------------------------------------------
    def _run_frame(self, prefix):
        merges = [n for n in self.nodes.values() if n["op"] == "Merge" and n["name"].startswith(prefix)
                  and self._frame_of(n["name"]) == prefix]
        cond_node = [n for n in self.nodes.values() if n["op"] == "LoopCond" and n["name"].startswith(prefix)][0]
        state = {}
        nexts = {}
        for m in merges:
            ent, nxt = None, None
            for r in m["inputs"]:
                nn, _ = _split_name(r)
                if self.nodes[nn]["op"] == "Enter":
                    ent = nn
                elif self.nodes[nn]["op"] == "NextIteration":
                    nxt = nn
            state[m["name"]] = self._eval(self.nodes[ent]["inputs"][0], None)
            nexts[m["name"]] = self.nodes[nxt]["inputs"][0]
        n_iter = 0
        while True:
            memo = dict(state)
            it = (prefix, memo)
            if not bool(self._eval(cond_node["name"], it)):
                break
            memo["__in_body__"] = True
            state = {m: self._eval(nexts[m], it) for m in state}
            n_iter += 1
            assert n_iter < 1000
        self.frames[prefix] = state
        return state

    # ------------------------------------------------------------------
    def _compute(self, n, it):
        op, name, a = n["op"], n["name"], n["attr"]
        ins = [r for r in n["inputs"] if not r.startswith("^")]
        self.visited[name] = op

        if op == "Const":
            v = a["value"]
            if v.dtype == np.float32:
                v = v.astype(self.dtype)
            return v
        if op == "Placeholder":
            return self.feeds[name].astype(self.dtype)
        if op == "PlaceholderWithDefault":
            if name in self.feeds:
                return self.feeds[name]
            return self._eval(ins[0], it)
        if op == "Exit":
            prefix = name[: name.rfind("/while/") + len("/while/")]
            state = self.frames[prefix] or self._run_frame(prefix)
            sw = self.nodes[_split_name(ins[0])[0]]
            return state[_split_name(sw["inputs"][0])[0]]
        if op == "Enter":
            return self._eval(ins[0], None)
        if op == "Merge":
            if it is not None or self._frame_of(name) is not None:
                raise RuntimeError("loop Merge %s must be pre-seeded" % name)
            for r in ins:
                v = self._eval(r, it)
                if v is not DEAD:
                    return v
            return DEAD
        if op == "Switch":
            data = self._eval(ins[0], it)
            pred = self._eval(ins[1], it)
            if data is DEAD or pred is DEAD:
                return (DEAD, DEAD)
            return (DEAD, data) if bool(pred) else (data, DEAD)

        # generic ops: lazy dead propagation
        x = []
        for r in ins:
            v = self._eval(r, it)
            if v is DEAD:
                return DEAD if op not in ("Split", "IdentityN") else (DEAD,) * 4
            x.append(v)

        if op in ("Identity", "StopGradient", "LoopCond", "NextIteration"):
            return x[0]
        if op == "IdentityN":
            return tuple(x)
        if op in ("Mul",):
            return x[0] * x[1]
        if op in ("AddV2", "Add"):
            return x[0] + x[1]
        if op == "Sub":
            return x[0] - x[1]
        if op == "RealDiv":
            return x[0] / x[1]
        if op == "Maximum":
            return np.maximum(x[0], x[1])
        if op == "Minimum":
            return np.minimum(x[0], x[1])
        if op == "SquaredDifference":
            return (x[0] - x[1]) ** 2
        if op == "Sqrt":
            return np.sqrt(x[0])
        if op == "Sigmoid":
            return (1.0 / (1.0 + np.exp(-x[0].astype(np.float64)))).astype(x[0].dtype)
        if op == "Tanh":
            return np.tanh(x[0])
        if op == "Relu":
            return np.maximum(x[0], 0)
        if op == "Sign":
            return np.sign(x[0])
        if op == "Reshape":
            return np.reshape(x[0], [int(s) for s in x[1]])
        if op == "Transpose":
            return np.transpose(x[0], [int(s) for s in x[1]])
        if op == "Mean":
            return np.mean(x[0], axis=tuple(int(s) for s in np.atleast_1d(x[1])), keepdims=bool(a.get("keep_dims", False)))
        if op == "Sum":
            return np.sum(x[0], axis=tuple(int(s) for s in np.atleast_1d(x[1])), keepdims=bool(a.get("keep_dims", False)))
        if op == "Max":
            return np.max(x[0], axis=tuple(int(s) for s in np.atleast_1d(x[1])), keepdims=bool(a.get("keep_dims", False)))
        if op == "All":
            return np.all(x[0], axis=tuple(int(s) for s in np.atleast_1d(x[1])))
        if op == "Cast":
            dst = {1: self.dtype, 3: np.int32, 9: np.int64, 10: np.bool_}[a["DstT"][1]]
            return np.asarray(x[0]).astype(dst)
        if op == "Shape":
            return np.asarray(np.shape(x[0]), np.int32)
        if op == "Size":
            return np.asarray(np.size(x[0]), np.int32)
        if op == "Pack":
            return np.stack([np.asarray(v) for v in x], axis=int(a.get("axis", 0)))
        if op == "ExpandDims":
            return np.expand_dims(x[0], int(x[1]))
        if op == "Fill":
            return np.full([int(s) for s in x[0]], x[1])
        if op == "Range":
            return np.arange(int(x[0]), int(x[1]), int(x[2]), dtype=np.int32)
        if op == "ConcatV2":
            return np.concatenate([np.atleast_1d(v) for v in x[:-1]], axis=int(x[-1]))
        if op == "Split":
            return tuple(np.split(x[1], int(a["num_split"]), axis=int(x[0])))
        if op == "StridedSlice":
            return _strided_slice(x[0], x[1], x[2], x[3], a)
        if op == "Equal":
            return np.equal(x[0], x[1])
        if op == "Less":
            return np.less(x[0], x[1])
        if op == "GreaterEqual":
            return np.greater_equal(x[0], x[1])
        if op == "LogicalNot":
            return np.logical_not(x[0])
        if op == "LogicalOr":
            return np.logical_or(x[0], x[1])
        if op == "LogicalAnd":
            return np.logical_and(x[0], x[1])
        if op == "Assert":
            assert bool(np.all(x[0])), "graph Assert failed at " + name
            return None
        if op == "Select":
            c = np.asarray(x[0])
            if c.ndim == 1 and np.ndim(x[1]) > 1:  # legacy Select: 1-D cond picks rows
                c = c.reshape((-1,) + (1,) * (np.ndim(x[1]) - 1))
            return np.where(c, x[1], x[2])
        if op == "BiasAdd":
            return x[0] + x[1]
        if op == "MirrorPad":
            mode = a["mode"].decode().lower()
            return np.pad(x[0], [(int(p[0]), int(p[1])) for p in x[1]], mode="reflect" if mode == "reflect" else "symmetric")
        if op == "Pad":
            return np.pad(x[0], [(int(p[0]), int(p[1])) for p in x[1]])
        if op == "Conv2D":
            assert a["data_format"] == b"NHWC" and list(a["strides"]) == [1, 1, 1, 1]
            pad = a["padding"].decode()
            k = x[1]
            t = torch.from_numpy(np.ascontiguousarray(np.transpose(x[0], (0, 3, 1, 2))))
            w = torch.from_numpy(np.ascontiguousarray(np.transpose(k, (3, 2, 0, 1))))
            if pad == "SAME":
                ph, pw = (k.shape[0] - 1) // 2, (k.shape[1] - 1) // 2
                assert k.shape[0] % 2 == 1 and k.shape[1] % 2 == 1
            else:
                ph = pw = 0
            y = F.conv2d(t, w, padding=(ph, pw))
            return np.ascontiguousarray(np.transpose(y.numpy(), (0, 2, 3, 1)))
        if op == "MaxPool":
            ks, st = list(a["ksize"]), list(a["strides"])
            assert a["padding"] == b"VALID" or (ks[1] == 1), name
            t = torch.from_numpy(np.ascontiguousarray(np.transpose(x[0], (0, 3, 1, 2))))
            y = F.max_pool2d(t, (ks[1], ks[2]), (st[1], st[2]))
            return np.ascontiguousarray(np.transpose(y.numpy(), (0, 2, 3, 1)))
        if op == "ResizeNearestNeighbor":
            assert a.get("half_pixel_centers", False) and not a.get("align_corners", False)
            H, W = x[0].shape[1:3]
            oh, ow = int(x[1][0]), int(x[1][1])
            # half_pixel_centers: src = floor((dst + 0.5) * in/out)
            ri = np.minimum(np.floor((np.arange(oh) + 0.5) * (H / oh)).astype(int), H - 1)
            ci = np.minimum(np.floor((np.arange(ow) + 0.5) * (W / ow)).astype(int), W - 1)
            return x[0][:, ri][:, :, ci]
        if op == "ReverseSequence":
            sd, bd = int(a["seq_dim"]), int(a.get("batch_dim", 0))
            out = x[0].copy()
            for b in range(x[0].shape[bd]):
                L = int(x[1][b])
                sl = [slice(None)] * x[0].ndim
                sl[bd] = b
                src = x[0][tuple(sl)]
                sd2 = sd - (1 if bd < sd else 0)
                idx = [slice(None)] * src.ndim
                idx[sd2] = slice(0, L)
                rev = np.flip(src[tuple(idx)], axis=sd2)
                dst = out[tuple(sl)]
                dst[tuple(idx)] = rev
            return out
        if op == "TensorArrayV3":
            return (TensorArray(), np.float32(0))
        if op == "TensorArrayScatterV3":
            ta = x[0]
            ta.items = {int(i): x[2][k] for k, i in enumerate(x[1])}
            return np.float32(0)
        if op == "TensorArrayReadV3":
            return x[0].items[int(x[1])]
        if op == "RandomUniform":
            raise RuntimeError("RandomUniform reached on a live branch: " + name)
        raise NotImplementedError(op + " @ " + name)
