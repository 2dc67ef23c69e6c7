"""TEST INFRASTRUCTURE (oracle) -- CPU restatement of the reference's two frozen
models.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this; the product path never does.

What it restates (the reference executes these through TensorFlow:
src/download_and_predict_job.py:353-357 `sess.run(predict_logits, ...)` and
:115-117 `sess.run(superresolve_logits, ...)`):

* predict_graph-<H>.pb : bidirectional ConvGRU (pb:down_16/bidirectional_rnn/*,
  readable spec src/train/src/model.py:208-290,540-579) + U-Net of
  partial-conv/Swish/GroupNorm/sSE blocks (pb:conv_median ... pb:conv2d/Sigmoid,
  readable spec src/train/train-model.py:140-231, src/train/src/model.py:100-121,396-538).
* superresolve_graph.pb : DSen2-style residual CNN (pb:in_conv ... pb:Add_2).

Pinning: TensorFlow is not installable here and the reference holds no golden
model outputs, so this restatement is "parity unpinned" against real TF; it IS
pinned against oracle/tfgraph_interp.py, a mechanical op-by-op interpreter of
the same GraphDefs (tests/test_oracle_model.py, fixtures in tests/golden/).
Arithmetic is torch-CPU float32 (or float64 with dtype=torch.float64).
"""
import numpy as np
import torch
import torch.nn.functional as F

GN_EPS = 1e-5


def _t(a, dtype):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dtype)


_QUANT = [None]   # set by PredictRef/SuperresolveRef(quant="fp16"): emulate fp16 conv operands


def _q(t):
    return t.to(torch.float16).to(t.dtype) if _QUANT[0] == "fp16" else t


def _conv(x, w, pad):
    """x NCHW, w HWIO numpy->OIHW. pad: 'zero' (SAME), 'reflect', 'valid'.
    With quant="fp16" both operands are rounded to fp16 first (fp32 accumulate) -- the
    arithmetic contract of the tensor-core path, used for tight GPU comparisons."""
    x, w = _q(x), _q(w)
    if pad == "reflect":
        x = F.pad(x, (1, 1, 1, 1), mode="reflect")
        return F.conv2d(x, w)
    if pad == "zero":
        return F.conv2d(x, w, padding=1)
    return F.conv2d(x, w)


def _gn(x, gamma, beta, groups=8):
    """pb:*_norm/*: biased moments over (C/8,H,W); (x-mu)/sqrt(var+1e-5)*gamma+beta."""
    B, C, H, W = x.shape
    xg = x.reshape(B, groups, C // groups, H, W)
    mu = xg.mean(dim=(2, 3, 4), keepdim=True)
    var = ((xg - mu) ** 2).mean(dim=(2, 3, 4), keepdim=True)
    y = ((xg - mu) / torch.sqrt(var + GN_EPS)).reshape(B, C, H, W)
    return y * gamma.view(1, C, 1, 1) + beta.view(1, C, 1, 1)


def _partial_scale(H, W, dtype):
    """pb:<blk>_conv/<blk>/mask/mul: 9/(cnt+1e-8)*clip(cnt,0,1), cnt = 3x3 SAME count."""
    cnt = F.conv2d(torch.ones(1, 1, H, W, dtype=dtype), torch.ones(1, 1, 3, 3, dtype=dtype), padding=1)
    return 9.0 / (cnt + 1e-8) * torch.clamp(cnt, 0, 1)


class PredictRef:
    def __init__(self, weights, dtype=torch.float32, quant=None):
        self.dt = dtype
        self.quant = quant
        self.w = {}
        for k, v in weights.items():
            if v.ndim == 4:
                self.w[k] = _t(np.transpose(v, (3, 2, 0, 1)), dtype)  # HWIO -> OIHW
            else:
                self.w[k] = _t(v, dtype)

    # -- ConvGRU cell, pb:.../while/<d>/conv_gru_cell/* -------------------------
    def _cell(self, d, x, h):
        w = self.w
        g = _conv(torch.cat([x, h], 1), w["gru.%s.gates_w" % d], "reflect")
        r, u = g[:, :32], g[:, 32:]
        r = torch.sigmoid(_gn(r, w["gru.%s.r_gamma" % d], w["gru.%s.r_beta" % d]))
        u = torch.sigmoid(_gn(u, w["gru.%s.u_gamma" % d], w["gru.%s.u_beta" % d]))
        y = _conv(torch.cat([x, r * h], 1), w["gru.%s.cand_w" % d], "reflect")
        s = torch.sigmoid((y * w["gru.%s.cand_sse_w" % d].view(1, 32, 1, 1)).sum(1, keepdim=True))
        y = _gn(y * s, w["gru.%s.y_gamma" % d], w["gru.%s.y_beta" % d])
        return u * h + (1 - u) * torch.tanh(y)

    def gru(self, seq, length):
        """seq [B,T,17,H,W], length [B] -> [B,64,H,W] = concat(fw_T, bw_T).
        Zoneout at inference (pb:.../while/{mul,mul_1,add_1}): h = .75 h + .25 h~;
        steps t >= length[b] keep the state (pb:.../while/Select_1); bw consumes
        ReverseSequence(seq, length)."""
        B, T = seq.shape[:2]
        outs = []
        steps = int(min(T, int(length.max())))
        for d in ("fw", "bw"):
            h = torch.zeros(B, 32, seq.shape[3], seq.shape[4], dtype=self.dt)
            for t in range(steps):
                if d == "fw":
                    x = seq[:, t]
                else:
                    idx = torch.where(t < length, length - 1 - t, torch.full_like(length, t))
                    x = seq[torch.arange(B), idx]
                hn = 0.75 * h + 0.25 * self._cell(d, x, h)
                keep = (t >= length).view(B, 1, 1, 1)
                h = torch.where(keep, h, hn)
            outs.append(h)
        return torch.cat(outs, 1)

    # -- conv block, pb:<blk>_conv .. csse_<blk>_mul ---------------------------
    def block(self, name, x, same, taps=None):
        w = self.w
        y = _conv(x, w[name + ".w"], "zero" if same else "valid")
        if same:
            y = y * _partial_scale(x.shape[2], x.shape[3], self.dt)
        y = y * torch.sigmoid(y)
        y = _gn(y, w[name + ".gamma"], w[name + ".beta"])
        s = torch.sigmoid((y * w[name + ".sse_w"].view(1, -1, 1, 1)).sum(1, keepdim=True) + w[name + ".sse_b"])
        out = y * s
        if taps is not None:
            taps[name] = out
        return out

    def forward(self, x, length=None, taps=None):
        """x [B,T+1,H,W,17] float (normalised), frames 0..T-1 sequence, frame T median.
        Returns [B,H-14,W-14] probabilities (pb:conv2d/Sigmoid)."""
        _QUANT[0] = self.quant
        x = _t(x, self.dt).permute(0, 1, 4, 2, 3)
        B, T1 = x.shape[:2]
        T = T1 - 1
        if length is None:
            length = np.full(B, T)
        length = torch.as_tensor(np.asarray(length), dtype=torch.long)
        gru = self.gru(x[:, :T], length)
        if taps is not None:
            taps["gru"] = gru
        med = self.block("conv_median", x[:, T], True, taps)
        cc = self.block("conv_concat", torch.cat([gru, med], 1), True, taps)
        c1 = self.block("conv1", F.max_pool2d(cc, 2), False, taps)
        c2 = self.block("conv2", F.max_pool2d(c1, 2), False, taps)
        u2 = self.block("up2", F.interpolate(c2, scale_factor=2, mode="nearest"), True, taps)
        u2o = self.block("up2_out", torch.cat([u2, c1[:, :, 2:-2, 2:-2]], 1), True, taps)
        u3 = self.block("up3", F.interpolate(u2o, scale_factor=2, mode="nearest"), True, taps)
        o = self.block("out", torch.cat([u3, cc[:, :, 6:-6, 6:-6]], 1), False, taps)
        logit = (o * self.w["head.w"].view(1, 64, 1, 1)).sum(1) + self.w["head.b"]
        _QUANT[0] = None
        return torch.sigmoid(logit).numpy()


class SuperresolveRef:
    """pb:superresolve_graph: [reflect-pad 1 + conv3x3 + bias] x6 with
    ReLU / x0.1 residuals, tanh, + bilinear input (Placeholder_1)."""

    def __init__(self, weights, dtype=torch.float32, quant=None):
        self.dt = dtype
        self.quant = quant
        self.w = {k: (_t(np.transpose(v, (3, 2, 0, 1)), dtype) if v.ndim == 4 else _t(v, dtype))
                  for k, v in weights.items()}

    def _c(self, n, x):
        x = F.pad(x, (1, 1, 1, 1), mode="reflect")
        if self.quant == "fp16":
            h = lambda t: t.to(torch.float16).to(t.dtype)
            return F.conv2d(h(x), h(self.w["sr.%s.w" % n]), self.w["sr.%s.b" % n])
        return F.conv2d(x, self.w["sr.%s.w" % n], self.w["sr.%s.b" % n])

    def forward(self, x10, bilinear6):
        """x10 [N,H,W,10], bilinear6 [N,H,W,6] -> [N,H,W,6] (pb:Add_2)."""
        x = _t(x10, self.dt).permute(0, 3, 1, 2)
        a = torch.relu(self._c("in", x))
        b = a + 0.1 * self._c("r02", torch.relu(self._c("r01", a)))
        c = b + 0.1 * self._c("r12", torch.relu(self._c("r11", b)))
        y = torch.tanh(self._c("out", c))
        return (y.permute(0, 2, 3, 1) + _t(bilinear6, self.dt)).numpy()
