"""CPU study: effect of storage precisions on the probability map (released weights, H=172).
Variants on top of fp16 conv operands + fp16 GRU pre-norm tensors (the shipped configuration):
  h16   : GRU state kept only in fp16 between steps (no fp32 copy)
  raw16 : U-Net pre-norm conv outputs stored in fp16
Not part of the product or the tests; documents the numbers quoted in DESIGN.md."""
import os, sys
import numpy as np
import torch
import torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import model_ref as M
from oracle import preproc_ref as P
from sentinel_tree_cover_b200.weights import load_npz
from sentinel_tree_cover_b200.api import MIN_ALL, MAX_ALL


def h16(t):
    return t.to(torch.float16).to(torch.float32)


class Variant(M.PredictRef):
    def __init__(self, w, g16=True, state16=False, raw16=False):
        super().__init__(w, quant="fp16")
        self.g16, self.state16, self.raw16 = g16, state16, raw16

    def _cell(self, d, x, h):
        w = self.w
        g = M._conv(torch.cat([x, h], 1), w["gru.%s.gates_w" % d], "reflect")
        if self.g16:
            # statistics from the fp32 accumulators (epilogue), normalisation applied to the fp16 copy
            r32, u32 = g[:, :32], g[:, 32:]
            g = h16(g)
        r, u = g[:, :32], g[:, 32:]
        def gn_from(x16, x32, gamma, beta):
            B, C, H, W = x32.shape
            xg = x32.reshape(B, 8, C // 8, H, W)
            mu = xg.mean(dim=(2, 3, 4), keepdim=True)
            var = (xg ** 2).mean(dim=(2, 3, 4), keepdim=True) - mu ** 2
            y = ((x16.reshape(B, 8, C // 8, H, W) - mu) / torch.sqrt(var + M.GN_EPS)).reshape(B, C, H, W)
            return y * gamma.view(1, C, 1, 1) + beta.view(1, C, 1, 1)
        if self.g16:
            r = torch.sigmoid(gn_from(r, r32, w["gru.%s.r_gamma" % d], w["gru.%s.r_beta" % d]))
            u = torch.sigmoid(gn_from(u, u32, w["gru.%s.u_gamma" % d], w["gru.%s.u_beta" % d]))
        else:
            r = torch.sigmoid(M._gn(r, w["gru.%s.r_gamma" % d], w["gru.%s.r_beta" % d]))
            u = torch.sigmoid(M._gn(u, w["gru.%s.u_gamma" % d], w["gru.%s.u_beta" % d]))
        y = M._conv(torch.cat([x, r * h], 1), w["gru.%s.cand_w" % d], "reflect")
        s = torch.sigmoid((y * w["gru.%s.cand_sse_w" % d].view(1, 32, 1, 1)).sum(1, keepdim=True))
        y = y * s
        if self.g16:
            y = gn_from(h16(y), y, w["gru.%s.y_gamma" % d], w["gru.%s.y_beta" % d])
        else:
            y = M._gn(y, w["gru.%s.y_gamma" % d], w["gru.%s.y_beta" % d])
        return u * h + (1 - u) * torch.tanh(y)

    def gru(self, seq, length):
        B, T = seq.shape[:2]
        outs = []
        for d in ("fw", "bw"):
            h = torch.zeros(B, 32, seq.shape[3], seq.shape[4])
            for t in range(T):
                x = seq[:, t] if d == "fw" else seq[:, T - 1 - t]
                h = 0.75 * h + 0.25 * self._cell(d, x, h)
                if self.state16:
                    h = h16(h)
            outs.append(h)
        return torch.cat(outs, 1)

    def block(self, name, x, same, taps=None):
        w = self.w
        y = M._conv(x, w[name + ".w"], "zero" if same else "valid")
        if same:
            y = y * M._partial_scale(x.shape[2], x.shape[3], self.dt)
        y = y * torch.sigmoid(y)
        if self.raw16:
            B, C, H, W = y.shape
            xg = y.reshape(B, 8, C // 8, H, W)
            mu = xg.mean(dim=(2, 3, 4), keepdim=True)
            var = (xg ** 2).mean(dim=(2, 3, 4), keepdim=True) - mu ** 2
            yn = ((h16(y).reshape(B, 8, C // 8, H, W) - mu) / torch.sqrt(var + M.GN_EPS)).reshape(B, C, H, W)
            y = yn * w[name + ".gamma"].view(1, C, 1, 1) + w[name + ".beta"].view(1, C, 1, 1)
        else:
            y = M._gn(y, w[name + ".gamma"], w[name + ".beta"])
        s = torch.sigmoid((y * w[name + ".sse_w"].view(1, -1, 1, 1)).sum(1, keepdim=True) + w[name + ".sse_b"])
        return y * s


def main():
    torch.set_num_threads(8)
    w = load_npz(os.path.join(ROOT, "tests", "golden", "weights_predict_172.npz"))
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    m = P.synth_monthly(n, 172, 7)
    x = np.stack([P.normalize_subtile(P.assemble(m[i:i + 1])[0], MIN_ALL, MAX_ALL) for i in range(n)])
    ref = M.PredictRef(w).forward(x)
    for name, kw in [("shipped (fp16 operands, fp16 GRU pre-norm)", {}),
                     ("+ fp16-only GRU state", {"state16": True}),
                     ("+ fp16 U-Net pre-norm", {"raw16": True}),
                     ("+ both", {"state16": True, "raw16": True})]:
        y = Variant(w, **kw).forward(x)
        d = np.abs(y - ref)
        print("%-48s max|dp| %.2e  p99.9 %.2e  mean %.2e" % (name, d.max(), np.quantile(d, 0.999), d.mean()), flush=True)


if __name__ == "__main__":
    main()
