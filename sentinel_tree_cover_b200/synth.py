"""Seeded synthetic inputs of the shapes the hot path consumes: used by bench.py, the tools and,
through oracle/preproc_ref.py, by the tests.  Host-side NumPy only; nothing here is on the arithmetic path."""
import numpy as np


def synth_monthly(B, H, seed, W=None):
    """[B,12,H,W,13] float32: 10 S2 bands (smooth seasonal signal), DEM, S1 VV/VH."""
    W = H if W is None else W
    r = np.random.default_rng(seed)
    base = r.uniform(0.02, 0.45, (B, 1, H, W, 10)).astype(np.float32)
    phase = r.uniform(0, 2 * np.pi, (B, 1, H, W, 1)).astype(np.float32)
    t = np.arange(12, dtype=np.float32).reshape(1, 12, 1, 1, 1)
    s2 = base + 0.05 * np.sin(2 * np.pi * (t / 12) + phase) + r.normal(0, 0.01, (B, 12, H, W, 10)).astype(np.float32)
    s2 = np.clip(s2, 0.001, 0.999).astype(np.float32)
    dem = np.repeat(r.uniform(0, 0.4, (B, 1, H, W, 1)).astype(np.float32), 12, axis=1)
    s1 = r.uniform(0.05, 0.95, (B, 12, H, W, 2)).astype(np.float32)
    return np.ascontiguousarray(np.concatenate([s2, dem, s1], -1), np.float32)
