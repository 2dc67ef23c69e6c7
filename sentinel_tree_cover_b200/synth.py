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


def synth_cloudy_cube(T, H, W, seed, urban=False):
    """Sentinel-2-like cube [T,H,W,10] with vegetation/soil/water spectra, Gaussian-blob clouds
    (+0.3..0.6 on every band) and displaced shadows (x0.3), plus a DEM [H,W].
    urban=True adds a built-up block (NDBI > 0, NDBI > NDVI) with bright roofs that look like small clouds."""
    r = np.random.default_rng(seed)
    veg = np.array([0.035, 0.06, 0.045, 0.30, 0.10, 0.22, 0.28, 0.32, 0.17, 0.08], np.float32)
    soil = np.array([0.09, 0.12, 0.15, 0.25, 0.18, 0.21, 0.23, 0.26, 0.30, 0.24], np.float32)
    wat = np.array([0.05, 0.06, 0.04, 0.02, 0.03, 0.025, 0.02, 0.02, 0.01, 0.008], np.float32)
    yy, xx = np.mgrid[0:H, 0:W]
    mix = 0.5 + 0.5 * np.sin(xx / 17.0) * np.cos(yy / 23.0)
    base = veg[None, None] * mix[..., None] + soil[None, None] * (1 - mix[..., None])
    lake = (yy - H * 0.7) ** 2 + (xx - W * 0.25) ** 2 < (min(H, W) * 0.12) ** 2
    base[lake] = wat
    if urban:
        town = (np.abs(yy - H * 0.3) < H * 0.16) & (np.abs(xx - W * 0.65) < W * 0.2)
        base[town] = np.array([0.10, 0.12, 0.16, 0.20, 0.19, 0.20, 0.21, 0.22, 0.28, 0.24], np.float32)
        roofs = town & (((yy // 3) % 5 == 0) & ((xx // 3) % 4 == 0))
        base[roofs] += 0.22
    cube = np.repeat(base[None], T, 0) * (1 + 0.08 * np.sin(2 * np.pi * np.arange(T) / T))[:, None, None, None]
    cube = cube + r.normal(0, 0.004, cube.shape)
    for t in range(T):
        for _ in range(r.integers(0, 3)):
            cy, cx, s = r.integers(0, H), r.integers(0, W), r.uniform(5, 14)
            blob = np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * s * s))
            cube[t] += (r.uniform(0.3, 0.6) * (blob > 0.4))[..., None]
            sy, sx = cy + int(1.5 * s), cx + int(1.2 * s)
            sh = np.exp(-((yy - sy) ** 2 + (xx - sx) ** 2) / (2 * s * s)) > 0.45
            cube[t][sh] *= 0.3
    dem = (40 * (0.5 + 0.5 * np.sin(xx / 31.0 + yy / 47.0))).astype(np.float32)
    return np.clip(cube, 0.001, 0.999).astype(np.float32), dem


def synth_raw_tile(seed, n=8, h=60, w=64, with_clm=False, ragged=False):
    """uint16 S2 10 m / 20 m stacks, uint16 S1, float32 DEM, dates; `ragged` makes S1 / DEM / 10 m one
    pixel larger or smaller than 2x the 20 m grid so that adjust_shape has work to do."""
    img, dem = synth_cloudy_cube(n, 2 * h, 2 * w, seed)
    r = np.random.default_rng(seed + 1000)
    raw = {}
    s2_10 = np.trunc(img[..., :4] * 65535).astype(np.uint16)
    s2_20 = np.trunc(img[:, ::2, ::2, 4:10] * 65535).astype(np.uint16)
    s1 = np.trunc(r.uniform(0.01, 0.6, (12, 2 * h, 2 * w, 2)) * 65535).astype(np.uint16)
    s1[r.random(s1.shape) < 0.001] = 65535                       # saturated returns -> median fill
    demf = (dem * 20 + r.normal(0, 3, dem.shape)).astype(np.float32)
    if ragged:
        s2_10 = np.pad(s2_10, ((0, 0), (0, 1), (1, 1), (0, 0)), "edge")      # 2h+1 x 2w+2
        s1 = s1[:, 1:-1, 2:-2]                                                # 2h-2 x 2w-4 (padded back with 'edge')
        demf = np.pad(demf, ((2, 2), (0, 0)), "edge")                         # 2h+4 x 2w
    raw["clouds"] = r.uniform(0, 0.3, (n, h // 4, w // 4)).astype(np.float32)
    raw["s1"], raw["s2_10"], raw["s2_20"], raw["dem"] = s1, s2_10, s2_20, demf
    raw["s2_dates"] = (np.arange(n) * (330 // n) + 10).astype(np.int64)
    if with_clm:
        c = np.zeros((n, h, w), np.float32)
        c[2:4, 10:20, 10:30] = 1.0                                 # two consecutive dates -> cleared by the pair rule
        c[5, 30:40, 5:25] = 1.0                                    # single-date Sen2Cor cloud -> kept
        c[0, 2:6, 40:50] = 1.0
        raw["cloudmask"] = c
    return raw
