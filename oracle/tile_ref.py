"""TEST INFRASTRUCTURE -- synthetic RAW tile (what temp/<x>/<y>/raw/*.hkl hold) for process_tile
(/root/reference/src/download_and_predict_job.py:640-997) and a dict-backed stand-in for hkl.load.
The golden outputs in tests/golden/process_tile.npz come from the reference function itself
(tools/make_golden_tile.py).  Tests only."""
import numpy as np
from oracle import cloud_ref


def synth_raw_tile(seed, n=8, h=60, w=64, with_clm=False, ragged=False):
    """uint16 S2 10 m / 20 m stacks, uint16 S1, float32 DEM, dates; `ragged` makes S1 / DEM / 10 m one
    pixel larger or smaller than 2x the 20 m grid so that adjust_shape has work to do."""
    img, dem = cloud_ref.synth_cloudy_cube(n, 2 * h, 2 * w, seed)
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


class FakeStore:
    """loader / exists pair keyed on the reference's file naming."""
    KEYS = [("raw/clouds/clouds_", "clouds"), ("raw/clouds/cloudmask_", "cloudmask"), ("raw/s1/", "s1"), ("raw/s2_10/", "s2_10"),
            ("raw/s2_20/", "s2_20"), ("raw/misc/s2_dates_", "s2_dates"), ("raw/misc/dem_", "dem")]

    def __init__(self, raw):
        self.raw = raw

    def _key(self, path):
        for frag, k in self.KEYS:
            if frag in path:
                return k
        raise KeyError(path)

    def load(self, path):
        return np.copy(self.raw[self._key(path)])

    def exists(self, path):
        try:
            return self._key(path) in self.raw
        except KeyError:
            return False
