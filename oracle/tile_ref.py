"""TEST INFRASTRUCTURE -- synthetic RAW tile (what temp/<x>/<y>/raw/*.hkl hold) for process_tile
(/root/reference/src/download_and_predict_job.py:640-997) and a dict-backed stand-in for hkl.load.
The golden outputs in tests/golden/process_tile.npz come from the reference function itself
(tools/make_golden_tile.py).  Tests only."""
import numpy as np


from sentinel_tree_cover_b200.synth import synth_raw_tile  # noqa: E402,F401  (seeded generator shared with bench.py)


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
