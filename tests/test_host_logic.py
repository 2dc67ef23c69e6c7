"""CPU: library loading / ABI surface, sharding (world_size 2 over gloo)."""
import os
import re
import subprocess
import sys
import numpy as np
import pytest

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def test_library_builds_loads_and_exports_every_declared_symbol():
    from sentinel_tree_cover_b200 import build, api
    lib_path = build.build()
    assert os.path.exists(lib_path)
    lib = api.load_library()
    header = open(os.path.join(ROOT, "include", "stc.h")).read()
    declared = set(re.findall(r"\b(stc_[a-z0-9_]+)\s*\(", header))
    bound = {s[0] for s in api.SYMBOLS}
    assert declared == bound, (declared ^ bound)
    for name in declared:
        assert hasattr(lib, name)
    assert b"sm_100a" in lib.stc_version()


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from sentinel_tree_cover_b200.api import StcSession
    with pytest.raises(RuntimeError):
        StcSession(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "sentinel_tree_cover_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M) and "oracle." not in src, fn


def test_shard_range_partitions():
    from sentinel_tree_cover_b200.shard import shard_range
    for n in (0, 1, 7, 190, 36100):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


_WORKER = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from sentinel_tree_cover_b200.shard import shard_range, broadcast_weights
from sentinel_tree_cover_b200.weights import random_predict_weights
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
w = random_predict_weights(5) if rank == 0 else None
w = broadcast_weights(w, dist)
ref = random_predict_weights(5)
assert sorted(w) == sorted(ref) and all(np.array_equal(w[k], ref[k]) for k in ref)
lo, hi = shard_range(37, rank, world)
t = torch.tensor([float(hi - lo)]); dist.all_reduce(t)
assert int(t.item()) == 37
tm = torch.tensor([1.0 + rank]); dist.all_reduce(tm, op=dist.ReduceOp.MAX)
assert tm.item() == float(world)
dist.barrier(); dist.destroy_process_group()
print("OK", rank)
'''


def test_weight_broadcast_and_sharding_world2_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29617", str(script), ROOT]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("OK") == 2


def test_python_random_shuffle_replay_matches_cpython():
    """The cloud-removal fit samples its pixels with Python's global `random` (cloud_removal.py:447-497); the library
    replays random.shuffle from the generator state on the host (MT19937 block regeneration, branch-free rejection
    loop).  Host-only entry point: runs without a GPU."""
    import ctypes as C
    import random
    from sentinel_tree_cover_b200 import api
    lib = api.load_library()
    random.seed(20240917)
    for _ in range(411):                                     # land in the middle of a 624-word block
        random.getrandbits(32)
    for n in (0, 1, 2, 3, 7, 31, 32, 33, 34, 63, 64, 65, 100, 1024, 1025, 65535, 65536, 65537, 200003, 1300001):
        state = np.array(random.getstate()[1], dtype=np.uint32)
        skip = state.copy()
        want = list(range(n)); random.shuffle(want)
        # data == NULL: the generator walk alone (what lets the shuffles run on worker threads) ends in the same state
        assert lib.stc_py_shuffle(skip.ctypes.data_as(C.c_void_p), None, n) == 0
        assert tuple(int(v) for v in skip) == random.getstate()[1], ("skip", n)
        got = np.arange(n, dtype=np.int32)
        rc = lib.stc_py_shuffle(state.ctypes.data_as(C.c_void_p), got.ctypes.data_as(C.c_void_p), n)
        assert rc == 0 and got.tolist() == want, n
        assert tuple(int(v) for v in state) == random.getstate()[1], n      # the generator is left where Python leaves it


@pytest.mark.parametrize("isa", ["0", "1"])
def test_generator_walk_on_every_instruction_set(isa):
    """The data-less walk (skip_shuffle) has an AVX-512, an AVX2 and a scalar scan; the default run takes the widest the
    machine has.  STC_PYRANDOM_ISA caps it, so the other two are walked here in a child process."""
    code = r'''
import ctypes as C, random, numpy as np
from sentinel_tree_cover_b200 import api
lib = api.load_library()
random.seed(77)
for _ in range(100): random.getrandbits(32)
for n in (5, 255, 256, 257, 4095, 4096, 16384, 70001, 600000):
    skip = np.array(random.getstate()[1], dtype=np.uint32)
    want = list(range(n)); random.shuffle(want)
    assert lib.stc_py_shuffle(skip.ctypes.data_as(C.c_void_p), None, n) == 0
    assert tuple(int(v) for v in skip) == random.getstate()[1], n
print("OK")
'''
    env = dict(os.environ, STC_PYRANDOM_ISA=isa, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr


def test_header_is_plain_c_and_the_library_links_from_c(tmp_path):
    """The drop-in boundary is a C ABI: include/stc.h compiles as strict C99, and a C program linked against libstc.so calls
    two host-only entry points (GeoTIFF encode -> decode round trip; no GPU, no Python in between)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    from sentinel_tree_cover_b200 import build
    lib = build.build()
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
    src = tmp_path / "client.c"
    src.write_text(r'''
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include "stc.h"
int main(void) {
  uint8_t img[6 * 5], *file = 0, *back = 0;
  int64_t len = 0;
  int rows = 0, cols = 0, i;
  double b[4];
  for (i = 0; i < 30; ++i) img[i] = (uint8_t)(i * 7);
  if (stc_geotiff_encode_u8(img, 6, 5, 10.0, -3.0, 10.5, -2.4, &file, &len) != STC_OK) return 1;
  if (stc_geotiff_decode_u8(file, len, 1, &back, &rows, &cols, b) != STC_OK) return 2;
  if (rows != 6 || cols != 5 || memcmp(img, back, 30) != 0) return 3;
  if (b[0] != 10.0 || b[1] > -2.999999 || b[2] < 10.499999 || b[3] != -2.4) return 4;
  if (stc_geotiff_encode_u8(0, 6, 5, 0, 0, 1, 1, &file, &len) != STC_ERR_ARG) return 5;
  stc_geotiff_free(file); stc_geotiff_free(back);
  printf("ok %lld\n", (long long)len);
  return 0;
}
''')
    exe = str(tmp_path / "client")
    inc = os.path.join(root, "include")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", inc, str(src), "-o", exe,
                    lib, "-Wl,-rpath," + os.path.dirname(lib)], check=True, capture_output=True)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and r.stdout.startswith("ok "), (r.returncode, r.stdout, r.stderr)
