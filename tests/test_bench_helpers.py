"""bench.py bookkeeping that does not need a GPU: the algorithmic FLOP / byte counts behind the roofline entries
(known answers from SURVEY.md section 8d) and the shape of the roofline objects."""
import os
import sys
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def test_conv_flops_match_survey_figures():
    # ~40.0 GFLOP per tile at H = 168 (GRU 19.13 + blocks 20.9), ~41.98 at H = 172 (T = 4)
    assert abs(bench.conv_flops_per_tile(168) / 1e9 - 40.0) < 0.1
    assert abs(bench.conv_flops_per_tile(172) / 1e9 - 41.98) < 0.1
    gru = 2 * 4 * 84736 * 168 * 168
    assert abs(gru / 1e9 - 19.13) < 0.01


def test_gates_roofline_object():
    peaks = {"tf": 1364.4, "hbm": 6541.1, "src": "measured"}
    r = bench.gates_roofline(total_ms=40.0, n_launch=320, chunk=32, peaks=peaks)     # 125 us per launch
    flops = 2 * 32 * 168 * 168 * (3 * 2 * 9 * 49 * 64 + 2 * 9 * 17 * 64) / 4.0
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and r["peak"] == 1364.4
    assert abs(r["achieved"] - flops / 125e-6 / 1e12) < 1e-6 and abs(r["frac"] - r["achieved"] / 1364.4) < 1e-12
    assert bench.gates_roofline(0.0, 0, 32, peaks)["achieved"] is None


def test_hbm_roofline_object(tmp_path):
    peaks = {"tf": 1364.4, "hbm": 6541.1, "src": "measured"}
    p = tmp_path / "trace.csv"
    p.write_text("label,slot,start_ms,end_ms\napply2,0,0.0,0.160\nconv_gates,0,0.2,0.3\napply2,0,1.0,1.164\n")
    r = bench.hbm_roofline(str(p), 32, peaks)
    by = 2 * 32 * 168 * 168 * ((128 + 3 * 256) / 4.0 + (3 * 192 + 256) / 4.0)       # 432 B per pixel and direction
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["achieved"] - by / 0.162e-3 / 1e9) < 1e-3
    assert 0 < r["frac"] < 1
    p.write_text("label,slot,start_ms,end_ms\nconv_gates,0,0.2,0.3\n")
    assert bench.hbm_roofline(str(p), 32, peaks) is None


def test_parity_object_and_its_cpu_leg():
    """The `parity` entry of the bench line: the CPU child leg writes the oracle maps for the first tile of the timed batch on
    the float32 patches and on the uint16-stored patches; parity_object compares each GPU wire format with ITS oracle."""
    r = bench.cpu_parity_outputs(2000, 1)
    want = np.load(r["path"])
    assert want.shape == (2, 1, 154, 154) and want.dtype == np.float32 and 0 <= want.min() and want.max() <= 1
    ok = bench.parity_object(want[0] + np.float32(2e-4), want[1] - np.float32(3e-4), want)
    assert ok["ok"] and abs(ok["max_abs_err"] - 2e-4) < 1e-6 and abs(ok["max_abs_err_uint16_wire"] - 3e-4) < 1e-6 and ok["tiles"] == 1
    assert ok["oracle_shift_under_uint16_storage"] == float(np.abs(want[0] - want[1]).max())
    bad = bench.parity_object(want[0], want[1] + np.float32(2e-3), want)
    assert not bad["ok"]
    u = bench.quantise_u16(np.array([0.0, 0.5, 1.0, 1.5, -0.1], np.float32))
    assert u.dtype == np.uint16 and u.tolist() == [0, 32768, 65535, 65535, 0]
