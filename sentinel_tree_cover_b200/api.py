"""ctypes binding of libstc.so + host-side mirror of the reference's hot-path
functions (same names, argument meaning, in-place / return contracts).

Reference functions mirrored (all in /root/reference/src/download_and_predict_job.py
unless noted):
  normalize_subtile(subtile)                        :316-325  (in place)
  predict_subtile(subtile, sess, op, size)          :328-369
  make_indices(arr)                                 :998-1006
  smooth_large_tile(arr, dates, interp)             :1057-1096
  superresolve_large_tile(arr, sess)                :95-147   (in place on arr[..., 4:])
The `sess` slot takes a StcSession (replaces the tf.compat.v1.Session globals,
:1785-1826).  There is NO CPU fallback: if libstc.so or a CUDA sm_100 device is
missing, StcSession() raises.
"""
import ctypes as C
import os
import numpy as np

from . import weights as _weights
from . import regrid as _regrid

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libstc.so")
_lib = None

# normalisation constants of the released model, src/download_and_predict_job.py:1829-1842
MIN_ALL = [0.006576638437476157, 0.0162050812542916, 0.010040436408026246,
           0.013351644159609368, 0.01965362020294499, 0.014229037918669413,
           0.015289539940489814, 0.011993591210803388, 0.008239871824216068,
           0.006546120393682765, 0.0, 0.0, 0.0, -0.1409399364817101,
           -0.4973397113668104, -0.09731556326714398, -0.7193834232943873]
MAX_ALL = [0.2691233691920348, 0.3740291447318227, 0.5171435111009385,
           0.6027466239414053, 0.5650263218127718, 0.5747005416952773,
           0.5933928435187305, 0.6034943160143434, 0.7472037842374304,
           0.7000076295109483, 0.4, 0.948334642387533,
           0.6729257769285485, 0.8177635298774327, 0.35768999002433816,
           0.7545951919107605, 0.7602693339366691]

_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)

# every symbol include/stc.h declares: (name, restype, argtypes)
SYMBOLS = [
    ("stc_create", C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    ("stc_destroy", None, [C.c_void_p]),
    ("stc_last_error", C.c_char_p, [C.c_void_p]),
    ("stc_version", C.c_char_p, []),
    ("stc_launch_count", C.c_int64, [C.c_void_p]),
    ("stc_set_conv_impl", C.c_int, [C.c_void_p, C.c_int]),
    ("stc_set_weight", C.c_int, [C.c_void_p, C.c_char_p, _f32p, C.c_int64]),
    ("stc_finalize_weights", C.c_int, [C.c_void_p, C.c_int]),
    ("stc_malloc", C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    ("stc_free", C.c_int, [C.c_void_p, C.c_void_p]),
    ("stc_malloc_host", C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]),
    ("stc_free_host", C.c_int, [C.c_void_p, C.c_void_p]),
    ("stc_h2d", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    ("stc_d2h", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    ("stc_sync", C.c_int, [C.c_void_p]),
    ("stc_timer_begin", C.c_int, [C.c_void_p]),
    ("stc_timer_end", C.c_int, [C.c_void_p, _f32p]),
    ("stc_conv_timing", C.c_int, [C.c_void_p, C.c_int, _f32p, C.POINTER(C.c_int64)]),
    ("stc_conv_timing_kind", C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, _f32p, C.POINTER(C.c_int64)]),
    ("stc_trace", C.c_int, [C.c_void_p, C.c_int, C.c_char_p]),
    ("stc_predict_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _f64p, _f64p, C.c_void_p]),
    ("stc_predict_feats_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _f64p, _f64p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("stc_float_to_int16_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    ("stc_feature_mosaic_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    ("stc_region_gather_dev", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    ("stc_region_predict_dev", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, _f64p, _f64p, C.c_void_p]),
    ("stc_region_blend_dev", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    ("stc_predict_dev", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _f64p, _f64p, C.c_void_p]),
    ("stc_assemble_dev", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    ("stc_assemble_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    ("stc_predict_patches_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, _f64p, _f64p, C.c_void_p]),
    ("stc_predict_patches_dev", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, _f64p, _f64p, C.c_void_p]),
    ("stc_predict_patches_u16_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, _f64p, _f64p, C.c_void_p]),
    ("stc_predict_patches_u16_dev", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, _f64p, _f64p, C.c_void_p]),
    ("stc_temporal_matmul_host", C.c_int, [C.c_void_p, C.c_void_p, _f32p, C.c_int, C.c_int, C.c_int64, C.c_void_p]),
    ("stc_temporal_matmul_dev", C.c_int, [C.c_void_p, C.c_void_p, _f32p, C.c_int, C.c_int, C.c_int64, C.c_void_p]),
    ("stc_indices_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]),
    ("stc_temporal_median_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p]),
    ("stc_superresolve_dev", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    ("stc_superresolve_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    ("stc_mosaic_diffs_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    ("stc_gauss_mosaic_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    ("stc_to_float32_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    ("stc_to_uint16_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    ("stc_convert_to_db_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_void_p]),
    ("stc_feather_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    ("stc_binary_dilate_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    ("stc_edt_sq_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    ("stc_missing_px_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    ("stc_median_fill_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    ("stc_build_sentinel2_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    ("stc_s1_fill_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]),
    ("stc_median_filter5_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    ("stc_clm_pairs_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]),
    ("stc_snow_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    ("stc_count_gt_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p]),
    ("stc_count_lt_axis0_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_void_p]),
    ("stc_elementwise_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_float]),
    ("stc_max_masked_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
    ("stc_s2_medians_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("stc_smooth_quarterly_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("stc_predict_postprocess_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                               _f64p, _f64p, C.c_void_p]),
    ("stc_process_subtiles_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, _f64p, _f64p,
                                            C.c_void_p, C.c_void_p]),
    ("stc_process_subtiles_feats_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                                  C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, _f64p, _f64p,
                                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("stc_np_sum_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    ("stc_normalize_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
    ("stc_bright_bare_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    ("stc_postprocess_subtile_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    ("stc_remove_clouds_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p]),
    ("stc_py_shuffle", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    ("stc_malloc_host_flags", C.c_int, [C.c_void_p, C.c_size_t, C.c_int, C.POINTER(C.c_void_p)]),
    ("stc_align_histograms_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    ("stc_order_stats_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]),
    ("stc_set_ancillary_masks_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]),
    ("stc_remove_clouds_clip_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.POINTER(C.c_int32)]),
    ("stc_cloud_masks_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_int]),
    ("stc_debug_read", C.c_int64, [C.c_void_p, C.c_char_p, C.c_void_p]),
    ("stc_tile_run_host", C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                    C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                    C.c_int, C.c_int, _f64p, _f64p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                    C.c_void_p]),
    ("stc_monthly_operator_plan", C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    ("stc_subtile_windows_plan", C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]),
    ("stc_adjust_shape_plan", C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    ("stc_pool_info", C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    ("stc_write_geotiff_u8", C.c_int, [C.c_char_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double]),
    ("stc_geotiff_encode_u8", C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                                        C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]),
    ("stc_geotiff_free", None, [C.c_void_p]),
    ("stc_read_geotiff_u8", C.c_int, [C.c_char_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    ("stc_geotiff_decode_u8", C.c_int, [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
]


def load_library(path=None):
    """dlopen libstc.so and bind every declared symbol.  Raises if missing."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or _LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError("libstc.so not built (%s); run `python __graft_entry__.py build`. "
                           "There is no CPU fallback." % p)
    lib = C.CDLL(p)
    for name, res, args in SYMBOLS:
        fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def _dptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _f64(a):
    a = np.ascontiguousarray(a, np.float64)
    return a, a.ctypes.data_as(_f64p)


class StcSession:
    """Opaque handle passed wherever the reference passes a tf.Session."""

    def __init__(self, device=0, predict_weights=None, superresolve_weights=None, conv_impl=None):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.stc_create(int(device), C.byref(h))
        if rc != 0:
            raise RuntimeError("stc_create(device=%d) failed with %d: a CUDA sm_100 (B200) device is required; "
                               "there is no CPU fallback" % (device, rc))
        self.h = h
        self.device = device
        self.length = 4            # args.length of the reference job (:354)
        self.min_all = list(MIN_ALL)
        self.max_all = list(MAX_ALL)
        if conv_impl is None and os.environ.get("STC_CONV_IMPL"):
            conv_impl = int(os.environ["STC_CONV_IMPL"])
        if conv_impl is not None:
            self._check(self.lib.stc_set_conv_impl(self.h, int(conv_impl)))
        if predict_weights is not None:
            self.load_predict(predict_weights)
        if superresolve_weights is not None:
            self.load_superresolve(superresolve_weights)

    # -- plumbing ----------------------------------------------------------------
    def _check(self, rc):
        if rc is not None and rc < 0:
            raise RuntimeError("libstc error %d: %s" % (rc, self.lib.stc_last_error(self.h).decode()))
        return rc

    def close(self):
        if getattr(self, "h", None):
            self.lib.stc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _set_weights(self, w, which):
        for k, v in w.items():
            v = np.ascontiguousarray(v, np.float32)
            self._check(self.lib.stc_set_weight(self.h, k.encode(), v.ctypes.data_as(_f32p), v.size))
        self._check(self.lib.stc_finalize_weights(self.h, which))

    def load_predict(self, src):
        """src: path to predict_graph-*.pb, path to .npz of canonical tensors, or dict."""
        if isinstance(src, str):
            src = _weights.load_predict_pb(src) if src.endswith(".pb") else _weights.load_npz(src)
        self._set_weights(src, 0)

    def load_superresolve(self, src):
        if isinstance(src, str):
            src = _weights.load_superresolve_pb(src) if src.endswith(".pb") else _weights.load_npz(src)
        self._set_weights(src, 1)

    def set_conv_impl(self, impl):
        self._check(self.lib.stc_set_conv_impl(self.h, int(impl)))

    def launch_count(self):
        return int(self.lib.stc_launch_count(self.h))

    # -- device memory (for benchmarks that keep inputs resident) ------------------
    def malloc(self, nbytes):
        p = C.c_void_p()
        self._check(self.lib.stc_malloc(self.h, nbytes, C.byref(p)))
        return p

    def free(self, p):
        self._check(self.lib.stc_free(self.h, p))

    def pinned_empty(self, shape, dtype=np.float32, write_combined=False):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        if write_combined:
            self._check(self.lib.stc_malloc_host_flags(self.h, n, 1, C.byref(p)))
        else:
            self._check(self.lib.stc_malloc_host(self.h, n, C.byref(p)))
        buf = (C.c_char * n).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
        self._pins = getattr(self, '_pins', [])
        self._pins.append(p)          # freed with the context's process; keeps the mapping alive
        return arr

    def h2d(self, dptr, arr):
        self._check(self.lib.stc_h2d(self.h, dptr, _dptr(arr), arr.nbytes))

    def d2h(self, arr, dptr):
        self._check(self.lib.stc_d2h(self.h, _dptr(arr), dptr, arr.nbytes))

    def sync(self):
        self._check(self.lib.stc_sync(self.h))

    def timer_begin(self):
        self._check(self.lib.stc_timer_begin(self.h))

    def timer_end(self):
        ms = C.c_float()
        self._check(self.lib.stc_timer_end(self.h, C.byref(ms)))
        return ms.value

    def conv_timing(self, enable_reset=-1):
        ms, n = C.c_float(), C.c_int64()
        self._check(self.lib.stc_conv_timing(self.h, enable_reset, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def conv_timing_kind(self, N, groups, mode):
        ms, n = C.c_float(), C.c_int64()
        self._check(self.lib.stc_conv_timing_kind(self.h, N, groups, mode, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def trace(self, enable, csv_path=None):
        """Kernel timeline of the model path (profiling aid, stc_trace)."""
        self._check(self.lib.stc_trace(self.h, int(enable), csv_path.encode() if csv_path else None))

    # -- model -------------------------------------------------------------------
    def predict(self, x, length=None, normalize=False, mins=None, maxs=None):
        """x [B,T+1,H,W,17] float32 -> [B,H-14,W-14] float32 (pb:conv2d/Sigmoid).  H and W are independent (multiples of 4,
        >= 28): square for the released graphs, 220 x 684 for the border re-segmentation pass.  `mins` / `maxs` override the
        session's 17 normalisation constants for this call (resegment_tiles_wide.py uses a different DEM maximum)."""
        x = np.ascontiguousarray(x, np.float32)
        B, T1, H, W, Cc = x.shape
        assert Cc == 17
        length = int(self.length if length is None else length)
        out = np.empty((B, H - 14, W - 14), np.float32)
        mn, mnp = _f64(self.min_all if mins is None else mins)
        mx, mxp = _f64(self.max_all if maxs is None else maxs)
        self._check(self.lib.stc_predict_host(self.h, _dptr(x), B, T1 - 1, H, W, length, int(bool(normalize)), mnp, mxp, _dptr(out)))
        return out

    def predict_feats(self, x, length=None, normalize=False):
        """Forward plus the two --gen_feats taps: (probs [B,H-14,W-14], early [B,H-14,W-14,64], late [..,64])
        = pb:conv2d/Sigmoid, pb:gru_drop/drop_block2d/cond/Merge (centre-cropped), pb:csse_out_mul/mul."""
        x = np.ascontiguousarray(x, np.float32)
        B, T1, H, W, Cc = x.shape
        assert Cc == 17
        length = int(self.length if length is None else length)
        probs = np.empty((B, H - 14, W - 14), np.float32)
        early = np.empty((B, H - 14, W - 14, 64), np.float32)
        late = np.empty((B, H - 14, W - 14, 64), np.float32)
        mn, mnp = _f64(self.min_all)
        mx, mxp = _f64(self.max_all)
        self._check(self.lib.stc_predict_feats_host(self.h, _dptr(x), B, T1 - 1, H, W, length, int(bool(normalize)), mnp, mxp,
                                                    _dptr(probs), _dptr(early), _dptr(late)))
        return probs, early, late

    def float_to_int16(self, arr, precision=1000):
        a = np.ascontiguousarray(arr, np.float32)
        out = np.empty(a.shape, np.int16)
        self._check(self.lib.stc_float_to_int16_host(self.h, _dptr(a), a.size, int(precision), _dptr(out)))
        return out

    def mosaic_feats(self, feats, xs, ys, out_shape, sigma=36):
        """Feature mosaic (load_mosaic_predictions, depth > 1): feats list of [S,S,D] int16 stacks in the reference's
        layer order -> [D, out_shape[0], out_shape[1]] int16."""
        n = len(feats)
        F = np.ascontiguousarray(np.stack([np.asarray(f) for f in feats]), np.int16)
        S, D = F.shape[1], F.shape[3]
        xs = np.ascontiguousarray(xs, np.int32)
        ys = np.ascontiguousarray(ys, np.int32)
        gauss = np.ascontiguousarray(fspecial_gauss(S, sigma), np.float32)
        out = np.empty((D, int(out_shape[0]), int(out_shape[1])), np.int16)
        self._check(self.lib.stc_feature_mosaic_host(self.h, _dptr(F), _dptr(xs), _dptr(ys), _dptr(gauss), n, S, D,
                                                     int(out_shape[0]), int(out_shape[1]), _dptr(out)))
        return out

    def predict_dev(self, x_dev, B, T, H, W, out_dev, length=None, normalize=False):
        length = int(self.length if length is None else length)
        mn, mnp = _f64(self.min_all)
        mx, mxp = _f64(self.max_all)
        self._check(self.lib.stc_predict_dev(self.h, x_dev, B, T, H, W, length, int(bool(normalize)), mnp, mxp, out_dev))

    def assemble(self, monthly):
        """monthly [B,12,H,W,13] -> [B,5,H,W,17] (see include/stc.h)."""
        m = np.ascontiguousarray(monthly, np.float32)
        B, n, H, W, Cc = m.shape
        assert n == 12 and Cc == 13
        out = np.empty((B, 5, H, W, 17), np.float32)
        self._check(self.lib.stc_assemble_host(self.h, _dptr(m), B, H, W, _dptr(out)))
        return out

    def predict_patches(self, monthly, out=None):
        """Fused tile path: monthly [B,12,H,W,13] -> tree-cover probabilities [B,H-14,W-14]."""
        u16 = monthly.dtype == np.uint16          # integer storage convention: x / 65535 (predict_subtile :345-347)
        want = np.uint16 if u16 else np.float32
        m = monthly if (monthly.dtype == want and monthly.flags.c_contiguous) else np.ascontiguousarray(monthly, want)
        B, n, H, W, Cc = m.shape
        assert n == 12 and Cc == 13
        if out is None:
            out = np.empty((B, H - 14, W - 14), np.float32)
        mn, mnp = _f64(self.min_all)
        mx, mxp = _f64(self.max_all)
        fn = self.lib.stc_predict_patches_u16_host if u16 else self.lib.stc_predict_patches_host
        self._check(fn(self.h, _dptr(m), B, H, W, mnp, mxp, _dptr(out)))
        return out

    def predict_patches_dev(self, m_dev, B, H, W, out_dev):
        mn, mnp = _f64(self.min_all)
        mx, mxp = _f64(self.max_all)
        self._check(self.lib.stc_predict_patches_dev(self.h, m_dev, B, H, W, mnp, mxp, out_dev))

    def debug_read(self, name):
        n = self.lib.stc_debug_read(self.h, name.encode(), None)
        self._check(n)
        out = np.empty(n, np.float32)
        self._check(self.lib.stc_debug_read(self.h, name.encode(), _dptr(out)))
        return out

    # -- preprocessing -------------------------------------------------------------
    def temporal_matmul(self, arr, M):
        """out[o] = sum_n M[o,n] * arr[n]  over the leading axis."""
        a = np.ascontiguousarray(arr, np.float32)
        M = np.ascontiguousarray(M, np.float32)
        n_out, n_in = M.shape
        assert a.shape[0] == n_in
        inner = int(np.prod(a.shape[1:]))
        out = np.empty((n_out,) + a.shape[1:], np.float32)
        self._check(self.lib.stc_temporal_matmul_host(self.h, _dptr(a), M.ctypes.data_as(_f32p), n_in, n_out, inner, _dptr(out)))
        return out

    def indices(self, arr):
        a = np.ascontiguousarray(arr, np.float32)
        npix = int(np.prod(a.shape[:-1]))
        out = np.empty(a.shape[:-1] + (4,), np.float32)
        self._check(self.lib.stc_indices_host(self.h, _dptr(a), npix, a.shape[-1], _dptr(out)))
        return out

    def temporal_median(self, arr):
        a = np.ascontiguousarray(arr, np.float32)
        inner = int(np.prod(a.shape[1:]))
        out = np.empty(a.shape[1:], np.float32)
        self._check(self.lib.stc_temporal_median_host(self.h, _dptr(a), a.shape[0], inner, _dptr(out)))
        return out

    def superresolve(self, x10, bilinear6=None):
        """pb:superresolve graph on [N,H,W,10] float32 (+ the bilinear input [N,H,W,6]; None = x10[..., 4:], which is
        what superresolve_large_tile feeds) -> the six resolved bands [N,H,W,6]."""
        x = np.ascontiguousarray(x10, np.float32)
        N, H, W, _ = x.shape
        out = np.empty((N, H, W, 6), np.float32)
        if bilinear6 is None:
            self._check(self.lib.stc_superresolve_host(self.h, _dptr(x), None, N, H, W, _dptr(out)))
        else:
            b = np.ascontiguousarray(bilinear6, np.float32)
            self._check(self.lib.stc_superresolve_host(self.h, _dptr(x), _dptr(b), N, H, W, _dptr(out)))
        return out


    def order_stats(self, data, ks):
        """Exact order statistics of the columns of `data` [rows, cols<=16] float32 (any row stride): returns [cols, 2] =
        the values of 0-based rank ks[c] and ks[c] + 1 in sorted order (NaN last).  np.median / np.percentile are one
        interpolation away from these (stc_order_stats_host)."""
        a = np.asarray(data, np.float32)
        if a.ndim == 1:
            a = a[:, None]
        if a.strides[1] != 4 or a.strides[0] % 4 or a.strides[0] < 4 * a.shape[1]:
            a = np.ascontiguousarray(a)
        rows, cols = a.shape
        k = np.ascontiguousarray(np.broadcast_to(np.asarray(ks, np.int32), (cols,)))
        out = np.empty((cols, 2), np.float32)
        self._check(self.lib.stc_order_stats_host(self.h, C.c_void_p(a.ctypes.data), rows, cols, a.strides[0] // 4, _dptr(k), _dptr(out)))
        return out

    def to_float32(self, arr_u16):
        a = np.ascontiguousarray(arr_u16, np.uint16)
        out = np.empty(a.shape, np.float32)
        self._check(self.lib.stc_to_float32_host(self.h, _dptr(a), a.size, _dptr(out)))
        return out

    def to_uint16(self, arr):
        a = np.ascontiguousarray(arr, np.float32)
        out = np.empty(a.shape, np.uint16)
        self._check(self.lib.stc_to_uint16_host(self.h, _dptr(a), a.size, _dptr(out)))
        return out

    def convert_to_db(self, arr, min_db):
        a = np.ascontiguousarray(arr, np.float32)
        out = np.empty(a.shape, np.float32)
        self._check(self.lib.stc_convert_to_db_host(self.h, _dptr(a), a.size, float(min_db), _dptr(out)))
        return out

    def feather(self, masks, closing_size):
        """[n,H,W] float32 0/1 masks -> feathered interpolation weights (see include/stc.h)."""
        m = np.ascontiguousarray(masks, np.float32)
        n, H, W = m.shape
        out = np.empty_like(m)
        self._check(self.lib.stc_feather_host(self.h, _dptr(m), n, H, W, int(closing_size), _dptr(out)))
        return out

    def binary_dilation(self, x, iterations=1, connectivity=1):
        """scipy.ndimage.binary_dilation(x, structure=cross|3x3, iterations=k) on [H,W] or [n,H,W]."""
        a = np.ascontiguousarray(np.asarray(x) != 0, np.uint8)
        shp = a.shape
        a3 = a.reshape((-1,) + shp[-2:])
        out = np.empty_like(a3)
        self._check(self.lib.stc_binary_dilate_host(self.h, _dptr(a3), a3.shape[0], shp[-2], shp[-1], int(iterations),
                                                    int(connectivity), _dptr(out)))
        return out.reshape(shp).astype(bool)

    def edt_capped(self, target, cap):
        """min(distance_transform_edt(1 - target), cap) in float64 (exact integer search on the GPU,
        square root on the host).  target [H,W] or [n,H,W], non-zero = distance-zero pixels."""
        a = np.ascontiguousarray(np.asarray(target) != 0, np.uint8)
        shp = a.shape
        a3 = a.reshape((-1,) + shp[-2:])
        d2 = np.empty(a3.shape, np.int32)
        self._check(self.lib.stc_edt_sq_host(self.h, _dptr(a3), a3.shape[0], shp[-2], shp[-1], int(np.ceil(cap)), _dptr(d2)))
        return np.minimum(np.sqrt(d2.astype(np.float64)), float(cap)).reshape(shp)

    def missing_px_counts(self, arr):
        """Per date of arr [n,H,W,C]: (#pixels with >1 of the first 10 bands == 0 or >= 1, #NaN values)."""
        a = np.ascontiguousarray(arr, np.float32)
        n, H, W, Cc = a.shape
        bad = np.zeros(n, np.int32)
        nans = np.zeros(n, np.int32)
        self._check(self.lib.stc_missing_px_host(self.h, _dptr(a), n, H, W, Cc, _dptr(bad), _dptr(nans)))
        return bad, nans

    def median_fill(self, arr):
        """In-place 0 / 1 sentinel fill with the running temporal median (deal_w_missing_px :1039-1047).
        arr must be a C-contiguous float32 [n,H,W,C] array; returns the per-date NaN counts afterwards."""
        if not (isinstance(arr, np.ndarray) and arr.dtype == np.float32 and arr.flags.c_contiguous and arr.ndim == 4):
            raise ValueError("median_fill needs a C-contiguous float32 [n,H,W,C] array (filled in place)")
        n, H, W, Cc = arr.shape
        nans = np.zeros(n, np.int32)
        self._check(self.lib.stc_median_fill_host(self.h, _dptr(arr), n, H, W, Cc, _dptr(nans)))
        return nans

    def build_sentinel2(self, s2_10, s2_20):
        """process_tile :743-782: (n,2h,2w,4) 10 m bands + (n,h,w,6) 20/40 m bands -> (n,2h,2w,10) float32."""
        a = np.ascontiguousarray(s2_10, np.float32)
        b = np.ascontiguousarray(s2_20, np.float32)
        n, h, w, c6 = b.shape
        if c6 != 6 or a.shape != (n, 2 * h, 2 * w, 4):
            raise ValueError("shapes %r / %r are not (n,2h,2w,4) / (n,h,w,6)" % (a.shape, b.shape))
        out = np.empty((n, 2 * h, 2 * w, 10), np.float32)
        self._check(self.lib.stc_build_sentinel2_host(self.h, _dptr(a), _dptr(b), n, h, w, _dptr(out)))
        return out

    CLOUD_STAGES = {"clm": 1, "shadows_raw": 2, "shadows_clean": 3, "clouds_raw": 4, "clouds_bright": 5,
                    "clouds_fp": 6, "clouds_shape": 7, "clouds_pre_haze": 8, "fcps0": 9}

    def set_ancillary_masks(self, forest=None, urban=None, shape=None):
        """ESA WorldCover rasters of the tile whose masks are computed next (stc_set_ancillary_masks_host): `forest` [H,W]
        0/1 = what adjust_cloudmask_in_forests returns (cloud_removal.py:758-771), `urban` = (core, near) [H,W] 0/1 = the two
        resized rasters of mask_nonurban_areas (:735-755); see ancillary_masks_from_rasters.  None clears a mask (the
        reference's behaviour when the .tif is missing).  They apply to cloud_masks / run_tile calls of the same H, W."""
        f = None if forest is None else np.ascontiguousarray(np.asarray(forest) != 0, np.uint8)
        c = n = None
        if urban is not None:
            c = np.ascontiguousarray(np.asarray(urban[0]) != 0, np.uint8)
            n = np.ascontiguousarray(np.asarray(urban[1]) != 0, np.uint8)
        shp = shape or (f.shape if f is not None else (c.shape if c is not None else (0, 0)))
        for m in (f, c, n):
            if m is not None and m.shape != tuple(shp):
                raise ValueError("ancillary mask shape %r is not %r" % (m.shape, tuple(shp)))
        self._check(self.lib.stc_set_ancillary_masks_host(self.h, _dptr(f) if f is not None else None, _dptr(c) if c is not None else None,
                                                          _dptr(n) if n is not None else None, int(shp[0]), int(shp[1])))

    def cloud_masks(self, img, dem, stage=None):
        """identify_clouds_shadows (cloud_removal.py:1215-1677) with the ancillary rasters last given to set_ancillary_masks.
        img [T,H,W,>=10] float32 reflectance, dem [H,W] -> (clouds float32 [T,H,W], fcps bool [T,H,W]);
        with `stage` (a CLOUD_STAGES name) also returns that intermediate uint8 mask (test tap)."""
        a = np.ascontiguousarray(np.asarray(img)[..., :10], np.float32)
        T, H, W, _ = a.shape
        d = np.ascontiguousarray(dem, np.float32)
        if d.shape != (H, W):
            raise ValueError("dem shape %r does not match image %r" % (d.shape, (H, W)))
        clouds = np.empty((T, H, W), np.float32)
        fcps = np.empty((T, H, W), np.uint8)
        tap = np.empty((T, H, W), np.uint8) if stage else None
        self._check(self.lib.stc_cloud_masks_host(self.h, _dptr(a), _dptr(d), T, H, W, _dptr(clouds), _dptr(fcps),
                                                  _dptr(tap) if stage else None, self.CLOUD_STAGES[stage] if stage else 0))
        if stage:
            return clouds, fcps.astype(bool), tap
        return clouds, fcps.astype(bool)

    def np_sum(self, data, mode=0):
        """np.sum of every row of a contiguous float32 [nseg, len] array in NumPy's pairwise order.
        mode 1: values < 255 scaled by 100 first; mode 2: nan-sum.  Returns (sums float32, valid counts)."""
        a = np.ascontiguousarray(data, np.float32)
        nseg, ln = a.shape
        sums = np.empty(nseg, np.float32)
        valid = np.empty(nseg, np.int32)
        self._check(self.lib.stc_np_sum_host(self.h, _dptr(a), nseg, ln, int(mode), _dptr(sums), _dptr(valid)))
        return sums, valid

    def normalize(self, x, mins, maxs):
        """normalize_subtile arithmetic in place on a C-contiguous float32 [..., C] array."""
        if not (isinstance(x, np.ndarray) and x.dtype == np.float32 and x.flags.c_contiguous):
            raise ValueError("normalize needs a C-contiguous float32 array (normalised in place)")
        Cc = x.shape[-1]
        mn = np.ascontiguousarray(mins, np.float64)
        mx = np.ascontiguousarray(maxs, np.float64)
        if mn.shape != (Cc,) or mx.shape != (Cc,):
            raise ValueError("need %d minima / maxima" % Cc)
        self._check(self.lib.stc_normalize_host(self.h, _dptr(x), x.size // Cc, Cc, _dptr(mn), _dptr(mx)))
        return x

    def bright_bare(self, img):
        """identify_bright_bare_surfaces: img [F,H,W,C] -> float64 ramp [(H-14),(W-14)]."""
        a = np.ascontiguousarray(img, np.float32)
        F, H, W, Cc = a.shape
        out = np.empty((H - 14, W - 14), np.float64)
        self._check(self.lib.stc_bright_bare_host(self.h, _dptr(a), F, H, W, Cc, _dptr(out)))
        return out

    def postprocess_subtile(self, preds, subtile_all, min_clear):
        """Post-filters of the subtile loop (:1408-1409,1451-1483) -> float32 [S,S]."""
        p = np.ascontiguousarray(preds, np.float32)
        a = np.ascontiguousarray(subtile_all, np.float32)
        m = np.ascontiguousarray(min_clear, np.float32)
        S = p.shape[0]
        F, H, W, Cc = a.shape
        if p.shape != (S, S) or (H, W) != (S + 14, S + 14) or m.shape != (S + 14, S + 14):
            raise ValueError("shapes %r / %r / %r are not (S,S) / (F,S+14,S+14,C) / (S+14,S+14)" % (p.shape, a.shape, m.shape))
        out = np.empty((S, S), np.float32)
        self._check(self.lib.stc_postprocess_subtile_host(self.h, _dptr(p), _dptr(a), _dptr(m), S, F, Cc, _dptr(out)))
        return out

    def remove_clouds(self, tiles, probs, pfcps, mt_state, want_mosaic=False, clip_when_all_kept=False):
        """remove_cloud_and_shadows core (cloud_removal.py:888-973).  tiles: C-contiguous float32 [n,H,W,10],
        rewritten in place.  mt_state: uint32[625] (Python `random.getstate()[1]`), advanced in place.
        Returns (areas [n,H,W] float32, to_remove list[, mosaic [H,W,10]])."""
        if not (isinstance(tiles, np.ndarray) and tiles.dtype == np.float32 and tiles.flags.c_contiguous and tiles.ndim == 4
                and tiles.shape[-1] == 10):
            raise ValueError("remove_clouds needs a C-contiguous float32 [n,H,W,10] array (blended in place)")
        n, H, W, _ = tiles.shape
        p = np.ascontiguousarray(probs, np.float32)
        f = np.ascontiguousarray(np.asarray(pfcps) != 0, np.uint8).reshape(-1, H, W)
        if p.shape != (n, H, W) or f.shape[0] not in (1, n):
            raise ValueError("mask shapes %r / %r do not match tiles %r" % (p.shape, f.shape, tiles.shape))
        if f.shape[0] != n:
            f = np.ascontiguousarray(np.repeat(f, n, 0))
        if not (isinstance(mt_state, np.ndarray) and mt_state.dtype == np.uint32 and mt_state.shape == (625,)):
            raise ValueError("mt_state must be a uint32[625] array")
        areas = np.empty((n, H, W), np.float32)
        rem = np.zeros(n, np.int32)
        mosaic = np.empty((H, W, 10), np.float32) if want_mosaic else None
        if clip_when_all_kept:               # process_tile: fold the np.clip(sentinel2, 0, 1) that follows into the download
            clipped = C.c_int32(0)
            self._check(self.lib.stc_remove_clouds_clip_host(self.h, _dptr(tiles), _dptr(p), _dptr(f), n, H, W, _dptr(mt_state),
                                                             _dptr(areas), _dptr(rem), C.byref(clipped)))
            return areas, [int(i) for i in np.flatnonzero(rem)], bool(clipped.value)
        self._check(self.lib.stc_remove_clouds_host(self.h, _dptr(tiles), _dptr(p), _dptr(f), n, H, W, _dptr(mt_state), _dptr(areas),
                                                    _dptr(rem), _dptr(mosaic) if want_mosaic else None))
        out = (areas, [int(i) for i in np.flatnonzero(rem)])
        return out + (mosaic,) if want_mosaic else out

    def run_tile(self, s2_10, s2_20, s1, dem, dates, clm=None, make_shadow=True, superresolve=True, size=158, length=4,
                 return_subtiles=False):
        """One whole tile on the device (stc_tile_run_host): raw uint16 cubes -> uint8 tree-cover tile, the body of the
        reference's main loop (src/download_and_predict_job.py:1995-2020).  s2_10 [n,h10,w10,4], s2_20 [n,h20,w20,6],
        s1 [12,hs,ws,2] uint16 as stored under raw/; dem float32; dates [n] int; clm optional uint8 [n,h20,w20].
        Python's global `random` state is handed to the library and advanced, as in remove_cloud_and_shadows.
        Returns (tile uint8 [2*w20, 2*h20], dates_kept[, subtile predictions [36,size,size]])."""
        import random
        a10 = np.ascontiguousarray(s2_10, np.uint16); a20 = np.ascontiguousarray(s2_20, np.uint16)
        a1 = np.ascontiguousarray(s1, np.uint16); d = np.ascontiguousarray(dem, np.float32)
        dt = np.ascontiguousarray(dates, np.int32)
        n, h10, w10, c4 = a10.shape
        n2, h20, w20, c6 = a20.shape
        m1, hs, ws, c2 = a1.shape
        if c4 != 4 or c6 != 6 or c2 != 2 or n2 != n or dt.shape != (n,) or d.ndim != 2:
            raise ValueError("run_tile: shapes %r / %r / %r / %r / %r" % (a10.shape, a20.shape, a1.shape, d.shape, dt.shape))
        cm = None
        if clm is not None:
            cm = np.ascontiguousarray(clm, np.uint8)
            if cm.shape != (n, h20, w20):
                raise ValueError("run_tile: Sen2Cor mask %r is not (n, h20, w20)" % (cm.shape,))
        version, internal, gaussian = random.getstate()
        state = np.array(internal, dtype=np.uint32)
        out = np.empty((2 * w20, 2 * h20), np.uint8)
        kept = np.zeros(n, np.int32)
        nk = C.c_int32(0)
        gauss = np.ascontiguousarray(fspecial_gauss(size, 36), np.float32)
        nt = 36 if size != 222 else 49
        sub = np.empty((nt, size, size), np.float32) if return_subtiles else None
        mn, mnp = _f64(self.min_all)
        mx, mxp = _f64(self.max_all)
        rc = self.lib.stc_tile_run_host(self.h, _dptr(a10), n, h10, w10, _dptr(a20), h20, w20, _dptr(a1), m1, hs, ws, _dptr(d), d.shape[0],
                                        d.shape[1], _dptr(cm) if cm is not None else None, _dptr(dt), _dptr(state), int(bool(make_shadow)),
                                        int(bool(superresolve)), int(size), int(length), mnp, mxp, _dptr(gauss), _dptr(out), out.shape[0],
                                        out.shape[1], _dptr(kept), C.byref(nk), _dptr(sub) if return_subtiles else None)
        random.setstate((version, tuple(int(v) for v in state), gaussian))      # the generator advanced even if a later stage failed
        self._check(rc)
        kept = kept[:nk.value].copy()
        return (out, kept, sub) if return_subtiles else (out, kept)

    def pool_info(self):
        v = [C.c_int64() for _ in range(4)]
        self._check(self.lib.stc_pool_info(self.h, *[C.byref(x) for x in v]))
        return dict(zip(("hits", "misses", "cached_bytes", "total_bytes"), [x.value for x in v]))

    def mosaic(self, preds, xs, ys, out_shape, sigma=36):
        """Gaussian overlap blend of subtile predictions (list/array [n,S,S], the arrays as
        saved by process_subtiles) placed at (xs[i], ys[i]) -> uint8 canvas `out_shape`."""
        n = len(preds)
        S = np.asarray(preds[0]).shape[0]
        P = np.empty((n, S, S), np.float32)
        placed = np.zeros(n, np.int32)
        for i, a in enumerate(preds):
            P[i] = np.asarray(a, np.float32)
        # :1570-1573: values < 255 are scaled by 100, an all-no-data subtile (sum == S*S*255) is skipped
        sums, _ = self.np_sum(P.reshape(n, S * S), mode=1)
        placed[:] = sums < S * S * 255
        xs = np.ascontiguousarray(xs, np.int32)
        ys = np.ascontiguousarray(ys, np.int32)
        gauss = np.ascontiguousarray(fspecial_gauss(S, sigma), np.float32)
        mult = np.ones(n, np.float32)
        if placed.all():                                     # an unplaced subtile makes calc_overlap raise -> no reweighting (:1597-1608)
            diffs = np.empty((n, S, S), np.float32)
            self._check(self.lib.stc_mosaic_diffs_host(self.h, _dptr(P), _dptr(xs), _dptr(ys), _dptr(placed), n, S, _dptr(diffs)))
            # np.nanmean of every difference map (:1512): pairwise float32 sum of the non-NaN values on the GPU,
            # float32(sum / count) like NumPy's scalar path; what is left here are n scalars
            sums, valid = self.np_sum(diffs.reshape(n, S * S), mode=2)
            with np.errstate(all="ignore"):
                ratios = (sums.astype(np.float64) / valid).astype(np.float32)
                mult = (np.median(ratios) / ratios).astype(np.float32)
                mult[mult > 1.5] = 1.5
        out = np.empty(out_shape, np.uint8)
        self._check(self.lib.stc_gauss_mosaic_host(self.h, _dptr(P), _dptr(xs), _dptr(ys), _dptr(placed), _dptr(gauss),
                                                   _dptr(np.ascontiguousarray(mult, np.float32)), n, S,
                                                   int(out_shape[0]), int(out_shape[1]), _dptr(out)))
        return out


# ======================================================================================
# reference-signature functions
# ======================================================================================
def normalize_subtile(subtile, min_all=MIN_ALL, max_all=MAX_ALL, sess=None):
    """src/download_and_predict_job.py:316-325 -- in place, returns the same array.  Runs on the GPU
    (stc_normalize_host); the batched predict path (`StcSession.predict(..., normalize=True)`,
    `predict_patches`) fuses the same arithmetic into the model's input packing instead."""
    if sess is None:
        raise RuntimeError("normalize_subtile needs an StcSession (sess=...); there is no CPU path")
    if isinstance(subtile, np.ndarray) and subtile.dtype == np.float32 and subtile.flags.c_contiguous:
        return sess.normalize(subtile, min_all, max_all)
    work = np.ascontiguousarray(subtile, np.float32)
    sess.normalize(work, min_all, max_all)
    subtile[...] = work
    return subtile


# Tensor names a caller of the reference passes as `op` (src/download_and_predict_job.py:1807-1809)
PREDICT_LOGITS = "predict/conv2d_13/Sigmoid:0"
PREDICT_LATEFEATS = "predict/csse_out_mul/mul:0"
PREDICT_EARLYFEATS = "predict/gru_drop/drop_block2d/cond/Merge:0"


def predict_subtile(subtile, sess, op=None, size=None):
    """src/download_and_predict_job.py:328-369.  `op` selects the output like the reference's tensor handle:
    None / PREDICT_LOGITS -> probabilities, PREDICT_LATEFEATS / PREDICT_EARLYFEATS -> the 64-channel feature taps
    of the --gen_feats path (:1430-1431), squeezed and centre-cropped to `size` the same way.  All-zero input ->
    int 255 fill."""
    SIZE = subtile.shape[1] - 14
    size = SIZE if size is None else size
    if op not in (None, PREDICT_LOGITS, PREDICT_LATEFEATS, PREDICT_EARLYFEATS):
        raise ValueError("predict_subtile: unknown output tensor %r" % (op,))
    if np.sum(subtile) != 0:
        if not isinstance(subtile.flat[0], np.floating):
            assert np.max(subtile) > 1
            # `subtile / 65535.` (float64) followed by astype(float32) == correctly rounded float32 x/65535
            subtile = sess.to_float32(np.ascontiguousarray(subtile).astype(np.uint16, copy=False))
        batch_x = subtile[np.newaxis].astype(np.float32)
        if op in (PREDICT_LATEFEATS, PREDICT_EARLYFEATS):
            _, early, late = sess.predict_feats(batch_x, length=sess.length)
            preds = (late if op == PREDICT_LATEFEATS else early).squeeze()    # early is already cropped to H-14
        else:
            preds = sess.predict(batch_x, length=sess.length).squeeze()
        clip = (preds.shape[0] - size) // 2
        if clip > 0:
            preds = preds[clip:-clip, clip:-clip]
        preds = np.float32(preds)
    else:
        preds = np.full((SIZE, SIZE), 255)
    return preds


def predict_subtile_monthly(subtile, sess):
    """Legacy 12-step contract, src/download_and_predict_job_multiyear.py:794-838 (the "12-step x 13-band" signature of
    BASELINE.json): subtile (13, S+14, S+14, 13) = 12 monthly frames + the median frame, 13 bands (10 S2, DEM, 2 S1).
    Indices (:812-816), clip / normalise with the 17 constants (:819-820), sess.run with lengths = 12 (:826), squeeze,
    preds[1:-1, 1:-1] (:831).  All-zero (or negative-sum) input -> int 255 fill of the UNcropped size, like the
    reference (:834-835 uses SIZE).  The ConvGRU weights do not depend on the sequence length, so the same
    session serves the quarterly (length 4) and this monthly (length 12) graph."""
    SIZE = subtile.shape[1] - 14
    if np.sum(subtile) > 0:
        if not isinstance(subtile.flat[0], np.floating):
            assert np.max(subtile) > 1
            subtile = sess.to_float32(np.ascontiguousarray(subtile).astype(np.uint16, copy=False))
        sub = np.ascontiguousarray(subtile, np.float32)
        x = np.empty(sub.shape[:3] + (17,), np.float32)
        x[..., :13] = sub
        x[..., 13:] = sess.indices(sub)                       # evi, bi, msavi2, grndvi on the device (bands 0,1,2,3,8)
        preds = sess.predict(x[np.newaxis], length=sub.shape[0] - 1, normalize=True).squeeze()
        preds = preds[1:-1, 1:-1]
    else:
        preds = np.full((SIZE, SIZE), 255)
    return preds


def float_to_int16(arr, sess, precision=1000):
    """src/download_and_predict_job.py:174-180 (the NaN replacement is in place there too)."""
    return sess.float_to_int16(arr, precision)


def make_indices(arr, sess):
    """src/download_and_predict_job.py:998-1006 (EVI, BI, MSAVI2, GRNDVI)."""
    return sess.indices(arr)


def id_missing_px(sentinel2, thresh, sess):
    """src/preprocessing/interpolation.py:5-23: dates with >= H^2/thresh pixels having more than one
    band (of the first 10) equal to 0 or >= 1.  The per-date pixel counts come from the GPU."""
    bad, _ = sess.missing_px_counts(sentinel2)
    return np.argwhere(bad >= (sentinel2.shape[1] ** 2) / thresh).flatten()


def deal_w_missing_px(arr, dates, interp, sess):
    """src/download_and_predict_job.py:1031-1054.  Date bookkeeping (np.delete) on the host, counting and
    the running-median fill on the GPU.  Like the reference, `arr` is filled IN PLACE when no date was
    dropped first (np.delete makes a copy otherwise)."""
    missing_px = id_missing_px(arr, 10, sess)
    if len(missing_px) > 0:
        dates = np.delete(dates, missing_px)
        arr = np.delete(arr, missing_px, 0)
        interp = np.delete(interp, missing_px, 0)
    if arr.dtype == np.float32 and arr.flags.c_contiguous:
        nans = sess.median_fill(arr)
    else:
        filled = np.ascontiguousarray(arr, np.float32)
        nans = sess.median_fill(filled)
        arr[...] = filled
    to_remove = np.argwhere(nans > 0).flatten()
    if len(to_remove) > 0:
        dates = np.delete(dates, to_remove)
        arr = np.delete(arr, to_remove, 0)
        interp = np.delete(interp, to_remove, 0)
    return arr, dates, interp


def build_sentinel2(s2_10, s2_20, sess):
    """The 20 m -> 10 m upsampling block of process_tile (src/download_and_predict_job.py:743-782)."""
    return sess.build_sentinel2(s2_10, s2_20)


def smooth_large_tile(arr, dates, interp, sess):
    """src/download_and_predict_job.py:1057-1096: median-fill missing px, indices,
    15-day regrid + Whittaker + monthly mean -> (12,H,W,14).  The date logic and the
    12 x n operator are built on the host (regrid.py); the arithmetic runs on the GPU."""
    arr, dates, interp = deal_w_missing_px(arr, dates, interp, sess)
    try:
        M, _ = _regrid.monthly_operator(dates)
    except Exception:
        M = None
    if M is None:
        out = np.zeros((12, arr.shape[1], arr.shape[2], 14 if arr.shape[-1] == 10 else arr.shape[-1]), np.float32)
        return out, [0, ], interp
    sm = sess.temporal_matmul(arr, M)
    if arr.shape[-1] == 10:
        idx = sess.temporal_matmul(sess.indices(arr), M)
        out = np.concatenate([sm, idx], axis=-1)
    else:
        out = sm
    return out, dates, interp


def superresolve_large_tile(arr, sess, wsize=110):
    """src/download_and_predict_job.py:95-147.  110-px windows, last row/col anchored to
    the edge.  Quirks kept on purpose: the bottom row of windows reads a PRE-resolution
    copy of the bottom band (and that copy is updated in place window by window); the
    right-most column is only processed in the bottom row (:133-143 never reach the other
    right-edge windows).  Writes bands 4: of `arr` in place and returns it.
    Windows that cannot see each other's output go through the network in ONE batched call: the interior
    windows tile the array without overlap and read `arr` before anything is written; the bottom row
    reads the copy, where only the last (edge-anchored) window overlaps its left neighbour and therefore
    runs second.  Results are written back in the reference's order (later windows overwrite earlier ones)."""
    from .windows import superres_windows
    xs, ys = superres_windows(arr.shape[1], wsize), superres_windows(arr.shape[2], wsize)
    bottom_band = np.copy(arr[:, xs[-1]:, ...])
    order = [(x, y) for x in xs for y in ys if not (y == ys[-1] and x != xs[-1])]
    n = arr.shape[0]

    def run(wins):
        srcs = [bottom_band[:, :, y:y + wsize, ...] if x == xs[-1] else arr[:, x:x + wsize, y:y + wsize, ...] for x, y in wins]
        padded = np.concatenate([np.pad(s, ((0, 0), (4, 4), (4, 4), (0, 0)), 'reflect') for s in srcs])
        resolved = sess.superresolve(padded)
        for k, s in enumerate(srcs):
            s[..., 4:] = resolved[k * n:(k + 1) * n, 4:-4, 4:-4, :]
        return srcs

    last_overlaps = len(ys) > 1 and ys[-1] < ys[-2] + wsize           # the edge-anchored bottom window reads its neighbour's output
    first = [w for w in order if not (last_overlaps and w == (xs[-1], ys[-1]))]
    done = dict(zip(first, run(first)))
    if last_overlaps:
        done[(xs[-1], ys[-1])] = run([(xs[-1], ys[-1])])[0]
    for x, y in order:                                               # the reference's write order
        arr[:, x:x + wsize, y:y + wsize, ...] = done[(x, y)]
    return arr


def to_float32(array, sess):
    """src/tof/tof_downloading.py:64-72: integer arrays -> x/65535 float32; float arrays pass through."""
    if not isinstance(array.flat[0], np.floating):
        assert np.max(array) > 1
        array = sess.to_float32(array)
    assert np.max(array) <= 1
    assert array.dtype == np.float32
    return array


def to_int16(array, sess):
    """src/tof/tof_downloading.py:51-61: float [0,1] -> uint16 trunc(x*65535)."""
    assert np.min(array) >= 0, np.min(array)
    assert np.max(array) <= 1
    return sess.to_uint16(array)


def convert_to_db(x, min_db, sess):
    """src/download_and_predict_job.py:74-89."""
    return sess.convert_to_db(x, min_db)


def process_sentinel_1_tile(sentinel1, dates, sess):
    """src/tof/tof_downloading.py:75-95: Sentinel-1 dates -> 24 steps -> median of consecutive pairs
    (= their mean) -> 12 monthly composites, as one 12 x n operator applied on the GPU."""
    M, _ = _regrid.s1_monthly_operator(dates)
    return sess.temporal_matmul(sentinel1, M)


def _nn_index(n_out, n_in):
    """Source index of an order-0 resize (skimage.transform.resize(x, shape, 0)): nearest sample to (i + .5) n_in / n_out - .5."""
    k = np.floor((np.arange(n_out) + 0.5) * (n_in / float(n_out)) - 0.5 + 0.5).astype(np.int64)
    return np.clip(k, 0, n_in - 1)


def _dilate_cross(a, k):
    """k iterations of the 4-connected binary dilation with border_value 0 (tiny host rasters only)."""
    a = np.asarray(a) != 0
    for _ in range(k):
        p = np.pad(a, 1)
        a = p[1:-1, 1:-1] | p[:-2, 1:-1] | p[2:, 1:-1] | p[1:-1, :-2] | p[1:-1, 2:]
    return a


def ancillary_masks_from_rasters(forest_rst, urban_rst, shape):
    """What the reference derives from the two raster WINDOWS it reads with rasterio (the read itself is the caller's I/O):
    forest = dilate 2 -> order-0 resize to the tile (adjust_cloudmask_in_forests, cloud_removal.py:758-771); urban core =
    dilate 1 -> resize, urban near = dilate 5 more -> resize (mask_nonurban_areas, :735-755).  A ~40 x 40 raster: host NumPy.
    Returns (forest, (core, near)) ready for StcSession.set_ancillary_masks; None in -> None out."""
    def rs(a):
        return np.ascontiguousarray(a[_nn_index(shape[0], a.shape[0])][:, _nn_index(shape[1], a.shape[1])], np.uint8)
    forest = rs(_dilate_cross(forest_rst, 2)) if forest_rst is not None else None
    urban = None
    if urban_rst is not None:
        r1 = _dilate_cross(urban_rst, 1)
        urban = (rs(r1), rs(_dilate_cross(r1, 5)))
    return forest, urban


def identify_clouds_shadows(img, dem, bbx, sess, forest_mask=None, urban_mask=None):
    """src/preprocessing/cloud_removal.py:1215-1677, same arguments and return value
    `(clouds float32 [T,H,W], fcps bool [T,H,W])`.  The reference uses `bbx` only to cut the windows of
    forestmask.tif / urbanmask.tif (:1131-1135, :1254-1257) and falls back to zeros when the files are absent (they are
    not shipped with the tree); here the caller passes the arrays instead: `forest_mask` [H,W] and `urban_mask` =
    (core, near) [H,W] as produced by ancillary_masks_from_rasters.  Both None = the reference's fallback.
    Runs entirely on the GPU (stc_cloud_masks_host)."""
    del bbx
    H, W = np.asarray(dem).shape
    sess.set_ancillary_masks(forest_mask, urban_mask, (H, W))
    try:
        return sess.cloud_masks(img, dem)
    finally:
        sess.set_ancillary_masks(None, None)


def identify_bright_bare_surfaces(img, sess):
    """src/download_and_predict_job.py:1099-1122: NIR/SWIR < 0.9, mean RGB > 0.2, EVI < 0.3 in more
    than one frame -> open by dilations -> ramp min(EDT,3)/3, cropped by 7 px (float64).
    One GPU call (stc_bright_bare_host)."""
    return sess.bright_bare(img)


def postprocess_subtile(preds, subtile_all, min_clear_images_per_date, sess, size=158):
    """Post-filters of the subtile loop, src/download_and_predict_job.py:1408-1409,1451-1483:
    no-image 40x40 block vote -> 255, bright-bare-surface attenuation, round to 3 decimals.
    `subtile_all` is the (5, size+14, size+14, 17) stack BEFORE normalize_subtile,
    `min_clear_images_per_date` the (size+14, size+14) map before its [6:-6] crop.
    One GPU call (stc_postprocess_subtile_host)."""
    if np.asarray(preds).shape != (size, size):
        raise ValueError("preds shape %r is not (%d, %d)" % (np.asarray(preds).shape, size, size))
    return sess.postprocess_subtile(preds, subtile_all, min_clear_images_per_date)


def id_areas_to_interp(tiles, probs, shadows, image_dates, pfcps, sess):
    """src/preprocessing/cloud_removal.py:774-798: feathered interpolation masks (closing 15)."""
    a = np.clip(np.copy(probs).astype(np.float32), 0, 1)
    return sess.feather(a, 15)


def remove_cloud_and_shadows(tiles, probs, shadows, image_dates, pfcps, sentinel1, mosaic=None, sess=None, clip_when_all_kept=False):
    """src/preprocessing/cloud_removal.py:888-973, same arguments and return value
    `(tiles, areas_interpolated, to_remove)`; `tiles` is blended IN PLACE like the reference.
    `shadows`, `image_dates` and `sentinel1` are accepted and unused (the reference body never reads
    them: the S1 stack was dropped from the fit, :339-343, :400-407).  The reference samples its
    regression pixels with Python's global `random`; this wrapper hands the generator state to the
    library and stores the advanced state back, so `random.seed(k)` before the call gives the
    reference's sample and leaves `random` where the reference would leave it.
    The reference's debug dumps (mosaic.npy, tiles.npy, ... into the CWD, :925-927,972) are not written."""
    import random
    if sess is None:
        raise RuntimeError("remove_cloud_and_shadows needs an StcSession (sess=...); there is no CPU path")
    if mosaic is not None:
        raise NotImplementedError("a precomputed mosaic is never passed by the reference's callers "
                                  "(src/download_and_predict_job.py:935-944) and is not supported")
    del shadows, image_dates, sentinel1
    if isinstance(tiles, np.ndarray) and tiles.dtype == np.float32 and tiles.flags.c_contiguous:
        work = tiles
    else:
        work = np.ascontiguousarray(tiles, np.float32)
    version, internal, gauss = random.getstate()
    state = np.array(internal, dtype=np.uint32)
    clipped = False
    if clip_when_all_kept:       # tile.process_tile: fold the np.clip that follows (:996) into the download; extra return value
        areas, to_remove, clipped = sess.remove_clouds(work, probs, pfcps, state, clip_when_all_kept=True)
    else:
        areas, to_remove = sess.remove_clouds(work, probs, pfcps, state)
    random.setstate((version, tuple(int(v) for v in state), gauss))
    if work is not tiles:
        tiles[...] = work
    if clip_when_all_kept:
        return tiles, areas, to_remove, clipped
    return tiles, areas, to_remove


def fspecial_gauss(size, sigma):
    """src/download_and_predict_job.py:1489-1501 (float64)."""
    x, y = np.mgrid[-size // 2 + 1:size // 2 + 1, -size // 2 + 1:size // 2 + 1]
    return np.exp(-((x ** 2 + y ** 2) / (2.0 * sigma ** 2)))


def load_mosaic_predictions(out_folder, depth, sess, size=None):
    """src/download_and_predict_job.py:1515-1641: walk <out_folder>/<x>/<y>.npy in the reference's os.listdir order
    and blend on the GPU.  depth == 1: probabilities -> the uint8 tile; depth > 1: int16 feature stacks
    [S,S,>=depth] -> int16 [depth, max_x, max_y] (:1587-1592,1628-1635)."""
    x_tiles = [int(x) for x in os.listdir(out_folder) if '.DS' not in x]
    preds, xs, ys = [], [], []
    for x_tile in x_tiles:
        y_tiles = [int(y[:-4]) for y in os.listdir(out_folder + str(x_tile) + "/") if '.DS' not in y]
        for y_tile in y_tiles:
            f = out_folder + str(x_tile) + "/" + str(y_tile) + ".npy"
            if os.path.exists(f):
                preds.append(np.load(f)); xs.append(x_tile); ys.append(y_tile)
    S = size or preds[0].shape[0]
    max_x = np.max(x_tiles) + S
    max_y = np.max(y_tiles) + S          # like the reference: from the last listed x folder (:1535-1538)
    if depth == 1:
        return sess.mosaic(preds, xs, ys, (max_x, max_y))
    return sess.mosaic_feats([p[..., :depth] for p in preds], xs, ys, (max_x, max_y))


def write_tif(arr, point, x, y, out_folder, suffix="_FINAL"):
    """src/downloading/io.py:229-263, same arguments and return value: `arr` (the uint8 tile of load_mosaic_predictions) is
    written TRANSPOSED as `<out_folder><x>X<y>Y<suffix>.tif` -- single band, uint8, LZW, EPSG:4326, geotransform from the
    bounding box point = [west, south, east, north].  Host-only (no session): the LZW / GeoTIFF writer is part of libstc
    (csrc/stc_geotiff.cpp), rasterio / GDAL are not needed."""
    file = out_folder + f"{str(x)}X{str(y)}Y{suffix}.tif"
    west, east = point[0], point[2]
    north, south = point[3], point[1]
    a = np.ascontiguousarray(np.asarray(arr).T.astype(np.uint8))
    rc = load_library().stc_write_geotiff_u8(os.fsencode(file), _dptr(a), a.shape[0], a.shape[1],
                                             float(west), float(south), float(east), float(north))
    if rc:
        raise RuntimeError("write_tif: %s (%d)" % ({-2: "bad argument", -3: "cannot write " + file, -4: "out of memory"}.get(rc, "error"), rc))
    return file


def read_tif(path, band=1, return_bounds=False):
    """`rasterio.open(path).read(band)` for the uint8 tile products (src/resegment_tiles_wide.py:750): [rows, cols] uint8 through
    the TIFF reader of libstc (strips / tiles, none / LZW / PackBits, predictor 1 / 2) -- no rasterio / GDAL.  With
    `return_bounds` also [west, south, east, north] from the GeoTIFF tags (NaN when absent)."""
    lib = load_library()
    buf, rows, cols = C.c_void_p(), C.c_int(), C.c_int()
    bounds = (C.c_double * 4)()
    rc = lib.stc_read_geotiff_u8(os.fsencode(path), int(band), C.byref(buf), C.byref(rows), C.byref(cols), bounds)
    if rc:
        raise RuntimeError("read_tif: %s (%d): %s" % ({-2: "bad argument", -3: "unreadable, unsupported or corrupt TIFF", -4: "out of memory"}.get(rc, "error"), rc, path))
    try:
        img = np.frombuffer(C.string_at(buf, rows.value * cols.value), np.uint8).reshape(rows.value, cols.value).copy()
    finally:
        lib.stc_geotiff_free(buf)
    return (img, [float(b) for b in bounds]) if return_bounds else img


def geotiff_bytes(arr, point):
    """The bytes write_tif would put on disk for `arr` (already in raster orientation [rows, cols]) -- for callers that
    upload the product instead of keeping a file."""
    a = np.ascontiguousarray(arr, np.uint8)
    buf, n = C.c_void_p(), C.c_int64()
    lib = load_library()
    rc = lib.stc_geotiff_encode_u8(_dptr(a), a.shape[0], a.shape[1], float(point[0]), float(point[1]), float(point[2]), float(point[3]),
                                   C.byref(buf), C.byref(n))
    if rc:
        raise RuntimeError("geotiff_bytes: error %d" % rc)
    try:
        return C.string_at(buf, n.value)
    finally:
        lib.stc_geotiff_free(buf)
