"""TEST INFRASTRUCTURE (oracle) -- import shim that lets the reference's OWN
NumPy/SciPy preprocessing functions run unmodified in this container.

Used only by tools/make_golden.py (to generate tests/golden/*.npz) and by
tests that are skipped when /root/reference is absent.  Nothing here travels
to the GPU box as a dependency of the product path.

The reference imports network / geo / TF packages at module import time
(src/download_and_predict_job.py:1-53) that are not installed here and are
never touched by the numeric functions we call.  We register inert stub
modules for them, give `bottleneck` NumPy equivalents, and provide
`skimage.transform.resize` via scipy.ndimage (the algorithm scikit-image >= 0.19
publishes: anti-aliasing Gaussian when an axis shrinks, then ndimage.zoom with
grid_mode; order 0 exact; order 1 is the one place the shim is not guaranteed
identical to scikit-image).
"""
import os
import sys
import types
import importlib

REF_ROOT = os.environ.get("STC_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "src", "preprocessing"))


class _Anything(types.ModuleType):
    """Module stub: any attribute is another permissive stub / callable."""

    def __getattr__(self, name):
        if name.startswith("__") and name not in ("__version__",):
            raise AttributeError(name)
        v = _Anything(self.__name__ + "." + name)
        setattr(self, name, v)
        return v

    def __call__(self, *a, **k):
        return _Anything(self.__name__ + "()")

    def __iter__(self):
        return iter(())

    def __mro_entries__(self, bases):
        return (object,)


def _stub(name, **attrs):
    parts = name.split(".")
    for i in range(1, len(parts) + 1):
        n = ".".join(parts[:i])
        if n not in sys.modules:
            m = _Anything(n)
            m.__path__ = []
            sys.modules[n] = m
            if i > 1:
                setattr(sys.modules[".".join(parts[:i - 1])], parts[i - 1], m)
    for k, v in attrs.items():
        setattr(sys.modules[name], k, v)
    return sys.modules[name]


_installed = False


def install():
    global _installed
    if _installed:
        return
    import numpy as np
    import pandas  # noqa: F401  (must be imported before stubs shadow anything)
    import scipy.ndimage as ndi

    for m in ["sentinelhub", "sentinelhub.config", "sentinelhub.geo_utils", "sentinelhub.api",
              "pyproj", "shapely", "shapely.geometry", "rasterio", "rasterio.transform",
              "reverse_geocoder", "pycountry", "pycountry_convert", "hickle", "boto3",
              "boto3.s3", "boto3.s3.transfer", "botocore", "botocore.errorfactory",
              "botocore.exceptions", "tqdm", "yaml", "osgeo", "geopandas", "matplotlib",
              "matplotlib.pyplot", "seaborn", "skimage", "skimage.transform", "skimage.measure",
              "skimage.exposure", "sklearn.cross_decomposition"]:
        if m in ("tqdm", "yaml", "matplotlib", "matplotlib.pyplot"):
            try:
                importlib.import_module(m)
                continue
            except Exception:
                pass
        try:
            if m.startswith("sklearn"):
                importlib.import_module(m)
                continue
        except Exception:
            pass
        _stub(m)
    tf = _stub("tensorflow")
    tf.__version__ = "1.15"
    _stub("tensorflow.compat")
    _stub("tensorflow.compat.v1")

    def resize(img, shape, order=1, anti_aliasing=None, **kw):
        # scikit-image >= 0.19 (skimage/transform/_warps.py: resize): Gaussian pre-filter of sigma (in/out - 1) / 2 on the axes
        # that shrink (anti_aliasing defaults to on for non-boolean input unless integer input with order 0), then ndimage.zoom
        # with grid_mode=True.  Every call the goldens made before this branch existed enlarges or keeps the size.
        img = np.asarray(img)
        factors = np.divide(img.shape, shape)
        if anti_aliasing is None:
            anti_aliasing = (img.dtype != bool and not (np.issubdtype(img.dtype, np.integer) and order == 0)
                             and bool(np.any(factors > 1)))
        # skimage's default mode 'reflect' is ndimage's 'mirror' (skimage.transform._warps._to_ndimage_mode)
        mode = "nearest" if order == 0 else "mirror"
        if anti_aliasing:
            img = ndi.gaussian_filter(img.astype(np.float64) if img.dtype.kind != "f" else img, np.maximum(0, (factors - 1) / 2),
                                      cval=0, mode="mirror")
        zoom = [s / float(i) for s, i in zip(shape, img.shape)]
        return ndi.zoom(img, zoom, order=order, mode=mode, grid_mode=True, prefilter=False)
    sys.modules["skimage.transform"].resize = resize

    try:
        import bottleneck  # noqa: F401
    except Exception:
        bn = _stub("bottleneck")
        bn.__version__ = "shim"
        bn.median = np.median
        for f in ("nanmedian", "nanmean", "nanstd", "nanmax", "nanmin", "nansum", "nanvar"):
            setattr(bn, f, getattr(np, f))

    for p in (REF_ROOT, os.path.join(REF_ROOT, "src")):
        if p not in sys.path:
            sys.path.insert(0, p)
    _installed = True


def ref(modname):
    """Import a reference module, e.g. ref('preprocessing.indices'),
    ref('downloading.utils'), ref('download_and_predict_job')."""
    install()
    return importlib.import_module(modname)
