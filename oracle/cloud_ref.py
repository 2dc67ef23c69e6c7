"""TEST INFRASTRUCTURE (oracle) -- NumPy/SciPy restatement of the multi-temporal cloud / shadow
mask, /root/reference/src/preprocessing/cloud_removal.py:1215-1677 (`identify_clouds_shadows`),
both for the configuration the reference tree runs in as shipped (`urbanmask.tif` / `forestmask.tif` absent: the forest
mask and the potential-false-positive (urban) masks are all zero, :1131-1135, :1254-1257) and with the two ancillary
rasters given as arrays (`forest`, `urban=(core, near)`: what adjust_cloudmask_in_forests :758-771 returns and the two
resized rasters of mask_nonurban_areas :735-755), including the Fmask-4 parallax test of detect_pfcp (:1109-1212).
Pinned against the reference function itself executed through oracle/refshim.py (tests/test_cloud_masks.py; with the
rasters: the two loader functions monkey-patched to return seeded arrays).  Tests only; never imported by the product.

The function returns the final (clouds, fcps) and, with `stages=True`, the intermediate
arrays the CUDA pipeline is checked against stage by stage.
"""
import numpy as np
from scipy.ndimage import binary_dilation as dil, distance_transform_edt as edt


def _erode(x, k):
    """1 - binary_dilation(x == 0, iterations=k): erosion that never eats in from the border."""
    return 1 - dil(x == 0, iterations=k)


def _winsum3(a):
    """cloud_removal.py:1244-1249 with windowsize 3: 3x3 box sum, np.pad 'reflect' (no edge repeat)."""
    p = np.pad(a, 1, mode="reflect")
    p[3:] -= p[:-3]
    p[:, 3:] -= p[:, :-3]
    return p.cumsum(0)[2:].cumsum(1)[:, 2:]


def shadow_window(t, T):
    lo, hi = max(0, t - 4), min(T, t + 3)
    if hi - lo == 3:
        if hi == T:
            lo = max(lo - 1, 0)
        if lo == 0:
            hi = min(hi + 1, T)
    return list(range(lo, hi))


def cloud_windows(t, T):
    """others, close (initial list) of cloud_removal.py:1343-1361."""
    lo, hi = max(0, t - 2), min(T, t + 3)
    if hi - lo == 3:
        if hi == T:
            lo = max(lo - 2, 0)
        if lo == 0:
            hi = min(hi + 2, T)
    others = list(range(lo, hi))
    close = [max(0, t - 1), min(T - 1, t + 1)]
    if close[1] - close[0] < 2:
        if close[0] == 0:
            close = [close[0] + 1, close[1] + 1]
        else:
            close = [close[0] - 1, close[1] - 1]
    if close[-1] >= T - 2 and T > 3:
        close = [close[0] - 1] + close
    return others, close


def nn_resize_index(n_out, n_in):
    """Source index of every output sample of an order-0 resize (scipy.ndimage.zoom(order=0, grid_mode=True,
    mode='nearest'), the stand-in for skimage.transform.resize(..., 0)): the same map in every column / row."""
    from scipy.ndimage import zoom
    return zoom(np.arange(n_in, dtype=np.float64), n_out / float(n_in), order=0, mode="nearest", grid_mode=True, prefilter=False).astype(np.int64)


def rasters_to_masks(forest_rst, urban_rst, shape):
    """What the reference derives from the two ESA WorldCover windows (cloud_removal.py:735-771): forest = dilate 2 ->
    resize; urban core = dilate 1 -> resize; urban near = dilate 5 more -> resize.  Returns uint8 (forest, core, near)."""
    def rs(a):
        return a[nn_resize_index(shape[0], a.shape[0])][:, nn_resize_index(shape[1], a.shape[1])]
    forest = rs(dil(forest_rst, iterations=2)).astype(np.uint8) if forest_rst is not None else None
    core = near = None
    if urban_rst is not None:
        r1 = dil(urban_rst, iterations=1)
        core = rs(r1).astype(np.uint8)
        near = rs(dil(r1, iterations=5)).astype(np.uint8)
    return forest, core, near


def detect_pfcp(arr, dem, urban):
    """cloud_removal.py:1109-1212 with the urban raster given as (core, near) masks (None: the except-branch, zeros)."""
    from scipy import ndimage, signal
    T, H, W, _ = arr.shape
    ndvi = (arr[..., 3] - arr[..., 2]) / (arr[..., 3] + arr[..., 2])
    ndbi = (arr[..., 8] - arr[..., 3]) / (arr[..., 8] + arr[..., 3])
    ndwi = np.median((arr[..., 1] - arr[..., 3]) / (arr[..., 1] + arr[..., 3]), axis=0)
    pfps = np.median(np.logical_and(ndbi > 0, ndbi > ndvi), axis=0)
    pfps = pfps * (ndwi < 0)
    if urban is None:
        pfps = np.zeros_like(dem)
    else:
        core, near = urban
        pfps[core == 1] = 1.
        pfps[near == 0] = 0.
    pfps[(dem / 90) > 0.10] = 0.
    pfps = np.tile(pfps[np.newaxis], (T, 1, 1))
    cdis = np.zeros((T, H, W), np.float32)
    H2, W2 = H + H % 2, W + W % 2
    ru, cu = nn_resize_index(H2, H), nn_resize_index(W2, W)
    rd, cd = nn_resize_index(H, H2), nn_resize_index(W, W2)
    mean_op = np.ones((7, 7)) / 49

    def var7(x):
        return signal.convolve2d(x ** 2, mean_op, mode="same", boundary="symm") - signal.convolve2d(x, mean_op, mode="same", boundary="symm") ** 2

    for t in range(T):
        def pool(b, blur):
            x = np.copy(arr[t, ..., b])
            if (H % 2 + W % 2) > 0:
                x = x[ru][:, cu]
            if blur:
                x = ndimage.gaussian_filter(x, sigma=0.5, truncate=3)
            return np.mean(x.reshape(H2 // 2, 2, W2 // 2, 2), axis=(1, 3))
        b8, b8a, b7 = pool(3, True), pool(7, False), pool(6, False)
        r8a, r8a7 = var7(b8 / b8a), var7(b7 / b8a)
        cdi = (r8a7 - r8a) / (r8a7 + r8a)
        pf = (cdi >= -0.4).repeat(2, axis=0).repeat(2, axis=1)[rd][:, cd]
        cdis[t] = pf * (ndvi[t] < 0.4)
    s2 = np.ones((3, 3), bool)
    for t in range(T):
        cdis[t] = dil(cdis[t], iterations=6, structure=s2)
        pfps[t] = dil(pfps[t], iterations=6, structure=s2)
    return pfps * cdis, pfps


def identify_clouds_shadows(img, dem, stages=False, forest=None, urban=None):
    img = np.asarray(img, np.float32)
    T, H, W, _ = img.shape
    st = {}
    with np.errstate(all="ignore"):
        ndwi = (img[..., 1] - img[..., 3]) / (img[..., 1] + img[..., 3])
        water = np.nanmedian(ndwi, axis=0)
        forest = np.zeros_like(dem) if forest is None else np.asarray(forest)
        # Hollstein "okay" cloud mask (:1230-1242)
        clm = (img[..., 7] > 0.166) * (img[..., 1] > 0.28) * (img[..., 5] / img[..., 8] < 4.292)
        for t in range(T):
            clm[t] = dil(_erode(clm[t], 2), iterations=10)
        st["water"], st["clm"] = water, clm.astype(np.uint8)

        # ---- shadows (:1265-1324) ----
        shadows = np.zeros((T, H, W), np.float32)
        b4 = img[..., [0, 1, 7, 8]]
        allref = np.copy(b4)
        allref[clm > 0] = np.nan
        allref = np.nanmedian(allref, axis=0)
        allref[np.isnan(allref)] = np.median(b4, axis=0)[np.isnan(allref)]
        for t in range(T):
            win = shadow_window(t, T)
            r = np.copy(b4)[win]
            r[clm[win] > 0] = np.nan
            rmax = np.nanmax(r, axis=0)
            rmed = np.nanmedian(r, axis=0)
            rmed[np.isnan(rmed)] = np.min(b4, axis=0)[np.isnan(rmed)]
            x = img[t]
            s = ((x[..., 8] - rmed[..., 3]) < -0.04) * ((x[..., 7] - rmed[..., 2]) < -0.04) * (x[..., 0] < 0.09) * \
                ((x[..., 0] - rmed[..., 0]) < -0.02) * (x[..., 7] < 0.17)
            d8a, d11 = (x[..., 7] - rmax[..., 2]) < -0.04, (x[..., 8] - rmax[..., 3]) < -0.04
            dark = d11 * d8a * (x[..., 0] < 0.03) * (x[..., 7] < 0.18)
            dark[water > 0] = 0
            s = np.maximum(s, dark)
            s[water > 0] = 0
            slope = d8a * d11 * (x[..., 0] < 0.07) * (x[..., 7] < 0.18)
            slope = slope * (np.sum(x[..., :3], axis=-1) < 0.28)
            slope[water > 0] = 0
            slope = slope * (dem >= 25)
            s = np.maximum(s, slope)
            wsh = ((x[..., 0] - allref[..., 0]) < -0.05) * ((x[..., 1] - allref[..., 1]) < -0.05) * (x[..., 7] < 0.03) * \
                  ((allref[..., 1] - x[..., 1]) > 0.02) * (water > 0)
            shadows[t] = s + wsh
        st["shadows_raw"] = shadows.copy()
        for t in range(T):
            s = dil(_erode(shadows[t], 2), iterations=3)
            d = edt(1 - s)
            shadows[t] = 1 - (d > 5)
        st["shadows_clean"] = shadows.copy()

        # ---- clouds (:1342-1447) ----
        clouds = np.zeros((T, H, W), np.float32)
        rgb = img[..., [0, 1, 2]]
        p25 = [np.percentile(img[..., b], 25, axis=0) for b in range(3)]
        for t in range(T):
            others, close = cloud_windows(t, T)
            ref = np.copy(rgb)
            if T > 2:
                ref[shadows > 0] = np.nan
                up = [np.nanmin(ref[others, ..., b], axis=0) for b in range(3)]
                nanrep = np.isnan(up[0])
                for b in range(3):
                    up[b][nanrep] = p25[b][nanrep]
                rc = np.nanmin(ref[close], axis=0).astype(np.float32)
                lo_i, hi_i = close[0], close[-1]
                for _ in range(10):
                    if np.sum(np.isnan(rc) > 0):
                        lo_i, hi_i = max(lo_i - 1, 0), min(hi_i + 1, T)
                        cl2 = [k for k in range(lo_i, hi_i) if k != t]
                        new = np.nanmin(ref[cl2], axis=0).astype(np.float32)
                        rc[np.isnan(rc)] = new[np.isnan(rc)]
                if np.sum(np.isnan(rc) > 0):
                    rc[np.isnan(rc)] = np.min(img[..., :3], axis=0)[np.isnan(rc)]
            else:
                rc = np.min(ref, axis=0).astype(np.float32)
                up = [rc[..., 0], rc[..., 1], rc[..., 2]]
            thr = np.minimum((rc[..., 0] / 0.02 / 100) + 0.005, 0.10)
            thr = np.maximum(thr, 0.05)
            thr[forest == 1] -= 0.02
            thr = np.maximum(thr, 0.04)
            x = img[t]
            ci = ((x[..., 0] - up[0]) > 0.08) * ((x[..., 1] - up[1]) > 0.08) * ((x[..., 2] - up[2]) > 0.07)
            mean_i, mean_c, mod = 0., 1., 0.
            while (mean_c - mean_i) > 0.075:
                cc = ((x[..., 0] - rc[..., 0]) > (thr + mod + 0.01)) * ((x[..., 1] - rc[..., 1]) > (thr + mod + 0.01)) * \
                     ((x[..., 2] - rc[..., 2]) > (thr + mod))
                mean_i, mean_c = np.mean(ci > 0), np.mean(cc > 0)
                mod += 0.0025
            cc = cc * (np.sum(x[..., :3], axis=-1) < 0.75)
            cc_nf = _erode(cc, 2)
            cc = cc.astype(cc_nf.dtype) if cc.dtype == bool else cc
            cc = np.where(forest == 0, cc_nf, cc)
            clouds[t] = np.maximum(ci, cc)
        st["clouds_raw"] = clouds.copy()

        # ---- brightness z-score clouds (:1458-1481) ----
        bm = np.sum(img[..., :3], axis=-1)
        bm[np.logical_or(clouds > 0, shadows > 0)] = np.nan
        medb = np.nanmedian(bm, axis=(1, 2))
        bc = np.zeros_like(clouds, dtype=np.float32)
        for t in range(T):
            ratio = np.sum(img[t, ..., :3], axis=-1) / medb[t]
            ratio[water > 0] = 1.
            if np.sum(clouds[t] < 0.90):
                sel = ratio[clouds[t] == 0]
                z = (ratio - np.nanmean(sel)) / np.nanstd(sel)
            else:
                z = (ratio - np.nanmean(ratio)) / np.nanstd(ratio)
            bc[t][z > 3.5] = 1.
            bc[t] *= (water < 0)
        multi = np.sum((bc - clouds) > 0, axis=0)
        for t in range(T):
            bc[t][multi > 1] = 0.
        clouds = np.maximum(clouds, bc)
        # whiteness false positives (:1484-1492)
        for t in range(T):
            mb = np.mean(img[t, ..., :3], axis=-1)
            vr = np.max(img[t, ..., :3], axis=-1) - np.min(img[t, ..., :3], axis=-1)
            fp = (mb < 0.4) * ((vr / mb) > 0.5)
            clouds[t] = clouds[t] * (1 - fp)
        st["clouds_bright"] = clouds.copy()

        # ---- false-positive removal (:1497-1551); fcps == 0 without an urban mask ----
        if urban is None:                      # pfps == 0 (:1133-1135), so fcps == 0 whatever the parallax test says
            fcps, pfcps = np.zeros((T, H, W)), np.zeros((T, H, W))
        else:
            fcps, pfcps = detect_pfcp(img, dem, urban)
        st["fcps0"], st["pfcps0"] = (fcps > 0).astype(np.uint8), (pfcps > 0).astype(np.uint8)
        for t in range(T):
            lo, hi = max(t - 1, 0), min(t + 2, T)
            bmin = np.min(img[lo:hi, ..., :3], axis=(0, 3))
            isnt = ((np.mean(img[t, ..., :3], axis=-1) - bmin) < 0.4)
            rm = np.logical_and(fcps[t] > 0, isnt)
            clouds[t][rm] = 0.
            shadows[t][rm] = 0.
        nsr = (img[..., 3] / (img[..., 8] + 0.01)) < 0.75
        nsr = dil(nsr, iterations=3)          # note: 3-D dilation over (T,H,W) with the 3-D cross (:1518)
        for t in range(T):
            lo, hi = max(t - 1, 0), min(t + 2, T)
            bmin = np.min(img[lo:hi, ..., :3], axis=(0, 3))
            isnt = ((np.mean(img[t, ..., :3], axis=-1) - bmin) < 0.4)
            nsr[t][water < 0] = 0.
            clouds[t][np.logical_and(nsr[t] > 0, isnt)] = 0.
        for t in range(T):
            fp = dil((water > 0) * (img[t, ..., 8] < 0.11), iterations=10)
            clouds[t][fp] = 0.
        for t in range(T):
            clouds[t][_winsum3(clouds[t]) < 5] = 0.
        for t in range(T):
            bt = dil(np.sum(img[t, ..., :3], axis=-1) < 0.21, iterations=3)
            bt = (bt * (1 - forest)).astype(np.uint8)
            clouds[t][bt] = 0.                 # uint8 fancy index: rows 0/1 of the date are zeroed (:1546-1551)
        st["clouds_fp"] = clouds.copy()

        # ---- shape clean-up (:1590-1612) ----
        for t in range(T):
            c = _erode(clouds[t], 1)
            pf = dil(pfcps[t], iterations=5)
            urban = _erode(c * pf, 3)
            nu = c * (1 - pf)
            ws = _winsum3(nu)
            large, small = np.copy(nu), np.copy(nu)
            large[ws < 6] = 0.
            small[ws >= 6] = 0.
            small = dil(small, iterations=1)
            large = dil(large, iterations=5)
            nu = np.maximum(large, small)
            d = edt(1 - nu)
            nu = 1 - (d > 3)
            clouds[t] = nu + urban
        st["clouds_shape"] = clouds.copy()
        # ---- shadow plausibility (:1617-1626) ----
        for t in range(T):
            ms, mc = np.mean(shadows[t]), np.mean(clouds[t])
            if ms > (mc + 0.3) and mc < 0.3:
                far = np.logical_or(dil(np.copy(clouds[t]), iterations=50), dem >= 30)
                shadows[t] = shadows[t] * far
            if np.mean(clouds[t]) < 0.05 and ((np.mean(shadows[t]) / np.mean(clouds[t])) > 3):
                far = np.logical_or(dil(np.copy(clouds[t]), iterations=50), dem >= 30)
                shadows[t] = shadows[t] * far
        clouds = np.maximum(clouds, shadows)
        fcps = dil(np.maximum(fcps, nsr), iterations=2)     # 3-D dilation again (:1630)
        # ---- dark-blue shadow recovery (:1638-1648) ----
        for t in range(T):
            if np.mean(clouds[t]) < 0.9:
                inv = 1 / img[t, ..., 0][clouds[t] == 0]
                refv = np.mean(inv) + 2 * np.std(inv)
                s = (1 / img[t, ..., 0] > refv) * (img[t, ..., 7] < 0.17)
                s = dil(_erode(s, 2), iterations=2)
                s[water > 0] = 0.
                clouds[t] = np.maximum(clouds[t], s)
        clouds[clouds > 1] = 1.
        st["clouds_pre_haze"] = clouds.copy()
        # ---- haze (:1652-1676) ----
        mbr = np.mean(img[..., :3], axis=-1)
        mean_b, std_b, std_w = [], [], []
        for t in range(T):
            if np.mean(clouds[t]) < 1:
                sel = clouds[t] == 0
                mean_b.append(np.mean(mbr[t][sel]))
                std_b.append(np.std(mbr[t][sel]))
                std_w.append(np.std(np.ptp(img[t, ..., :3][sel], axis=1)))
        hb = mean_b / np.median(mean_b)
        hs = std_b / np.median(std_b)
        hw = std_w / np.median(std_w)
        haze = np.logical_or((hb >= 1.5) * (hs <= 0.67) * (hw < 1), (hb >= 1.3) * (hs <= 0.5))
        for k in range(len(haze)):
            if haze[k]:
                clouds[k] = 1.
    if stages:
        return clouds, fcps, st
    return clouds, fcps


from sentinel_tree_cover_b200.synth import synth_cloudy_cube  # noqa: E402,F401  (seeded generator shared with bench.py)
