"""CPU oracle: restated local-map fusion.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/inference/geotiff_raster.py: add_weights_band_to_rasters (:103-118),
_weighted_average_of_rasters (:294-347) applied pairwise in file order as rasterio.merge (rasterio==1.2.6,
un-vendored: parity with the real package unpinned) does with integer pixel offsets, then the first four
bands of finalize_merged_raster (:273-291, without the hard / admissibility bands).
"""
import numpy as np


def weight_image(D: int) -> np.ndarray:
    c = (np.arange(-D // 2, D // 2, 1) + 0.5) / D
    xx, yy = np.meshgrid(c, c)
    r = np.sqrt(xx ** 2 + yy ** 2)
    w = 1.5 - r
    w[r > 0.5] = np.nan
    return w


def _merge_pair(old, new):
    """geotiff_raster.py:294-347 on [2C,h,w] arrays (C score bands then C weight bands), NaN = nodata."""
    old, new = old.copy(), new.copy()
    old_nd, new_nd = np.isnan(old), np.isnan(new)
    C = old.shape[0] // 2
    uw = np.zeros_like(old[:C])
    with np.errstate(invalid="ignore", divide="ignore"):
        for b in range(C):
            wi = C + b
            old[b] = old[b] * old[wi] * (1 - old_nd[b])
            new[b] = new[b] * new[wi] * (1 - new_nd[b])
            w1 = old[wi] * (1 - old_nd[b])
            w2 = new[wi] * (1 - new_nd[b])
            uw[b] = np.nansum(np.stack([w1, w2]), axis=0)
            uw[b][old_nd[b] & new_nd[b]] = np.nan
        old[old_nd] = np.nan
        new[new_nd] = np.nan
        out = np.nansum(np.stack([old, new]), axis=0)
        out[old_nd & new_nd] = np.nan
        out[:C] = out[:C] / uw
    return out


def fuse_sequential(rasters: np.ndarray, offsets: np.ndarray, H: int, W: int) -> np.ndarray:
    """rasters [P,3,D,D] float64, offsets [P,2] -> [4,H,W]: 3 averaged bands + one weight band."""
    P, C, D, _ = rasters.shape
    w = weight_image(D)
    dest = np.full((2 * C, H, W), np.nan)
    for p in range(P):
        r0, c0 = offsets[p]
        new = np.concatenate([rasters[p], np.stack([w] * C)], axis=0)
        dest[:, r0:r0 + D, c0:c0 + D] = _merge_pair(dest[:, r0:r0 + D, c0:c0 + D], new)
    return dest[:C + 1]
