"""Host-side pre/post-processing of the CLI scripts (everything around the GPU hot path).

Mirrors UnMicst1-5.py:807-825,848-862 / UnMicst.py:621-660.  Nothing is resampled on the host: the raw
integer samples go to the GPU together with a PreMap (img_as_float scale + optional rescale_intensity
stretch, evaluated in float64 in-kernel) and, for --scalingFactor != 1, the size to resize them to; the
library applies scikit-image's ``resize`` defaults (bilinear, half-pixel grid, mirror boundary, Gaussian
anti-aliasing on shrink) on the fly and resizes the uint8 pages back.  ``resample`` below restates the same
on scipy.ndimage for the one case that needs the whole resized image on the host (--outlier percentiles).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

from .engine import PreMap

_INT_SCALE = {np.dtype(np.uint8): 1.0 / 255, np.dtype(np.uint16): 1.0 / 65535}


def as_float_scale(dtype) -> float:
    """Multiplier skimage's img_as_float applies (1/255, 1/65535, 1 for floats)."""
    return _INT_SCALE.get(np.dtype(dtype), 1.0)


def coerce_raw(raw: np.ndarray) -> np.ndarray:
    """`if I.dtype == 'float32': I = np.uint16(I)` (UnMicst1-5.py:807-808)."""
    return raw.astype(np.uint16) if raw.dtype == np.float32 else raw


def scaled_shape(raw_shape, factor: float) -> Tuple[int, int]:
    return int(float(raw_shape[0] * float(factor))), int(float(raw_shape[1] * float(factor)))


def resample(img: np.ndarray, out_shape: Tuple[int, int]) -> np.ndarray:
    """skimage.transform.resize(img, out_shape) on a float image (order 1, mode 'reflect',
    anti_aliasing when shrinking, clip to the input range)."""
    from scipy import ndimage as ndi
    img = np.asarray(img, dtype=np.float64)
    oh, ow = out_shape
    if (oh, ow) == img.shape:
        return img.copy()
    lo, hi = float(img.min()), float(img.max())
    fy, fx = img.shape[0] / oh, img.shape[1] / ow
    if fy > 1 or fx > 1:
        img = ndi.gaussian_filter(img, (max(0.0, (fy - 1) / 2), max(0.0, (fx - 1) / 2)), mode="mirror")
    out = ndi.zoom(img, (oh / img.shape[0], ow / img.shape[1]), order=1, mode="mirror", grid_mode=True)
    return np.clip(out, lo, hi)


def network_input(raw: np.ndarray, factor: float, stretch: bool, outlier: float = -1, engine=None):
    """What `singleImageInference` is fed, as (samples, PreMap, infer_shape).

    The raw integer samples always go to the GPU as they are; ``infer_shape`` (None at scalingFactor 1) is the size the
    library resizes them to on the fly (UnMicst1-5.py:813-815), the PreMap carries img_as_float and, with
    ``stretch`` (legacy / duo / Cyto2: rescale_intensity of the RESIZED image to (0, 0.983) with its max or the
    --outlier percentile, :817-821), the in_range.  solo feeds the un-stretched image (`cells = I`, :816).
    With an ``engine`` the min/max of the resized image come from the GPU (umx_resample_minmax); percentiles and
    engine-less calls resample on the host."""
    raw = coerce_raw(raw)
    scale = as_float_scale(raw.dtype)
    shape = scaled_shape(raw.shape, factor)
    infer_shape = None if shape == raw.shape else shape
    if not stretch:
        return raw, PreMap(in_scale=scale, rescale=False), infer_shape
    if infer_shape is None:
        lo = float(raw.min()) * scale
        top = float(raw.max()) * scale if outlier == -1 else float(np.percentile(raw.astype(np.float64) * scale, outlier))
    elif outlier == -1 and engine is not None and raw.dtype in (np.uint8, np.uint16, np.float32, np.float64):
        lo, top = engine.resample_minmax(raw, shape, scale)
    else:
        arr = resample(raw.astype(np.float64) * scale, shape)
        lo = float(arr.min())
        top = float(arr.max()) if outlier == -1 else float(np.percentile(arr, outlier))
    if top > lo:
        return raw, PreMap(in_scale=scale, rescale=True, imin=lo, imax=top, omin=0.0, omax=0.983), infer_shape
    return raw, PreMap(in_scale=scale, rescale=False), infer_shape


# uint8 -> resize (img_as_float: v * (1/255)) -> uint8(255 * x): the reference quantises twice
# (UnMicst1-5.py:848-853); some levels come back one lower.  Identity-size case as a table.
REQUANT_LUT = np.uint8(255 * (np.arange(256, dtype=np.float64) * (1.0 / 255)))


def back_to_raw_size(pm_u8: np.ndarray, raw_shape: Tuple[int, int]) -> np.ndarray:
    """resize(PM, (rawVert, rawHorz)) followed by np.uint8(255 * PM) for one uint8 page."""
    if pm_u8.shape == tuple(raw_shape):
        return REQUANT_LUT[pm_u8]
    return np.uint8(255 * resample(pm_u8.astype(np.float64) * (1.0 / 255), tuple(raw_shape)))


def preview_page(raw: np.ndarray) -> np.ndarray:
    """uint8(255 * im2double(raw)/max(im2double(raw))) — second page of the qc preview."""
    raw = coerce_raw(raw)
    d = raw.astype(np.float64)
    if raw.dtype == np.uint16:
        d = d / 65535
    elif raw.dtype == np.uint8:
        d = d / 255
    return np.uint8(255 * (d / np.max(d)))
