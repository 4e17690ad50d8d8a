"""ORACLE (test infrastructure): host pre/post-processing of the reference CLI
scripts restated on numpy/scipy — UnMicst1-5.py:807-825,848-850 and
UnMicst.py:621-633, toolbox/imtools.py:42-53.

The arithmetic lives in scikit-image (un-pinned dependency, Dockerfile:2; the
python:3.8 base resolves scikit-image 0.21), which is not installable here.
Restated from its published behaviour:
  resize(img, shape)   img_as_float (u16 * 1/65535, u8 * 1/255), order 1,
                       mode 'reflect' (= ndimage 'mirror' on the half-pixel
                       grid), scipy.ndimage.zoom(..., grid_mode=True); when
                       shrinking, a Gaussian of sigma=(factor-1)/2 first.
  rescale_intensity    clip to in_range, (x-imin)/(imax-imin)*(omax-omin)+omin
"parity unpinned" against scikit-image itself; the identity case (factor 1),
which is all the golden vectors exercise, is exact by construction.
"""
from __future__ import annotations

import numpy as np
from scipy import ndimage as ndi


def im2double(a: np.ndarray) -> np.ndarray:
    """toolbox/imtools.py:42-53."""
    if a.dtype == np.uint16:
        return a.astype(np.float64) / 65535
    if a.dtype == np.uint8:
        return a.astype(np.float64) / 255
    if a.dtype == np.float32:
        return a.astype(np.float64)
    return a


def img_as_float(a: np.ndarray) -> np.ndarray:
    """skimage.util.img_as_float: integer types are scaled by multiplying with 1/max."""
    if a.dtype == np.uint16:
        return a.astype(np.float64) * (1.0 / 65535)
    if a.dtype == np.uint8:
        return a.astype(np.float64) * (1.0 / 255)
    return a.astype(np.float64)


def resize(a: np.ndarray, shape) -> np.ndarray:
    """skimage.transform.resize(a, shape) with its defaults (order 1, 'reflect', anti-aliasing on shrink)."""
    img = img_as_float(a)
    out_h, out_w = int(shape[0]), int(shape[1])
    if (out_h, out_w) == img.shape:
        return img.copy()
    factors = (img.shape[0] / out_h, img.shape[1] / out_w)
    if any(f > 1 for f in factors):
        sigma = tuple(max(0.0, (f - 1) / 2) for f in factors)
        img = ndi.gaussian_filter(img, sigma, mode="mirror", cval=0)
    zoom = (out_h / img.shape[0], out_w / img.shape[1])
    out = ndi.zoom(img, zoom, order=1, mode="mirror", cval=0, grid_mode=True)
    return np.clip(out, img.min(), img.max())


def rescale_intensity(a: np.ndarray, in_range, out_range) -> np.ndarray:
    imin, imax = float(in_range[0]), float(in_range[1])
    omin, omax = float(out_range[0]), float(out_range[1])
    x = np.clip(a, imin, imax)
    if imin != imax:
        x = (x - imin) / (imax - imin)
        return x * (omax - omin) + omin
    return np.clip(x, omin, omax)


def prepare_rescaled(raw: np.ndarray, scaling_factor: float = 1.0, outlier: float = -1) -> np.ndarray:
    """UnMicst.py:621-631 / batchUNet2DtCycif.py:525-529: resize then stretch to (0, 0.983)."""
    h = int(float(raw.shape[0]) * float(scaling_factor))
    w = int(float(raw.shape[1]) * float(scaling_factor))
    img = resize(raw, (h, w))
    top = np.max(img) if outlier == -1 else np.percentile(img, outlier)
    return im2double(rescale_intensity(img, (np.min(img), top), (0, 0.983)))


def prepare_solo(raw: np.ndarray, scaling_factor: float = 1.0) -> np.ndarray:
    """UnMicst1-5.py:811-816: solo feeds the resized but un-stretched image (`cells = I`)."""
    h = int(float(raw.shape[0]) * float(scaling_factor))
    w = int(float(raw.shape[1]) * float(scaling_factor))
    return resize(raw, (h, w))


def preview_raw(raw: np.ndarray) -> np.ndarray:
    """UnMicst1-5.py:825 — rawI = im2double(raw)/max, later written as uint8(255*rawI)."""
    d = im2double(raw)
    return d / np.max(d)
