"""ORACLE (test infrastructure): NumPy restatement of the reference's 2-D tiler and
ramp-weighted overlap stitcher, toolbox/PartitionOfImage.py:23-122 (class PI2D),
and of the host tile loop that drives it, UnMicst1-5.py:687-710 /
UnMicst.py:520-541 / UnMicst2.py:666-689.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs use it.

Contract restated:
  P = imSize, m = int(P/8), sub = P - 2m                      PartitionOfImage.py:25-28
  W[r,c] = min(1, min(r, P-1-r, c, P-1-c) / (2m))            :30-39 (ring i -> i/(2m), border 0)
  npr = ceil(H/sub), npc = ceil(W/sub)                        :49-50
  padded frame (npr*sub+2m) x (npc*sub+2m), zeros, image at (m,m)   :52-63
  tile (i,j), row-major: rows [i*sub, i*sub+P), cols [j*sub, j*sub+P)   :65-72
  accumulate: Count[tile] += W ; Output[tile] += P*W  (float16 storage)  :86-98
  result: Output[m:m+H, m:m+W] / Count[m:m+H, m:m+W]          :108-115
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Tuple

import numpy as np


def ramp_weight(patch: int, margin: int) -> np.ndarray:
    """The P x P blending window built ring by ring at PartitionOfImage.py:30-39."""
    w = np.ones((patch, patch), dtype=np.float64)
    w[0, :] = w[-1, :] = 0.0
    w[:, 0] = w[:, -1] = 0.0
    for ring in range(1, 2 * margin):
        v = ring / (2 * margin)
        lo, hi = ring, patch - 1 - ring
        w[lo, lo:hi + 1] = v
        w[hi, lo:hi + 1] = v
        w[lo:hi + 1, lo] = v
        w[lo:hi + 1, hi] = v
    return w


@dataclass
class TileGrid:
    patch: int
    margin: int
    sub: int
    rows: int          # image H
    cols: int          # image W
    npr: int
    npc: int
    frame_rows: int
    frame_cols: int

    @property
    def num_tiles(self) -> int:
        return self.npr * self.npc

    def origin(self, t: int) -> Tuple[int, int]:
        i, j = divmod(t, self.npc)
        return i * self.sub, j * self.sub


def tile_grid(rows: int, cols: int, patch: int, margin: int) -> TileGrid:
    sub = patch - 2 * margin
    npr = int(np.ceil(rows / sub))
    npc = int(np.ceil(cols / sub))
    return TileGrid(patch, margin, sub, rows, cols, npr, npc, npr * sub + 2 * margin, npc * sub + 2 * margin)


def pad_frame(image: np.ndarray, g: TileGrid) -> np.ndarray:
    """Zero frame with the image at offset (m, m); [H,W] or [C,H,W] (PartitionOfImage.py:56-63)."""
    m = g.margin
    if image.ndim == 2:
        f = np.zeros((g.frame_rows, g.frame_cols), dtype=np.float64)
        f[m:m + g.rows, m:m + g.cols] = image
    else:
        f = np.zeros((image.shape[0], g.frame_rows, g.frame_cols), dtype=np.float64)
        f[:, m:m + g.rows, m:m + g.cols] = image
    return f


def cut_tile(frame: np.ndarray, g: TileGrid, t: int) -> np.ndarray:
    r0, c0 = g.origin(t)
    return frame[..., r0:r0 + g.patch, c0:c0 + g.patch]


def analytic_count(g: TileGrid) -> np.ndarray:
    """Sum of the ramp windows of every tile over the padded frame, in float64 —
    what ``Count`` holds before its float16 rounding (geometry only)."""
    w = ramp_weight(g.patch, g.margin)
    cnt = np.zeros((g.frame_rows, g.frame_cols), dtype=np.float64)
    for t in range(g.num_tiles):
        r0, c0 = g.origin(t)
        cnt[r0:r0 + g.patch, c0:c0 + g.patch] += w
    return cnt


def infer_image(image: np.ndarray, forward: Callable[[np.ndarray], np.ndarray], patch: int, n_channels: int,
                mean: float, std: float, batch: int, accum_dtype=np.float16,
                classes: List[int] = None, mode: str = "accumulate") -> np.ndarray:
    """singleImageInference for all requested classes in ONE network pass per batch.

    ``forward`` maps [B,P,P,C] float32 -> [B,P,P,K] float32 (the Session.run of
    UnMicst1-5.py:704).  Per tile: (patch - mean)/std in float64 including the zero
    padding (:700), cast to float32 by the feed, batches of ``batch`` tiles
    (hp['batchSize'], :697-703).  Stitching follows patchOutput/getValidOutput with
    ``accum_dtype`` storage (float16 = the reference's behaviour; float32/float64 =
    the variant the shipped goldens actually agree with best).
    ``mode`` 'replace' (PartitionOfImage.py:99-100, never used by the CLI): Output[tile] = P, no weights, no divide.
    Returns [K', H, W] in accum_dtype, K' = len(classes) (all classes by default)."""
    rows, cols = image.shape[-2:]
    g = tile_grid(rows, cols, patch, int(patch / 8))
    frame = pad_frame(image, g)
    w = ramp_weight(g.patch, g.margin)
    count = np.zeros((g.frame_rows, g.frame_cols), dtype=accum_dtype)
    out = None
    pending: List[int] = []
    feed = np.zeros((batch, patch, patch, n_channels), dtype=np.float64)
    for t in range(g.num_tiles):
        p = (cut_tile(frame, g, t) - mean) / std
        j = len(pending)
        if p.ndim == 2:
            for c in range(n_channels):
                feed[j, :, :, c] = p
        else:
            for c in range(n_channels):
                feed[j, :, :, c] = p[c]
        pending.append(t)
        if len(pending) == batch or t == g.num_tiles - 1:
            probs = forward(feed.astype(np.float32))
            if classes is None:
                classes = list(range(probs.shape[-1]))
            if out is None:
                out = np.zeros((len(classes), g.frame_rows, g.frame_cols), dtype=accum_dtype)
            for k, tt in enumerate(pending):
                r0, c0 = g.origin(tt)
                if mode == "replace":
                    for ci, cls in enumerate(classes):
                        out[ci, r0:r0 + patch, c0:c0 + patch] = probs[k, :, :, cls]
                    continue
                count[r0:r0 + patch, c0:c0 + patch] += w
                for ci, cls in enumerate(classes):
                    out[ci, r0:r0 + patch, c0:c0 + patch] += np.multiply(probs[k, :, :, cls], w)
            pending = []
    m = g.margin
    if mode == "replace":
        return out[:, m:m + rows, m:m + cols]
    c = count[m:m + rows, m:m + cols]
    return np.divide(out[:, m:m + rows, m:m + cols], c)


def quantize_u8(p: np.ndarray) -> np.ndarray:
    """np.uint8(255 * PM) of UnMicst1-5.py:848 (product keeps the array's dtype, then truncates)."""
    return np.uint8(255 * p)
