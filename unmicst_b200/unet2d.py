"""Drop-in for the reference's class-level inference API (UnMicst1-5.py:656-710):

    UNet2D.singleImageInferenceSetup(modelPath, gpuIndex, mean, std)
    UNet2D.singleImageInference(image, mode, pmIndex) -> float16 [H, W]
    UNet2D.singleImageInferenceCleanup()
    UNet2D.hp / UNet2D.DatasetMean / UNet2D.DatasetStDev

Same names, argument meaning and error behaviour (Python exceptions).  Additions: the network
runs ONCE for all classes (``singleImageInferenceAll``; per-class calls on the same image reuse
that result instead of re-running the network 2-3 times), and ``gpuIndex`` may be a
list of devices (tile-row bands across GPUs).
"""
from __future__ import annotations

import os
import zlib
from typing import Optional, Sequence, Union

import numpy as np

from . import modelzoo
from .engine import Engine, MultiEngine, PreMap, pick_gpu_most_free, sample_probe_tiles


class UNet2D:
    hp = None
    DatasetMean = 0
    DatasetStDev = 0
    Model = None
    Engine = None
    _cache_key = None
    _cache_val = None
    _pending = None          # (devices, precision): precision 'auto' builds the engine on the first image it sees

    @staticmethod
    def singleImageInferenceSetup(modelPath: str, gpuIndex: Union[int, Sequence[int]] = -1, mean: float = -1,
                                  std: float = -1, precision: str = "default", allow_synthetic: Optional[bool] = None):
        if allow_synthetic is None:
            allow_synthetic = os.environ.get("UNMICST_ALLOW_SYNTHETIC", "0") == "1"
        model = modelzoo.load_model(modelPath, mean, std, allow_synthetic=allow_synthetic)
        UNet2D.Model = model
        UNet2D.hp = model.hp
        UNet2D.DatasetMean = model.mean
        UNet2D.DatasetStDev = model.std
        print(UNet2D.DatasetMean)
        print(UNet2D.DatasetStDev)
        if isinstance(gpuIndex, (list, tuple)):
            devices = [int(d) for d in gpuIndex]
        else:
            devices = [pick_gpu_most_free() if gpuIndex is None or gpuIndex < 0 else int(gpuIndex)]
        UNet2D._cache_key = UNet2D._cache_val = None
        UNet2D._pending = (devices, precision)
        if precision != "auto":
            UNet2D._build_engine(None)
        print("Model restored.")

    @staticmethod
    def _build_engine(probe_tiles):
        devices, precision = UNet2D._pending
        if len(devices) > 1:
            UNet2D.Engine = MultiEngine(UNet2D.Model, devices, precision, probe_tiles=probe_tiles)
        else:
            UNet2D.Engine = Engine(UNet2D.Model, devices[0], precision, probe_tiles=probe_tiles)
        rep = getattr(UNet2D.Engine, "auto_report", None)
        if rep:
            print(f"precision auto -> {rep['chosen']}" + (f" (single-MMA layers: {', '.join(rep['single_layers'])})" if rep.get("single_layers") else "")
                  + f"; max|dp| vs split on {rep['probe_tiles']} tiles of this image: {rep.get('mixed_vs_split3_max_abs_dp', rep['single_vs_split3_max_abs_dp']):.1e}")

    @staticmethod
    def _ensure_engine(image, premap=None, infer_shape=None):
        """precision 'auto': calibrate on tiles of the image about to be processed (engine.calibrate)."""
        if UNet2D.Engine is None:
            if UNet2D._pending is None:
                raise RuntimeError("call UNet2D.singleImageInferenceSetup first")
            hp = UNet2D.hp
            tiles = sample_probe_tiles(image, int(hp["imSize"]), int(hp["nChannels"]), UNet2D.DatasetMean, UNet2D.DatasetStDev,
                                       premap, infer_shape=infer_shape)
            UNet2D._build_engine(tiles)
        return UNet2D.Engine

    @staticmethod
    def singleImageInferenceCleanup():
        if UNet2D.Engine is not None:
            UNet2D.Engine.close()
        UNet2D.Engine = None
        UNet2D._pending = None
        UNet2D._cache_key = UNet2D._cache_val = None

    @staticmethod
    def singleImageInferenceAll(image: np.ndarray, premap=None, as_uint8: bool = False,
                                infer_shape=None, cli_quant: bool = False):
        """All K class maps in one network pass: float32 [K,H,W] (or uint8 floor(255 p)).  ``premap``: one PreMap or one
        per input channel; ``infer_shape``: run at this size (--scalingFactor, resized on the GPU); ``cli_quant``: the
        uint8 pages the reference CLI writes (resized back to the raw grid, quantised twice, UnMicst1-5.py:848-853)."""
        UNet2D._ensure_engine(image, premap, infer_shape)
        as_uint8 = as_uint8 or cli_quant
        u8, f32 = UNet2D.Engine.infer_image(image, UNet2D.DatasetMean, UNet2D.DatasetStDev, premap=premap,
                                            want_u8=as_uint8, want_f32=not as_uint8, infer_shape=infer_shape, cli_quant=cli_quant)
        return u8 if as_uint8 else f32

    @staticmethod
    def singleImageInferenceStream(image: np.ndarray, premap=None, infer_shape=None, cli_quant: bool = True):
        """Generator of (row0, row1, uint8 [K, rows, W]) bands in row order, for writers that want to start before the
        last tile is done.  One GPU: bands stream off the device as they complete; several GPUs: one band (everything)."""
        UNet2D._ensure_engine(image, premap, infer_shape)
        if isinstance(UNet2D.Engine, Engine):
            yield from UNet2D.Engine.stream_image(image, UNet2D.DatasetMean, UNet2D.DatasetStDev, premap=premap,
                                                  infer_shape=infer_shape, cli_quant=cli_quant)
        else:
            u8 = UNet2D.singleImageInferenceAll(image, premap, True, infer_shape, cli_quant)
            yield 0, u8.shape[1], u8

    @staticmethod
    def singleImageInference(image: np.ndarray, mode: str = "accumulate", pmIndex: int = 0) -> np.ndarray:
        print("Inference...")
        if mode not in ("accumulate", "replace"):
            raise ValueError("mode must be 'accumulate' or 'replace' (PartitionOfImage.py:92-100)")
        # one pass serves every class: per-class calls on the very same array object reuse it (the reference's CLI asks
        # class by class, UnMicst1-5.py:845-849).  Identity + a full checksum: an in-place edit is noticed.
        a = np.asarray(image)
        UNet2D._ensure_engine(image)
        key = (mode, id(image), a.__array_interface__["data"][0], a.shape, a.dtype.str, a.strides, zlib.adler32(np.ascontiguousarray(a).data))
        if key != UNet2D._cache_key:
            _, UNet2D._cache_val = UNet2D.Engine.infer_image(image, UNet2D.DatasetMean, UNet2D.DatasetStDev, want_u8=False,
                                                             want_f32=True, stitch_mode=mode)
            UNet2D._cache_key = key
        return UNet2D._cache_val[pmIndex].astype(np.float16)
