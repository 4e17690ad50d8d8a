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
from typing import Optional, Sequence, Union

import numpy as np

from . import modelzoo
from .engine import Engine, MultiEngine, PreMap, pick_gpu_most_free


class UNet2D:
    hp = None
    DatasetMean = 0
    DatasetStDev = 0
    Model = None
    Engine = None
    _cache_key = None
    _cache_val = None

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
            UNet2D.Engine = MultiEngine(model, list(gpuIndex), precision) if len(gpuIndex) > 1 else Engine(model, gpuIndex[0], precision)
        else:
            dev = pick_gpu_most_free() if gpuIndex is None or gpuIndex < 0 else int(gpuIndex)
            UNet2D.Engine = Engine(model, dev, precision)
        UNet2D._cache_key = UNet2D._cache_val = None
        print("Model restored.")

    @staticmethod
    def singleImageInferenceCleanup():
        if UNet2D.Engine is not None:
            UNet2D.Engine.close()
        UNet2D.Engine = None
        UNet2D._cache_key = UNet2D._cache_val = None

    @staticmethod
    def _key(image: np.ndarray, premap):
        a = np.asarray(image)
        probe = a.reshape(-1)[:: max(1, a.size // 4099)]
        return (a.__array_interface__["data"][0], a.shape, a.dtype.str, a.strides, float(np.sum(probe, dtype=np.float64)),
                None if premap is None else tuple(premap.__dict__.values()))

    @staticmethod
    def singleImageInferenceAll(image: np.ndarray, premap: Optional[PreMap] = None, as_uint8: bool = False):
        """All K class maps in one network pass: float32 [K,H,W] (or uint8 floor(255 p))."""
        if UNet2D.Engine is None:
            raise RuntimeError("call UNet2D.singleImageInferenceSetup first")
        u8, f32 = UNet2D.Engine.infer_image(image, UNet2D.DatasetMean, UNet2D.DatasetStDev, premap=premap,
                                            want_u8=as_uint8, want_f32=not as_uint8)
        return u8 if as_uint8 else f32

    @staticmethod
    def singleImageInference(image: np.ndarray, mode: str = "accumulate", pmIndex: int = 0) -> np.ndarray:
        print("Inference...")
        if mode != "accumulate":
            raise NotImplementedError("only the 'accumulate' stitching mode of PI2D is implemented (the CLI never uses 'replace')")
        key = UNet2D._key(image, None)
        if key != UNet2D._cache_key:
            UNet2D._cache_val = UNet2D.singleImageInferenceAll(image)
            UNet2D._cache_key = key
        return UNet2D._cache_val[pmIndex].astype(np.float16)
