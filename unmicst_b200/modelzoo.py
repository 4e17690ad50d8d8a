"""Model folders of the reference (``models/<name>``) without TensorFlow.

Covers what ``UNet2D.singleImageInferenceSetup`` reads (UnMicst1-5.py:656-681,
UnMicst.py:489-512): ``hp.data`` / ``datasetMean.data`` / ``datasetStDev.data``
pickles (toolbox/ftools.py:36-39) and the ``model.ckpt`` tensor bundle, plus
two things the reference never needed: telling the two graph generations apart
from tensor names, and a seeded synthetic weight generator for the model
folders whose ``.data`` shard is not shipped.
"""
from __future__ import annotations

import os
import pickle
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import tfbundle

HP_KEYS = ("imSize", "nClasses", "nChannels", "nExtraConvs", "nLayers", "featMapsFact",
           "downSampFact", "ks", "nOut0", "stdDev0", "batchSize")

LEGACY = "legacy"   # UnMicst.py graph: ReLU, BN after ReLU, 1x1 shortcut, extra convs
V2 = "v2"           # UnMicst1-5.py / UnMicst2.py / UnMicstCyto2.py graph: leaky, BN before act

# hp of the shipped model folders so shapes are known even when
# the folder itself is not at hand (bench and tests on the GPU box).
KNOWN_HP: Dict[str, Dict] = {
    "nucleiDAPI": dict(imSize=128, nClasses=3, nChannels=1, nExtraConvs=1, nLayers=2, featMapsFact=2,
                       downSampFact=2, ks=5, nOut0=16, stdDev0=0.03, batchSize=16),
    "nucleiDAPI1-5": dict(imSize=64, nClasses=3, nChannels=1, nExtraConvs=0, nLayers=4, featMapsFact=2,
                          downSampFact=2, ks=3, nOut0=80, stdDev0=0.03, batchSize=32),
    "nucleiDAPILAMIN": dict(imSize=128, nClasses=3, nChannels=2, nExtraConvs=0, nLayers=5, featMapsFact=2,
                            downSampFact=2, ks=3, nOut0=36, stdDev0=1e-06, batchSize=24),
    "CytoplasmIncell2": dict(imSize=256, nClasses=2, nChannels=1, nExtraConvs=0, nLayers=3, featMapsFact=2,
                             downSampFact=2, ks=3, nOut0=30, stdDev0=0.007, batchSize=16),
    "CytoplasmIncell": dict(imSize=128, nClasses=2, nChannels=1, nExtraConvs=1, nLayers=2, featMapsFact=2,
                            downSampFact=2, ks=3, nOut0=24, stdDev0=0.03, batchSize=16),
    "CytoplasmZeissNikon": dict(imSize=256, nClasses=2, nChannels=1, nExtraConvs=1, nLayers=3, featMapsFact=2,
                                downSampFact=2, ks=3, nOut0=24, stdDev0=0.03, batchSize=32),
    "mousenucleiDAPI": dict(imSize=256, nClasses=3, nChannels=1, nExtraConvs=1, nLayers=3, featMapsFact=2,
                            downSampFact=2, ks=3, nOut0=20, stdDev0=0.03, batchSize=16),
}
KNOWN_VARIANT = {"nucleiDAPI": LEGACY, "CytoplasmIncell": LEGACY, "CytoplasmZeissNikon": LEGACY,
                 "mousenucleiDAPI": LEGACY, "nucleiDAPI1-5": V2, "nucleiDAPILAMIN": V2, "CytoplasmIncell2": V2}
KNOWN_NORM = {"nucleiDAPI": (0.19808180266398068, 0.16236284911018245), "nucleiDAPI1-5": (0.34, 0.25),
              "nucleiDAPILAMIN": (0.18, 0.17), "CytoplasmIncell2": (0.07, 0.07),
              "CytoplasmIncell": (0.1454310746195677, 0.12094951659914749),
              "CytoplasmZeissNikon": (0.3110483062036952, 0.14476281597088098),
              "mousenucleiDAPI": (0.0942104550464887, 0.08848931361402539)}


def load_pickle(path: str):
    """``loadData`` of toolbox/ftools.py:36-39 (without the print)."""
    with open(path, "rb") as f:
        return pickle.load(f)


def channel_plan(hp: Dict) -> List[int]:
    """nOutX of UnMicst1-5.py:69-73 — [C, n0, n0*f, ...] (nLayers + 2 entries)."""
    n = [int(hp["nChannels"]), int(hp["nOut0"])]
    for _ in range(int(hp["nLayers"])):
        n.append(n[-1] * int(hp["featMapsFact"]))
    return n


def detect_variant(names) -> str:
    """Graph generation from tensor names alone."""
    names = set(names)
    if "downsampling/ld0/kernelD0" in names:
        return V2
    if "downsampling/ld0/kernel1" in names:
        return LEGACY
    raise ValueError("neither downsampling/ld0/kernelD0 (v2) nor downsampling/ld0/kernel1 (legacy) present")


def _bn_names(scope: str) -> List[str]:
    return [f"{scope}/{p}" for p in ("beta", "gamma", "moving_mean", "moving_variance")]


def expected_tensors(hp: Dict, variant: str) -> Dict[str, Tuple[int, ...]]:
    """Names and shapes the inference graph reads, derived from hp.

    legacy: UnMicst.py:80-171; v2: UnMicst1-5.py:83-222."""
    n = channel_plan(hp)
    L, k, E, K = int(hp["nLayers"]), int(hp["ks"]), int(hp["nExtraConvs"]), int(hp["nClasses"])
    t: Dict[str, Tuple[int, ...]] = {}
    if variant == LEGACY:
        for i in range(L):
            t[f"downsampling/ld{i}/kernel1"] = (k, k, n[i], n[i + 1])
            for e in range(E):
                t[f"downsampling/ld{i}/kernelExtra{e}"] = (k, k, n[i + 1], n[i + 1])
            t[f"downsampling/ld{i}/shortcutWeights"] = (1, 1, n[i], n[i + 1])
            scope = "batch_normalization" if i == 0 else f"batch_normalization_{i}"
            for nm in _bn_names(scope):
                t[nm] = (n[i + 1],)
        t["lb/kernel1"] = (k, k, n[L], n[L + 1])
        for i in range(L):
            t[f"upsampling/lu{i}/kernel1"] = (k, k, n[i + 1], n[i + 2])
            t[f"upsampling/lu{i}/kernel2"] = (k, k, n[i] + n[i + 1], n[i + 1])
            for e in range(E):
                t[f"upsampling/lu{i}/kernel2Extra{e}"] = (k, k, n[i + 1], n[i + 1])
        t["lt/kernel"] = (1, 1, n[1], K)
    elif variant == V2:
        for i in range(L):
            t[f"downsampling/ld{i}/kernelD{i}"] = (k, k, n[i], n[i + 1])
            for e in range(E):
                t[f"ld{i}/kernelExtra{e}"] = (k, k, n[i + 1], n[i + 1])
            t[f"ld{i}/shortcutWeights"] = (k, k, n[i], n[i + 1])
            for nm in _bn_names(f"ld{i}/batch_normalization"):
                t[nm] = (n[i + 1],)
        t["lb/kernel1"] = (k, k, n[L], n[L + 1])
        for nm in _bn_names("conv"):
            t[nm] = (n[L + 1],)
        for i in range(L):
            t[f"lu{i}/kernelU{i}"] = (k, k, n[i + 1], n[i + 2])
            t[f"lu{i}/kernel2"] = (k, k, n[i] + n[i + 1], n[i + 1])
            for nm in _bn_names(f"lu{i}/conv2"):
                t[nm] = (n[i + 1],)
            for e in range(E):
                t[f"lu{i}/kernel2Extra{e}"] = (k, k, n[i + 1], n[i + 1])
        t["lt/kernel"] = (1, 1, n[1], K)
        for nm in _bn_names("batch_normalization"):
            t[nm] = (K,)
    else:
        raise ValueError(f"unknown graph variant {variant!r}")
    return t


def synthetic_weights(hp: Dict, variant: str, seed: int = 0, logit_gain: float = 1.0) -> Dict[str, np.ndarray]:
    """Seeded stand-in weights with the real shapes (SURVEY.md App. F.4 recipe).

    Tensors are drawn in sorted-name order (the key order of a ``.index`` table).
    ``logit_gain`` > 1 scales the last linear map so the softmax becomes as steep
    as the real models' (stress variant)."""
    rng = np.random.default_rng(seed)
    shapes = expected_tensors(hp, variant)
    out: Dict[str, np.ndarray] = {}
    for name in sorted(shapes):
        shp = shapes[name]
        if name.endswith("gamma"):
            v = rng.uniform(0.8, 1.2, shp)
        elif name.endswith("beta") or name.endswith("moving_mean"):
            v = rng.normal(0.0, 0.1, shp)
        elif name.endswith("moving_variance"):
            v = rng.uniform(0.5, 1.5, shp)
        else:
            kh, kw = shp[0], shp[1]
            is_up = ("kernelU" in name) or (variant == LEGACY and "/lu" in name and name.endswith("kernel1"))
            cin = shp[3] if is_up else shp[2]
            v = rng.normal(0.0, 1.0 / np.sqrt(kh * kw * cin), shp)
        out[name] = np.ascontiguousarray(v, dtype=np.float32)
    if logit_gain != 1.0:
        key = "batch_normalization/gamma" if variant == V2 else "lt/kernel"
        out[key] = (out[key] * np.float32(logit_gain)).astype(np.float32)
        if variant == V2:
            out["batch_normalization/beta"] = (out["batch_normalization/beta"] * np.float32(logit_gain)).astype(np.float32)
    return out


@dataclass
class Model:
    """Everything ``singleImageInferenceSetup`` leaves in ``UNet2D.*``."""
    name: str
    hp: Dict
    variant: str
    weights: Dict[str, np.ndarray]
    mean: float
    std: float
    synthetic: bool = False

    @property
    def channels(self) -> List[int]:
        return channel_plan(self.hp)


def check_weights(hp: Dict, variant: str, weights: Dict[str, np.ndarray]) -> None:
    want = expected_tensors(hp, variant)
    for name, shp in want.items():
        if name not in weights:
            raise KeyError(f"checkpoint lacks tensor {name!r} required by the {variant} graph")
        if tuple(weights[name].shape) != tuple(shp):
            raise ValueError(f"tensor {name!r}: checkpoint shape {tuple(weights[name].shape)} != graph shape {shp} "
                             f"(hp.data does not describe this checkpoint?)")


def load_model(model_path: str, mean: float = -1, std: float = -1, ckpt_name: str = "model.ckpt",
               hp_override: Optional[Dict] = None, allow_synthetic: bool = False, seed: int = 0) -> Model:
    """Read a reference model folder.  mean/std == -1 mean "from the folder"
    exactly like UnMicst1-5.py:661-669.  With ``allow_synthetic`` a folder whose
    ``.data`` shard is absent gets seeded stand-in weights (flagged in the result)."""
    name = os.path.basename(os.path.normpath(model_path))
    hp = dict(load_pickle(os.path.join(model_path, "hp.data")))
    if hp_override:
        hp.update(hp_override)
    m = load_pickle(os.path.join(model_path, "datasetMean.data")) if mean == -1 else mean
    s = load_pickle(os.path.join(model_path, "datasetStDev.data")) if std == -1 else std
    prefix = os.path.join(model_path, ckpt_name)
    synthetic = False
    if os.path.exists(tfbundle.data_path(prefix)):
        weights = tfbundle.load_bundle(prefix)
        variant = detect_variant(weights.keys())
    else:
        if not allow_synthetic:
            raise FileNotFoundError(f"{tfbundle.data_path(prefix)} is missing; the reference downloads it at "
                                    f"docker-build time (Dockerfile:5-6)")
        variant = detect_variant(tfbundle.read_index(prefix + ".index").keys())
        weights = synthetic_weights(hp, variant, seed)
        synthetic = True
    check_weights(hp, variant, weights)
    return Model(name, hp, variant, weights, float(m), float(s), synthetic)


def synthetic_model(name: str, seed: int = 0, logit_gain: float = 1.0) -> Model:
    """A model of a known shipped architecture with seeded stand-in weights."""
    hp = dict(KNOWN_HP[name])
    variant = KNOWN_VARIANT[name]
    mean, std = KNOWN_NORM[name]
    return Model(name, hp, variant, synthetic_weights(hp, variant, seed, logit_gain), mean, std, True)
