"""ORACLE (test infrastructure, not product code): CPU fp32 restatement of the
two UNet2D inference graphs of the reference, op for op, on torch-CPU.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The product path (unmicst_b200/) never
does and fails loudly when its CUDA library is missing.

Graphs restated (TensorFlow itself is an un-vendored dependency of the
reference — Dockerfile:1 pins tensorflow 2.7.1, conda.yml:4 pins 1.15 — and
cannot be installed here, so the op semantics below are TF's published ones,
SURVEY.md App. A.3):
  legacy  UnMicst.py:51-187      (ReLU, BN after ReLU, 1x1 shortcut, extra convs)
  v2      UnMicst1-5.py:55-237   (leaky 0.2, BN before activation, kxk shortcut,
                                  BN on bottom / conv2 / logits); identical at
                                  inference in UnMicst2.py:52-235 and
                                  UnMicstCyto2.py:49-232
Pinned by tests/test_golden.py against the reference's shipped outputs
(UNet sample data/prob_maps/*.tif) for the legacy graph + models/nucleiDAPI.
The v2 graph has no golden vectors in the reference (weights not shipped):
"parity unpinned" for v2 beyond sharing every primitive with the pinned legacy
graph and the serialized-graph attributes (eps=1e-3, alpha=0.2) read from
models/*/model.ckpt.meta.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3       # FusedBatchNorm[V3] epsilon in every shipped .meta
LEAKY_ALPHA = 0.2   # LeakyRelu alpha / Maximum(alpha*x, x) const in the v2 .meta files


def _t(w: np.ndarray, dtype) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(w)).to(dtype)


def conv_same(x: torch.Tensor, w_hwio: torch.Tensor) -> torch.Tensor:
    """tf.nn.conv2d(NHWC, HWIO, strides 1, 'SAME') for odd k, on NCHW torch tensors."""
    k = w_hwio.shape[0]
    return F.conv2d(x, w_hwio.permute(3, 2, 0, 1), padding=(k - 1) // 2)


def conv_transpose_same_s2(x: torch.Tensor, w_hwoi: torch.Tensor) -> torch.Tensor:
    """tf.nn.conv2d_transpose(x, W[h,w,Cout,Cin], [B,2M,2M,Cout], strides 2, 'SAME').

    Gradient of a stride-2 SAME conv: out[2i-pb+a, 2j-pb+b, co] += x[i,j,ci]*W[a,b,co,ci]
    with pb=(k-2)//2; rows/cols outside [0,2M) are dropped (UnMicst1-5.py:192-195)."""
    k = w_hwoi.shape[0]
    m = x.shape[-1]
    pb = (k - 2) // 2
    full = F.conv_transpose2d(x, w_hwoi.permute(3, 2, 0, 1), stride=2)
    return full[..., pb:pb + 2 * m, pb:pb + 2 * m]


def batch_norm_inference(x: torch.Tensor, w: Dict[str, torch.Tensor], scope: str) -> torch.Tensor:
    """tf.layers.batch_normalization(training=False): gamma*(x-mu)/sqrt(var+eps)+beta."""
    g, b = w[scope + "/gamma"], w[scope + "/beta"]
    mu, var = w[scope + "/moving_mean"], w[scope + "/moving_variance"]
    scale = g / torch.sqrt(var + BN_EPS)
    return x * scale.view(1, -1, 1, 1) + (b - mu * scale).view(1, -1, 1, 1)


def _channels(hp: Dict) -> List[int]:
    n = [int(hp["nChannels"]), int(hp["nOut0"])]
    for _ in range(int(hp["nLayers"])):
        n.append(n[-1] * int(hp["featMapsFact"]))
    return n


def forward(weights: Dict[str, np.ndarray], hp: Dict, variant: str, tiles_nhwc: np.ndarray,
            dtype=torch.float32, taps: Optional[Dict[str, np.ndarray]] = None) -> np.ndarray:
    """Session.run(UNet2D.nn, {tfData: tiles, tfTraining: 0}) — [B,S,S,C] -> [B,S,S,K] softmax.

    ``taps`` (optional dict) receives named intermediate activations (NHWC numpy)
    for layer-by-layer debugging of the CUDA path."""
    w = {k: _t(v, dtype) for k, v in weights.items()}
    L, E = int(hp["nLayers"]), int(hp["nExtraConvs"])
    x = torch.from_numpy(np.ascontiguousarray(tiles_nhwc)).to(dtype).permute(0, 3, 1, 2).contiguous()

    def tap(name, t):
        if taps is not None:
            taps[name] = t.permute(0, 2, 3, 1).contiguous().numpy()

    ds = [x]
    with torch.no_grad():
        if variant == "legacy":
            for i in range(L):
                src = ds[i]
                c = conv_same(src, w[f"downsampling/ld{i}/kernel1"])
                for e in range(E):
                    c = conv_same(F.relu(c), w[f"downsampling/ld{i}/kernelExtra{e}"])
                s = conv_same(src, w[f"downsampling/ld{i}/shortcutWeights"])
                scope = "batch_normalization" if i == 0 else f"batch_normalization_{i}"
                y = batch_norm_inference(F.relu(c + s), w, scope)
                ds.append(F.max_pool2d(y, 2))
                tap(f"ld{i}", ds[-1])
            u = F.relu(conv_same(ds[L], w["lb/kernel1"]))
            tap("lb", u)
            for i in range(L - 1, -1, -1):
                us = F.relu(conv_transpose_same_s2(u, w[f"upsampling/lu{i}/kernel1"]))
                tap(f"lu{i}.up", us)
                cc = torch.cat([ds[i], us], dim=1)
                u = F.relu(conv_same(cc, w[f"upsampling/lu{i}/kernel2"]))
                for e in range(E):
                    u = F.relu(conv_same(u, w[f"upsampling/lu{i}/kernel2Extra{e}"]))
                tap(f"lu{i}", u)
            t = conv_same(u, w["lt/kernel"])
        elif variant == "v2":
            lk = lambda z: F.leaky_relu(z, LEAKY_ALPHA)
            for i in range(L):
                src = ds[i]
                c = conv_same(src, w[f"downsampling/ld{i}/kernelD{i}"])
                for e in range(E):
                    c = conv_same(lk(c), w[f"ld{i}/kernelExtra{e}"])
                s = conv_same(src, w[f"ld{i}/shortcutWeights"])
                y = lk(batch_norm_inference(c + s, w, f"ld{i}/batch_normalization"))
                ds.append(F.max_pool2d(y, 2))
                tap(f"ld{i}", ds[-1])
            u = lk(batch_norm_inference(conv_same(ds[L], w["lb/kernel1"]), w, "conv"))
            tap("lb", u)
            for i in range(L - 1, -1, -1):
                us = lk(conv_transpose_same_s2(u, w[f"lu{i}/kernelU{i}"]))
                tap(f"lu{i}.up", us)
                cc = torch.cat([ds[i], us], dim=1)
                u = lk(batch_norm_inference(conv_same(cc, w[f"lu{i}/kernel2"]), w, f"lu{i}/conv2"))
                for e in range(E):
                    u = lk(conv_same(u, w[f"lu{i}/kernel2Extra{e}"]))
                tap(f"lu{i}", u)
            t = batch_norm_inference(conv_same(u, w["lt/kernel"]), w, "batch_normalization")
        else:
            raise ValueError(f"unknown graph variant {variant!r}")
        tap("logits", t)
        p = torch.softmax(t, dim=1)
    return p.permute(0, 2, 3, 1).contiguous().to(torch.float32).numpy()
