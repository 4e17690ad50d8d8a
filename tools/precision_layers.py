"""Per-layer precision/cost study behind `--precision auto` (GPU box): for a steep-softmax stand-in model and probe
tiles cut from the synthetic slide, print what each tensor-path layer contributes to max|dp| when it alone runs with one
MMA per product, what it saves, which mask the error-budgeted selection picks, and how every mode compares with the
fp32 oracle on the same tiles.

usage: precision_layers.py [model] [logit_gain] [size] [budget]      -> JSON on stdout
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bench import make_image, WORKLOADS  # noqa: E402
from oracle import unet_oracle  # noqa: E402
from unmicst_b200 import modelzoo  # noqa: E402
from unmicst_b200.engine import Engine, PreMap, calibrate, sample_probe_tiles  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "nucleiDAPI1-5"
gain = float(sys.argv[2]) if len(sys.argv) > 2 else 12.0
size = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
budget = float(sys.argv[4]) if len(sys.argv) > 4 else 1e-3
wl = {"nucleiDAPI1-5": "solo20k", "nucleiDAPILAMIN": "duo4k", "CytoplasmIncell2": "cyto2tma", "nucleiDAPI": "legacy20k"}[name]

m = modelzoo.synthetic_model(name, seed=0, logit_gain=gain)
img = make_image(wl, size, size)
S, C = m.hp["imSize"], m.hp["nChannels"]
if name == "nucleiDAPI1-5":
    pm = PreMap(in_scale=1.0 / 65535)
else:
    pm = PreMap(in_scale=1.0 / 65535, rescale=True, imin=float(img.min()) / 65535, imax=float(img.max()) / 65535)
tiles = sample_probe_tiles(img, S, C, m.mean, m.std, pm, n=64)
taps = {}
want = unet_oracle.forward(m.weights, m.hp, m.variant, tiles, taps=taps)
out = {"model": name, "logit_gain": gain, "max_abs_logit": float(np.abs(taps["logits"]).max()), "tiles": int(len(tiles))}
prec, mask, rep = calibrate(m, 0, tiles, budget=budget, verbose=True)
out["calibration"] = rep
modes = [("split3", 0), ("single", 0)] + ([("mixed", mask)] if prec == "mixed" else [])
out["vs_oracle"] = {}
for p, k in modes:
    with Engine(m, 0, p, 64, single_mask=k) as e:
        got = e.forward_tiles(tiles)
    out["vs_oracle"][p] = {"max_abs_dp": float(np.abs(got - want).max()),
                           "argmax_agreement": float((got.argmax(-1) == want.argmax(-1)).mean())}
print(json.dumps(out))
