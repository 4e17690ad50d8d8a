"""Layer-by-layer comparison of the CUDA path with the oracle (bring-up aid; run on the GPU box)."""
import sys
import numpy as np

sys.path.insert(0, ".")
from oracle import unet_oracle
from unmicst_b200 import modelzoo
from unmicst_b200.engine import Engine

name = sys.argv[1] if len(sys.argv) > 1 else "nucleiDAPI1-5"
prec = sys.argv[2] if len(sys.argv) > 2 else "split3"
n = int(sys.argv[3]) if len(sys.argv) > 3 else 9
m = modelzoo.synthetic_model(name, seed=0)
rng = np.random.default_rng(7)
S, C = m.hp["imSize"], m.hp["nChannels"]
x = rng.normal(size=(n, S, S, C)).astype(np.float32)
taps = {}
want = unet_oracle.forward(m.weights, m.hp, m.variant, x, taps=taps)
L = m.hp["nLayers"]
names = {f"ld{i}": f"ld{i}.conv0" for i in range(L)}
names["lb"] = "lb.conv"
for i in range(L):
    names[f"lu{i}.up"] = f"lu{i}.convT"
    names[f"lu{i}"] = f"lu{i}.conv2"
with Engine(m, precision=prec) as e:
    got = e.forward_tiles(x)
    for k, v in taps.items():
        if k not in names:
            continue
        try:
            g = e.debug_buffer(names[k], n, v.shape[1:])
        except Exception as ex:
            print(f"{k:10s} {names[k]:12s} unavailable: {ex}")
            continue
        err = np.abs(g - v)
        print(f"{k:10s} {names[k]:12s} shape {v.shape[1:]} max|ref| {np.abs(v).max():8.3f}  max err {err.max():.3e}  mean err {err.mean():.3e}  nan {int(np.isnan(g).sum())}")
print("probs max err", np.abs(got - want).max(), "argmax agree", (got.argmax(-1) == want.argmax(-1)).mean())
