"""max|dp| of the fp16 single-pass and hi/lo split3 tensor paths vs the fp32 oracle for several
softmax steepness levels (synthetic weights; run on the GPU box)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from oracle import unet_oracle
from unmicst_b200 import modelzoo
from unmicst_b200.engine import Engine

name = sys.argv[1] if len(sys.argv) > 1 else "nucleiDAPI1-5"
for gain in (1.0, 6.0, 12.0, 30.0):
    m = modelzoo.synthetic_model(name, seed=1, logit_gain=gain)
    rng = np.random.default_rng(8)
    S, C = m.hp["imSize"], m.hp["nChannels"]
    x = rng.normal(size=(16, S, S, C)).astype(np.float32)
    taps = {}
    want = unet_oracle.forward(m.weights, m.hp, m.variant, x, taps=taps)
    line = f"gain {gain:5.1f} max|logit| {np.abs(taps['logits']).max():6.1f}"
    for prec in ("fp32", "split3", "single"):
        with Engine(m, precision=prec) as e:
            got = e.forward_tiles(x)
        line += f" | {prec}: dp {np.abs(got - want).max():.2e} argmax {(got.argmax(-1) == want.argmax(-1)).mean():.5f}"
    print(line, flush=True)
