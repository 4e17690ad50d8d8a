"""Time whole-network variants of the per-source correction terms (GPU box): op_terms given as name=t0t1 pairs."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unmicst_b200 import modelzoo
from unmicst_b200.engine import Engine, tensor_ops
m = modelzoo.synthetic_model("nucleiDAPI1-5", seed=0, logit_gain=14.83)
ops = dict((n, i) for i, n in tensor_ops(m))
x = np.random.default_rng(0).normal(size=(2048, 64, 64, 1)).astype(np.float32)
base = {n: 0 for n in ops}
base.update({"lu1.conv2": 3, "lu0.convT": 15, "lu0.conv2": 2 | 3 << 2})
def run(label, over):
    t = dict(base); t.update(over)
    with Engine(m, 0, "mixed", 2048, op_terms={ops[n]: v for n, v in t.items()}) as e:
        e.forward_tiles(x); e.profile_enable(True); e.forward_tiles(x)
        prof = {p["name"].split("+")[0]: p["ms"] for p in e.profile_read()}
        info = {n: e.op_info(ops[n]) for n in ("lu0.convT", "lu0.conv2", "lu1.conv2")}
    print(f"{label:34s} total {sum(prof.values()):7.2f} ms | " + " ".join(f"{n} {prof[n]:.2f} (res {info[n]['resident']} st {info[n]['stages']}/{info[n]['b_stages']})" for n in info), flush=True)
run("base lu0c (2,3) lu0T 15", {})
run("lu0c (2,2)", {"lu0.conv2": 2 | 2 << 2})
run("lu0c (2,0)", {"lu0.conv2": 2})
run("lu0T 2", {"lu0.convT": 2 | 2 << 2})
run("lu0T 2, lu0c (2,2)", {"lu0.convT": 10, "lu0.conv2": 10})
run("lu1c (2,0)", {"lu1.conv2": 2})
run("all single", {"lu1.conv2": 0, "lu0.convT": 0, "lu0.conv2": 0})
