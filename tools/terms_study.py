"""Which hi/lo correction terms matter, per layer and per concat source (GPU box): max|dp| vs the full split on probe
tiles of the synthetic slide when ONE layer runs with a subset of the terms (UMX_TC_TERMS), and the layer's time."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_image, choose_model, premap_for
from unmicst_b200.engine import Engine, sample_probe_tiles

name, wl = (sys.argv[1], sys.argv[2]) if len(sys.argv) > 2 else ("nucleiDAPI1-5", "solo20k")
img = make_image(wl, 4096, 4096)
pm = premap_for(name, img)
from unmicst_b200.modelzoo import KNOWN_HP
S, C = KNOWN_HP[name]["imSize"], KNOWN_HP[name]["nChannels"]
probe = None
def probe_of(m):
    global probe
    if probe is None:
        probe = sample_probe_tiles(img, S, C, m.mean, m.std, pm, n=64)
    return probe
model, gain = choose_model(name, probe_of)
tiles = probe_of(model)
big = np.concatenate([tiles] * 16)
def run(env):
    os.environ.pop("UMX_TC_TERMS", None)
    if env: os.environ["UMX_TC_TERMS"] = env
    with Engine(model, 0, "split3", 1024) as e:
        out = e.forward_tiles(tiles)
        e.forward_tiles(big); e.profile_enable(True); e.forward_tiles(big); prof = {p["name"]: p["ms"] for p in e.profile_read()}
    return out, prof
ref, prof_ref = run(None)
layers = sys.argv[3].split(",") if len(sys.argv) > 3 else ["lu1.conv2", "lu0.convT", "lu0.conv2", "ld1.conv0", "lu1.convT", "lu2.conv2"]
res = {}
for l in layers:
    for t0 in range(4):
        for t1 in range(4):
            if (t0, t1) == (3, 3): continue
            if l.endswith("convT") or l.startswith("ld"):
                if t1 != 3: continue          # single source
            out, prof = run(f"{l}:{t0}:{t1}")
            key = [k for k in prof if k.split("+")[0] == l][0]
            res[f"{l}:{t0}:{t1}"] = (float(np.abs(out - ref).max()), prof[key], prof_ref[key])
            print(f"{l:10s} src0 terms {t0} src1 terms {t1}: dp {res[f'{l}:{t0}:{t1}'][0]:.2e}  {prof[key]:.3f} ms (full split {prof_ref[key]:.3f})", flush=True)
print(json.dumps({"gain": gain, "res": res}))
