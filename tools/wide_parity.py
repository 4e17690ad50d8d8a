"""How far does the probe-tile error extrapolate?  max|dp| of the calibrated engine vs the full split over N tiles drawn
uniformly from the whole benchmark slide (per tile, not stitched), next to the 64-probe-tile figure.  GPU box.
usage: wide_parity.py [n_tiles] [budget]"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import make_image, choose_model, premap_for
from unmicst_b200.engine import Engine, calibrate, sample_probe_tiles, tile_geometry

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
budget = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-3
name = "nucleiDAPI1-5"
img = make_image("solo20k", 20000, 20000)
pm = premap_for(name, img)
probe = None
def probe_of(m):
    global probe
    if probe is None:
        probe = sample_probe_tiles(img, 64, 1, m.mean, m.std, pm, n=64)
    return probe
model, gain = choose_model(name, probe_of)
prec, terms, rep = calibrate(model, 0, probe_of(model), budget=budget)
_, _, npr, npc = tile_geometry(20000, 20000, 64)
idx = np.random.default_rng(7).choice(npr * npc, size=n, replace=False)
worst, worst_single, hist = 0.0, 0.0, np.zeros(8, dtype=np.int64)
with Engine(model, 0, prec, 2048, op_terms=terms) as e, Engine(model, 0, "split3", 2048) as ref, Engine(model, 0, "single", 2048) as es:
    for i in range(0, n, 2048):
        t = sample_probe_tiles(img, 64, 1, model.mean, model.std, pm, indices=idx[i:i + 2048])
        a, b, c = e.forward_tiles(t), ref.forward_tiles(t), es.forward_tiles(t)
        d = np.abs(a - b).max(axis=(1, 2, 3))
        worst = max(worst, float(d.max())); worst_single = max(worst_single, float(np.abs(c - b).max()))
        hist += np.histogram(d, bins=[0, 2.5e-4, 5e-4, 7.5e-4, 1e-3, 1.25e-3, 1.5e-3, 2e-3, 1.0])[0]
print(json.dumps({"tiles": n, "budget": budget, "probe_64_tiles_max_abs_dp_vs_split3": rep.get("mixed_vs_split3_max_abs_dp"),
                  "wide_max_abs_dp_vs_split3": worst, "wide_single_max_abs_dp_vs_split3": worst_single,
                  "per_tile_max_histogram_edges": [0, 2.5e-4, 5e-4, 7.5e-4, 1e-3, 1.25e-3, 1.5e-3, 2e-3, 1.0], "per_tile_max_histogram": hist.tolist(),
                  "op_terms": rep.get("op_terms")}))
