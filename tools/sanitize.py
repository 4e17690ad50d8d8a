"""Small end-to-end runs for compute-sanitizer (memcheck / racecheck / synccheck): a 300 x 300 unmicst-solo image and a
legacy (5x5, real checkpoint) crop, in the arithmetic given on the command line; results are still checked against each
other so that a run that "passes" the sanitizer but computes garbage is noticed.
usage: sanitize.py split3|single|mixed|fp32"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from unmicst_b200 import modelzoo
from unmicst_b200.engine import Engine, PreMap

prec = sys.argv[1] if len(sys.argv) > 1 else "split3"
rng = np.random.default_rng(0)
img = (rng.random((300, 300)) * 65535).astype(np.uint16)
solo = modelzoo.synthetic_model("nucleiDAPI1-5", seed=0)
legacy = modelzoo.load_model(os.path.join(ROOT, "tests", "golden", "models", "nucleiDAPI"))
for name, m, shape in (("solo", solo, None), ("solo x1.5 + resize back", solo, (450, 450)), ("legacy", legacy, None)):
    kw = dict(single_mask=0b101010101010) if prec == "mixed" else {}
    with Engine(m, 0, prec, 64, **kw) as e:
        u8, _ = e.infer_image(img, premap=PreMap(in_scale=1.0 / 65535), infer_shape=shape, cli_quant=shape is not None)
        parts = [b for _, _, b in e.stream_image(img, premap=PreMap(in_scale=1.0 / 65535), chunk_tile_rows=2, infer_shape=shape, cli_quant=shape is not None)]
        assert np.array_equal(np.concatenate(parts, axis=1), u8), name
        batch = e.infer_images([img[:100, :120], img[100:164, :64]], [PreMap(in_scale=1.0 / 65535)] * 2)
    s = u8.astype(int).sum(0)
    assert s.min() >= 250 and s.max() <= 256, (name, s.min(), s.max())
    print(f"{name} [{prec}]: ok, class means {u8.mean((1, 2)).round(1)}", flush=True)
