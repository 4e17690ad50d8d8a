"""-m gpu: the command-line front-ends end to end (files in, BigTIFF pages out) and the
reference-shaped Python API."""
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import pi2d_oracle, prepost_oracle, unet_oracle
from unmicst_b200 import tiffio

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(script, *args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, script)] + list(args), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    return r.stdout


def test_wrapper_legacy_stack_output_matches_goldens(tmp_path, golden_dir, sample_goldens):
    img = os.path.join(golden_dir, "sample", "105.tif")
    out = str(tmp_path / "out")
    _run("unmicstWrapper.py", "--tool", "unmicst-legacy", "--channel", "1", "--stackOutput", "--outputPath", out,
         "--modelsDir", os.path.join(golden_dir, "models"), img)
    prob = os.path.join(out, "105_Probabilities_1.tif")
    prev = os.path.join(out, "qc", "105_Preview_1.tif")
    assert tiffio.count_pages(prob) == 3 and tiffio.count_pages(prev) == 2
    nuclei, contours, background = (tiffio.read_page(prob, i) for i in range(3))      # classOrder[::-1]
    # goldens were quantised once (batch script); the CLI quantises twice: +-1 each
    for got, key in ((nuclei, "nuclei"), (contours, "contours")):
        d = np.abs(got.astype(int) - sample_goldens[key].astype(int))
        assert d.max() <= 2 and (d > 1).mean() < 1e-3
    s = nuclei.astype(int) + contours.astype(int) + background.astype(int)
    assert s.min() >= 250 and s.max() <= 255
    assert np.array_equal(tiffio.read_page(prev, 0), contours)
    assert np.array_equal(tiffio.read_page(prev, 1), sample_goldens["raw"])


def test_legacy_script_without_stack_output(tmp_path, golden_dir, sample_goldens):
    img = os.path.join(golden_dir, "sample", "105.tif")
    out = str(tmp_path / "o2")
    _run("UnMicst.py", img, "--channel", "0", "--outputPath", out, "--modelsDir", os.path.join(golden_dir, "models"))
    c = os.path.join(out, "105_ContoursPM_1.tif")
    n = os.path.join(out, "105_NucleiPM_1.tif")
    assert tiffio.count_pages(c) == 2 and tiffio.count_pages(n) == 1
    assert np.abs(tiffio.read_page(c, 0).astype(int) - sample_goldens["contours"].astype(int)).max() <= 2
    assert np.array_equal(tiffio.read_page(c, 1), sample_goldens["raw"])
    assert np.abs(tiffio.read_page(n, 0).astype(int) - sample_goldens["nuclei"].astype(int)).max() <= 2


def test_solo_cli_synthetic_weights_and_scaling_factor(tmp_path, golden_dir, sample_raw):
    """unmicst-solo graph (weights not shipped -> seeded stand-ins), --scalingFactor 2 on a crop."""
    from unmicst_b200 import modelzoo
    crop = sample_raw[100:260, 200:420]
    img = str(tmp_path / "crop.ome.tif")
    tiffio.imsave(img, crop)
    out = str(tmp_path / "o3")
    env = dict(os.environ, UNMICST_ALLOW_SYNTHETIC="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "unmicstWrapper.py"), "--channel", "1", "--stackOutput",
                        "--scalingFactor", "2", "--outputPath", out, "--modelsDir", os.path.join(golden_dir, "models"), img],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    prob = os.path.join(out, "crop_Probabilities_1.tif")
    assert tiffio.count_pages(prob) == 3
    m = modelzoo.load_model(os.path.join(golden_dir, "models", "nucleiDAPI1-5"), allow_synthetic=True)
    cells = prepost_oracle.prepare_solo(crop, 2.0)
    fw = lambda x: unet_oracle.forward(m.weights, m.hp, m.variant, x)
    pm = pi2d_oracle.infer_image(cells, fw, 64, 1, m.mean, m.std, 32, accum_dtype=np.float64)
    for page, cls in enumerate((2, 1, 0)):
        want = np.uint8(255 * prepost_oracle.resize(np.uint8(255 * pm[cls]), crop.shape))
        got = tiffio.read_page(prob, page)
        assert got.shape == crop.shape
        assert np.abs(got.astype(int) - want.astype(int)).max() <= 2


def test_unet2d_api_shape_and_caching(golden_dir, sample_raw):
    from unmicst_b200.unet2d import UNet2D
    UNet2D.singleImageInferenceSetup(os.path.join(golden_dir, "models", "nucleiDAPI"), 0, -1, -1)
    try:
        assert UNet2D.hp["imSize"] == 128 and abs(UNet2D.DatasetMean - 0.19808) < 1e-4
        img = prepost_oracle.prepare_rescaled(sample_raw)[:300, :400].copy()
        n0 = UNet2D.Engine.launch_count
        a = UNet2D.singleImageInference(img, "accumulate", 1)
        n1 = UNet2D.Engine.launch_count
        b = UNet2D.singleImageInference(img, "accumulate", 2)
        assert UNet2D.Engine.launch_count == n1 > n0          # second class reuses the single pass
        assert a.dtype == np.float16 and a.shape == img.shape and b.shape == img.shape
        allp = UNet2D.singleImageInferenceAll(img)
        assert np.allclose(allp.sum(0), 1, atol=1e-5)
        with pytest.raises(NotImplementedError):
            UNet2D.singleImageInference(img, "replace", 0)
    finally:
        UNet2D.singleImageInferenceCleanup()


def test_multi_gpu_bands_match_single_gpu(nuclei_model, sample_raw):
    from unmicst_b200.engine import Engine, MultiEngine, device_count
    if device_count() < 2:
        pytest.skip("needs 2 GPUs")
    img = prepost_oracle.prepare_rescaled(sample_raw)
    with Engine(nuclei_model, 0) as e:
        one, _ = e.infer_image(img)
    me = MultiEngine(nuclei_model, [0, 1])
    try:
        two, _ = me.infer_image(img)
    finally:
        me.close()
    assert np.array_equal(one, two)
