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
        img[5:50, 5:50] = 0.5                                    # an in-place edit must not be served from the cache
        n2 = UNet2D.Engine.launch_count
        c = UNet2D.singleImageInference(img, "accumulate", 1)
        assert UNet2D.Engine.launch_count > n2 and not np.array_equal(a, c)
        r = UNet2D.singleImageInference(img, "replace", 1)       # PI2D's other mode (PartitionOfImage.py:99-100)
        assert r.shape == img.shape and not np.array_equal(r, c)
        with pytest.raises(ValueError):
            UNet2D.singleImageInference(img, "average", 0)
    finally:
        UNet2D.singleImageInferenceCleanup()


def _write_pages(path, pages):
    for i, pg in enumerate(pages):
        tiffio.imsave(path, pg, append=i > 0)


def test_duo_cli_two_channels_end_to_end(tmp_path, golden_dir, sample_raw):
    """unmicst-duo (UnMicst2.py:692-835): two channel pages, each stretched with its own range, 3 classes, qc preview."""
    from unmicst_b200 import modelzoo
    dna = sample_raw[:300, :420]
    lamin = (sample_raw[400:700, 100:520] // 2).astype(np.uint16)
    img = str(tmp_path / "pair.ome.tif")
    _write_pages(img, [dna, np.zeros_like(dna), lamin])
    out = str(tmp_path / "duo")
    env = dict(os.environ, UNMICST_ALLOW_SYNTHETIC="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "unmicstWrapper.py"), "--tool", "unmicst-duo", "--channel", "1", "3",
                        "--stackOutput", "--outputPath", out, "--modelsDir", os.path.join(golden_dir, "models"), img],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Using channels 1 and 3" in r.stdout
    prob = os.path.join(out, "pair_Probabilities_1.tif")
    prev = os.path.join(out, "qc", "pair_Preview_1.tif")
    assert tiffio.count_pages(prob) == 3 and tiffio.count_pages(prev) == 2
    m = modelzoo.load_model(os.path.join(golden_dir, "models", "nucleiDAPILAMIN"), allow_synthetic=True)
    stack = np.stack([prepost_oracle.prepare_rescaled(dna), prepost_oracle.prepare_rescaled(lamin)])
    fw = lambda x: unet_oracle.forward(m.weights, m.hp, m.variant, x)
    pm = pi2d_oracle.infer_image(stack, fw, 128, 2, m.mean, m.std, 24, accum_dtype=np.float64)
    for page, cls in enumerate((2, 1, 0)):
        want = np.uint8(255 * prepost_oracle.resize(np.uint8(255 * pm[cls]), dna.shape))
        assert np.abs(tiffio.read_page(prob, page).astype(int) - want.astype(int)).max() <= 2
    assert np.array_equal(tiffio.read_page(prev, 0), tiffio.read_page(prob, 1))
    assert np.array_equal(tiffio.read_page(prev, 1), np.uint8(255 * prepost_oracle.preview_raw(lamin)))     # rawI = last page read


def test_cyto2_cli_end_to_end(tmp_path, golden_dir, sample_raw):
    """UnMicstCyto2 (UnMicstCyto2.py:679-827): 2 classes, 0-based file suffix, no qc folder, --precision auto."""
    from unmicst_b200 import modelzoo
    crop = sample_raw[100:500, 200:680]
    img = str(tmp_path / "core.tif")
    tiffio.imsave(img, crop)
    out = str(tmp_path / "cyto")
    env = dict(os.environ, UNMICST_ALLOW_SYNTHETIC="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "unmicstWrapper.py"), "--tool", "UnMicstCyto2", "--channel", "1",
                        "--stackOutput", "--outputPath", out, "--precision", "auto",
                        "--modelsDir", os.path.join(golden_dir, "models"), img],
                       capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "precision auto ->" in r.stdout
    prob = os.path.join(out, "core_Probabilities_0.tif")
    assert tiffio.count_pages(prob) == 2 and not os.path.exists(os.path.join(out, "qc"))
    m = modelzoo.load_model(os.path.join(golden_dir, "models", "CytoplasmIncell2"), allow_synthetic=True)
    fw = lambda x: unet_oracle.forward(m.weights, m.hp, m.variant, x)
    pm = pi2d_oracle.infer_image(prepost_oracle.prepare_rescaled(crop), fw, 256, 1, m.mean, m.std, 16, accum_dtype=np.float64)
    for page, cls in enumerate((1, 0)):
        want = np.uint8(255 * prepost_oracle.resize(np.uint8(255 * pm[cls]), crop.shape))
        assert np.abs(tiffio.read_page(prob, page).astype(int) - want.astype(int)).max() <= 2
    assert np.array_equal(tiffio.read_page(os.path.join(out, "core_Preview_0.tif"), 1), np.uint8(255 * prepost_oracle.preview_raw(crop)))


def test_multi_gpu_auto_and_scaling_match_single_gpu(sample_raw):
    """--gpus 2 with --precision auto and --scalingFactor: one calibration for every band, bytes equal to one GPU."""
    from unmicst_b200 import modelzoo
    from unmicst_b200.engine import Engine, MultiEngine, PreMap, device_count, sample_probe_tiles
    if device_count() < 2:
        pytest.skip("needs 2 GPUs")
    m = modelzoo.synthetic_model("nucleiDAPI1-5", seed=0, logit_gain=20.0)
    crop = sample_raw[:400, :500]
    pm = PreMap(in_scale=1.0 / 65535)
    shape = (800, 1000)
    tiles = sample_probe_tiles(crop, 64, 1, m.mean, m.std, pm, n=24, infer_shape=shape)
    me = MultiEngine(m, [0, 1], "auto", probe_tiles=tiles)
    try:
        two, _ = me.infer_image(crop, premap=pm, infer_shape=shape, cli_quant=True)
        with Engine(m, 0, me.precision, op_terms=me.op_terms) as e:
            one, _ = e.infer_image(crop, premap=pm, infer_shape=shape, cli_quant=True)
    finally:
        me.close()
    assert two.shape == (3,) + crop.shape and np.array_equal(one, two)


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
