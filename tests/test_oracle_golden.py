"""Pins the oracle: CPU restatement vs the reference's own shipped outputs
(UNet sample data/prob_maps/*, produced by batchUNet2DtCycif.py:525-551 with the
legacy graph + models/nucleiDAPI).  Tolerance: +-1 uint8 level on every pixel."""
import numpy as np

from oracle import pi2d_oracle, prepost_oracle, unet_oracle


def _run(raw, model, accum):
    img = prepost_oracle.prepare_rescaled(raw)
    fw = lambda x: unet_oracle.forward(model.weights, model.hp, model.variant, x)
    return pi2d_oracle.infer_image(img, fw, model.hp["imSize"], model.hp["nChannels"], model.mean, model.std,
                                   model.hp["batchSize"], accum_dtype=accum)


def test_oracle_matches_shipped_probability_maps(sample_raw, sample_goldens, nuclei_model):
    pm = _run(sample_raw, nuclei_model, np.float16)          # the reference's fp16 accumulators
    for cls, key in ((1, "contours"), (2, "nuclei")):
        got = np.uint8(255 * pm[cls].astype(np.float64))
        d = np.abs(got.astype(int) - sample_goldens[key].astype(int))
        assert d.max() <= 1
        assert (d > 0).mean() < 0.02


def test_oracle_fp32_stitch_is_closer_to_goldens(sample_raw, sample_goldens, nuclei_model):
    pm = _run(sample_raw, nuclei_model, np.float32)
    for cls, key in ((1, "contours"), (2, "nuclei")):
        got = np.uint8(255 * pm[cls].astype(np.float64))
        d = np.abs(got.astype(int) - sample_goldens[key].astype(int))
        assert d.max() <= 1
        assert (d > 0).mean() < 0.006


def test_preview_page_is_raw_over_max(sample_raw, sample_goldens):
    got = np.uint8(255 * prepost_oracle.preview_raw(sample_raw))
    assert np.array_equal(got, sample_goldens["raw"])


def test_softmax_rows_sum_to_one_and_batch_independent(nuclei_model):
    rng = np.random.default_rng(3)
    x = rng.normal(size=(3, 128, 128, 1)).astype(np.float32)
    m = nuclei_model
    p = unet_oracle.forward(m.weights, m.hp, m.variant, x)
    assert np.allclose(p.sum(-1), 1, atol=1e-6)
    p1 = unet_oracle.forward(m.weights, m.hp, m.variant, x[1:2])
    assert np.allclose(p[1:2], p1, atol=1e-6)     # inference-mode BN: no cross-tile coupling


def test_pi2d_identity_network_returns_input():
    """PI2D.demo (PartitionOfImage.py:125-147): an identity 'network' must reproduce the image."""
    rng = np.random.default_rng(0)
    img = rng.random((150, 211))
    ident = lambda x: np.repeat(x, 2, axis=-1)
    out = pi2d_oracle.infer_image(img, ident, 64, 1, 0.0, 1.0, 8, accum_dtype=np.float64)
    assert np.abs(out[0] - img).max() < 1e-6     # feed is float32


def test_ramp_weight_closed_form():
    for S in (64, 128, 256):
        m = S // 8
        w = pi2d_oracle.ramp_weight(S, m)
        r = np.arange(S)
        d = np.minimum(r, S - 1 - r)
        closed = np.minimum(1.0, np.minimum(d[:, None], d[None, :]) / (2 * m))
        assert np.array_equal(w, closed)


def test_count_positive_on_valid_region():
    g = pi2d_oracle.tile_grid(100, 333, 64, 8)
    cnt = pi2d_oracle.analytic_count(g)
    assert cnt[g.margin:g.margin + 100, g.margin:g.margin + 333].min() >= 0.5
