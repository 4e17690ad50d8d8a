"""-m gpu: the CUDA path, called through the C-ABI, against the oracle.

Tolerances (BASELINE.json north_star): per-tile max|dp| <= 2e-3 vs fp32, K-class argmax
agreement >= 99.9 %, stitched uint8 within +-1 level."""
import numpy as np
import pytest

from oracle import pi2d_oracle, prepost_oracle, unet_oracle
from unmicst_b200 import modelzoo

pytestmark = pytest.mark.gpu

TOL_P = 2e-3
TOL_ARGMAX = 0.999


def _engine(model, **kw):
    from unmicst_b200.engine import Engine
    return Engine(model, device=0, **kw)


def _oracle_fw(model):
    return lambda x: unet_oracle.forward(model.weights, model.hp, model.variant, x)


def _real_tiles(raw, model, n, seed=0):
    img = prepost_oracle.prepare_rescaled(raw)
    S = model.hp["imSize"]
    g = pi2d_oracle.tile_grid(img.shape[0], img.shape[1], S, S // 8)
    frame = pi2d_oracle.pad_frame(img, g)
    rng = np.random.default_rng(seed)
    idx = rng.choice(g.num_tiles, size=min(n, g.num_tiles), replace=False)
    t = np.stack([(pi2d_oracle.cut_tile(frame, g, int(i)) - model.mean) / model.std for i in idx])
    return t[..., None].astype(np.float32)


def _check_probs(got, want):
    assert got.shape == want.shape
    assert np.isfinite(got).all()
    assert np.abs(got - want).max() <= TOL_P
    assert (got.argmax(-1) == want.argmax(-1)).mean() >= TOL_ARGMAX
    assert np.allclose(got.sum(-1), 1.0, atol=1e-5)


def test_forward_legacy_real_weights(sample_raw, nuclei_model):
    tiles = _real_tiles(sample_raw, nuclei_model, 20)
    with _engine(nuclei_model) as e:
        got = e.forward_tiles(tiles)
    _check_probs(got, _oracle_fw(nuclei_model)(tiles))


def test_forward_legacy_k3_two_class_real_weights(sample_raw, cyto_model):
    tiles = _real_tiles(sample_raw, cyto_model, 12, seed=1)
    with _engine(cyto_model) as e:
        got = e.forward_tiles(tiles)
    _check_probs(got, _oracle_fw(cyto_model)(tiles))


@pytest.mark.parametrize("name,n", [("nucleiDAPI1-5", 9), ("nucleiDAPILAMIN", 3), ("CytoplasmIncell2", 2)])
def test_forward_v2_synthetic_weights(name, n):
    m = modelzoo.synthetic_model(name, seed=0)
    rng = np.random.default_rng(7)
    S, C = m.hp["imSize"], m.hp["nChannels"]
    tiles = rng.normal(size=(n, S, S, C)).astype(np.float32)
    with _engine(m) as e:
        got = e.forward_tiles(tiles)
    _check_probs(got, _oracle_fw(m)(tiles))


@pytest.mark.parametrize("precision", ["fp32", "single", "auto"])
@pytest.mark.parametrize("name,n", [("nucleiDAPI1-5", 9), ("nucleiDAPILAMIN", 3), ("CytoplasmIncell2", 2)])
def test_forward_v2_every_precision(name, n, precision):
    """The bench runs `auto` (-> fp16 operands, one MMA per product, when calibration allows); the library default is
    the hi/lo split; `fp32` is the CUDA-core path.  All three must meet the same tolerance against the oracle."""
    m = modelzoo.synthetic_model(name, seed=0)
    rng = np.random.default_rng(17)
    S, C = m.hp["imSize"], m.hp["nChannels"]
    tiles = rng.normal(size=(n, S, S, C)).astype(np.float32)
    with _engine(m, precision=precision) as e:
        got = e.forward_tiles(tiles)
        if precision == "auto":
            assert e.auto_report["chosen"] in ("single", "mixed", "split3")
    _check_probs(got, _oracle_fw(m)(tiles))
    if precision == "fp32":
        assert np.abs(got - _oracle_fw(m)(tiles)).max() <= 2e-5


@pytest.mark.parametrize("precision", ["split3", "single"])
def test_forward_v2_many_tiles_odd_count(precision):
    """Every persistent CTA loops over several work items, rings wrap, the last 4x4-grid box is partly empty."""
    m = modelzoo.synthetic_model("nucleiDAPI1-5", seed=3)
    rng = np.random.default_rng(19)
    tiles = rng.normal(size=(77, 64, 64, 1)).astype(np.float32)
    with _engine(m, precision=precision) as e:
        got = e.forward_tiles(tiles)
        again = e.forward_tiles(tiles[::-1].copy())[::-1]
    _check_probs(got, _oracle_fw(m)(tiles))
    assert np.array_equal(got, again)            # tiles are independent: batch composition does not change a bit


def test_legacy_auto_rejects_plain_single(sample_raw, nuclei_model):
    """Real legacy weights are too steep for one MMA per product everywhere: calibration must not pick 'single', and
    whatever it picks (per-layer 'mixed' or 'split3') must meet the contract against the oracle on real tiles."""
    tiles = _real_tiles(sample_raw, nuclei_model, 12, seed=5)
    with _engine(nuclei_model, precision="auto") as e:
        rep = e.auto_report
        got = e.forward_tiles(tiles)
    assert rep["chosen"] in ("mixed", "split3")
    assert rep["single_vs_split3_max_abs_dp"] > rep["tolerance"]
    if rep["chosen"] == "mixed":
        assert (rep["single_layers"] or rep["partial_layers"]) and rep["mixed_vs_split3_max_abs_dp"] <= rep["tolerance"]
    _check_probs(got, _oracle_fw(nuclei_model)(tiles))


def test_auto_on_steep_weights_and_image_tiles(sample_raw):
    """`auto` as the bench and the CLI run it: calibrated on tiles cut from the image being processed (borders with
    their -mean/std padding included), weights as steep as the real models' (max|logit| ~ 20).
    One MMA per product everywhere is then outside the contract; the error-budgeted per-layer choice must stay inside
    it against the ORACLE, on the calibration tiles and on tiles it has never seen."""
    from unmicst_b200.engine import AUTO_TOLERANCE, PreMap, sample_probe_tiles
    m = modelzoo.synthetic_model("nucleiDAPI1-5", seed=0, logit_gain=22.0)
    pm = PreMap(in_scale=1.0 / 65535)
    tiles = sample_probe_tiles(sample_raw, 64, 1, m.mean, m.std, pm, n=48)
    unseen = sample_probe_tiles(sample_raw, 64, 1, m.mean, m.std, pm, n=40, seed=5)[8:]
    taps = {}
    want = unet_oracle.forward(m.weights, m.hp, m.variant, tiles, taps=taps)
    assert np.abs(taps["logits"]).max() > 12
    with _engine(m, precision="auto", probe_tiles=tiles) as e:
        rep = e.auto_report
        got, got_unseen = e.forward_tiles(tiles), e.forward_tiles(unseen)
        assert e.precision == rep["chosen"]
    assert rep["probe_tiles"] == len(tiles) and rep["budget"] == AUTO_TOLERANCE
    if rep["chosen"] == "mixed":
        assert (rep["single_layers"] or rep["partial_layers"]) and rep["mixed_vs_split3_max_abs_dp"] <= AUTO_TOLERANCE
        assert all(min(l["dp"].values()) >= 0 and l["split_ms"] > 0 for l in rep["layers"])
        assert rep["mixed_vs_split3_max_abs_dp"] <= 1.3 * rep["predicted_quadrature_dp"] + 1e-4      # quadrature model holds
    _check_probs(got, want)
    _check_probs(got_unseen, _oracle_fw(m)(unseen))


def test_per_source_correction_terms_v2():
    """umx_create_ex: which hi/lo correction terms each op adds, per concat source.  Terms 15 everywhere == split3 and
    0 everywhere == single, bit for bit; a handle built with a partial assignment equals the live split handle switched
    to the same terms (umx_set_op_terms, what calibration measures) bit for bit; everything stays within the contract."""
    from unmicst_b200.engine import tensor_ops
    m = modelzoo.synthetic_model("nucleiDAPI1-5", seed=0)
    rng = np.random.default_rng(31)
    tiles = rng.normal(size=(6, 64, 64, 1)).astype(np.float32)
    want = _oracle_fw(m)(tiles)
    ops = [i for i, _ in tensor_ops(m)]
    with _engine(m, precision="split3") as e:
        split = e.forward_tiles(tiles)
    with _engine(m, precision="single") as e:
        single = e.forward_tiles(tiles)
    with _engine(m, precision="mixed", op_terms={i: 15 for i in ops}) as e:
        assert np.array_equal(e.forward_tiles(tiles), split)
    with _engine(m, precision="mixed", op_terms={i: 0 for i in ops}) as e:
        assert np.array_equal(e.forward_tiles(tiles), single)
    partial = {i: t for i, t in zip(ops, [2 | 2 << 2, 0, 0, 0, 0, 3 | 0 << 2, 2 | 2 << 2, 2 | 1 << 2, 1 | 1 << 2, 3 | 0 << 2, 2 | 2 << 2, 3 | 2 << 2])}
    with _engine(m, precision="mixed", op_terms=partial) as e:
        built = e.forward_tiles(tiles)
    with _engine(m, precision="split3") as e:
        for i, t in partial.items():
            e.set_op_terms(i, t)
        live = e.forward_tiles(tiles)
    assert np.array_equal(built, live)
    _check_probs(built, want)
    assert not np.array_equal(built, split) and not np.array_equal(built, single)


def test_wide_correction_mma_of_narrow_layers(monkeypatch):
    """Narrow weight-stationary layers (Cyto2 lu0.conv2, N = 32) multiply a_hi by [w_hi | w_lo] in one MMA of twice the
    width and add the two accumulator halves in the epilogue (umx_op_info: resident == 2).  For every term assignment of
    that layer - including one whose first source has no a_hi*w_lo term, where the odd CTA supplies zeros - the result
    equals the three-MMA form up to the order of the fp32 additions."""
    from unmicst_b200.engine import tensor_ops
    m = modelzoo.synthetic_model("CytoplasmIncell2", seed=0)
    rng = np.random.default_rng(5)
    tiles = rng.normal(size=(2, 256, 256, 1)).astype(np.float32)
    want = _oracle_fw(m)(tiles)
    op = [i for i, n in tensor_ops(m) if n == "lu0.conv2"][0]
    cases = [15, 3, 1 | 1 << 2, 2 | 1 << 2, 2 | 3 << 2, 0]
    got = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("UMX_TC_NCAT", flag)
        with _engine(m, precision="split3") as e:
            assert e.op_info(op)["resident"] == (2 if flag == "1" else 1)
            for t in cases:
                e.set_op_terms(op, t)
                got[flag, t] = e.forward_tiles(tiles)
    for t in cases:
        assert np.abs(got["1", t] - got["0", t]).max() < 1e-5, t
    _check_probs(got["1", 15], want)
    assert not np.array_equal(got["1", 0], got["1", 15])            # the terms are really switched


def test_mixed_precision_masks_v2():
    """Any subset of layers may run with one MMA per product: all-zero mask == split3 bit for bit, a partial mask and
    the full mask stay within the contract."""
    m = modelzoo.synthetic_model("nucleiDAPI1-5", seed=0)
    rng = np.random.default_rng(29)
    tiles = rng.normal(size=(5, 64, 64, 1)).astype(np.float32)
    want = _oracle_fw(m)(tiles)
    with _engine(m, precision="split3") as e:
        ref = e.forward_tiles(tiles)
    with _engine(m, precision="mixed", single_mask=0) as e:
        assert np.array_equal(e.forward_tiles(tiles), ref)
    for mask in (0b0101010101010100, (1 << 40) - 1):
        with _engine(m, precision="mixed", single_mask=mask) as e:
            _check_probs(e.forward_tiles(tiles), want)


def test_image_single_precision_v2():
    m = modelzoo.synthetic_model("nucleiDAPI1-5", seed=0)
    rng = np.random.default_rng(23)
    img = rng.random((150, 333))
    with _engine(m, precision="single") as e:
        u8, f32 = e.infer_image(img, want_f32=True)
    want = pi2d_oracle.infer_image(img, _oracle_fw(m), 64, 1, m.mean, m.std, 32, accum_dtype=np.float64)
    assert np.abs(f32 - want).max() <= TOL_P
    assert np.abs(u8.astype(int) - np.uint8(255 * want).astype(int)).max() <= 1


def test_forward_v2_steep_softmax():
    """Stress variant: logits scaled until the softmax is as steep as the real models'."""
    m = modelzoo.synthetic_model("nucleiDAPI1-5", seed=1, logit_gain=12.0)
    rng = np.random.default_rng(8)
    tiles = rng.normal(size=(4, 64, 64, 1)).astype(np.float32)
    with _engine(m) as e:
        got = e.forward_tiles(tiles)
    _check_probs(got, _oracle_fw(m)(tiles))


def test_forward_batch_composition_and_empty(nuclei_model):
    rng = np.random.default_rng(2)
    tiles = rng.normal(size=(5, 128, 128, 1)).astype(np.float32)
    with _engine(nuclei_model, max_batch_tiles=2) as e:      # forces several launch groups
        a = e.forward_tiles(tiles)
        b = np.concatenate([e.forward_tiles(tiles[i:i + 1]) for i in range(5)])
        z = e.forward_tiles(tiles[:0])
    assert np.array_equal(a, b)
    assert z.shape == (0, 128, 128, 3)


def test_image_legacy_matches_goldens_and_oracle(sample_raw, sample_goldens, nuclei_model):
    m = nuclei_model
    img = prepost_oracle.prepare_rescaled(sample_raw)
    with _engine(m) as e:
        u8, f32 = e.infer_image(img, want_u8=True, want_f32=True)
    for cls, key in ((1, "contours"), (2, "nuclei")):
        d = np.abs(u8[cls].astype(int) - sample_goldens[key].astype(int))
        assert d.max() <= 1
    want = pi2d_oracle.infer_image(img, _oracle_fw(m), 128, 1, m.mean, m.std, 16, accum_dtype=np.float64)
    assert np.abs(f32 - want).max() <= TOL_P
    assert np.abs(u8.astype(int) - np.uint8(255 * want).astype(int)).max() <= 1


def test_image_premap_on_device_equals_host_prepared(sample_raw, nuclei_model):
    """u16 samples + in-kernel img_as_float/rescale_intensity == float64 image prepared on the host."""
    from unmicst_b200.engine import PreMap
    img = prepost_oracle.prepare_rescaled(sample_raw)
    f = sample_raw.astype(np.float64) * (1.0 / 65535)
    pm = PreMap(in_scale=1.0 / 65535, rescale=True, imin=float(f.min()), imax=float(f.max()), omin=0.0, omax=0.983)
    with _engine(nuclei_model) as e:
        a, _ = e.infer_image(img)
        b, _ = e.infer_image(sample_raw, premap=pm)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("shape", [(1, 1), (37, 205), (96, 96), (97, 193), (300, 64)])
def test_image_ragged_sizes_v2(shape):
    m = modelzoo.synthetic_model("nucleiDAPI1-5", seed=0)
    rng = np.random.default_rng(11)
    img = rng.random(shape)
    with _engine(m, max_batch_tiles=7) as e:
        u8, f32 = e.infer_image(img, want_f32=True)
    want = pi2d_oracle.infer_image(img, _oracle_fw(m), 64, 1, m.mean, m.std, 32, accum_dtype=np.float64)
    assert f32.shape == (3,) + shape
    assert np.abs(f32 - want).max() <= TOL_P
    assert np.abs(u8.astype(int) - np.uint8(255 * want).astype(int)).max() <= 1


def test_image_two_channel_duo():
    m = modelzoo.synthetic_model("nucleiDAPILAMIN", seed=0)
    rng = np.random.default_rng(12)
    img = rng.random((2, 150, 260))
    with _engine(m) as e:
        _, f32 = e.infer_image(img, want_u8=False, want_f32=True)
    want = pi2d_oracle.infer_image(img, _oracle_fw(m), 128, 2, m.mean, m.std, 24, accum_dtype=np.float64)
    assert np.abs(f32 - want).max() <= TOL_P


def test_tile_row_bands_reassemble_bit_exactly(sample_raw, nuclei_model):
    from unmicst_b200.engine import split_tile_rows, tile_geometry
    img = prepost_oracle.prepare_rescaled(sample_raw)
    H, W = img.shape
    _, _, npr, _ = tile_geometry(H, W, 128)
    with _engine(nuclei_model, max_batch_tiles=25) as e:
        whole, _ = e.infer_image(img)
        for parts in (2, 3, npr):
            out = np.zeros_like(whole)
            covered = np.zeros(H, dtype=int)
            for band in split_tile_rows(npr, parts):
                e.infer_image(img, tile_rows=band, out_u8=out)
                r0, r1 = e.band_rows(H, band)
                covered[r0:r1] += 1
            assert (covered == 1).all()
            assert np.array_equal(out, whole)


def test_errors_are_reported_not_fatal(nuclei_model):
    from unmicst_b200._lib import EngineError, UMX_ENOTENSOR
    from unmicst_b200.engine import Engine
    import copy
    bad = copy.copy(nuclei_model)
    bad.weights = {k: v for k, v in nuclei_model.weights.items() if k != "lb/kernel1"}
    with pytest.raises(EngineError) as ei:
        Engine(bad)
    assert ei.value.code == UMX_ENOTENSOR and "lb/kernel1" in str(ei.value)
    with _engine(nuclei_model) as e:
        with pytest.raises(ValueError):
            e.forward_tiles(np.zeros((1, 64, 64, 1), np.float32))
        with pytest.raises(EngineError):
            e.infer_image(np.zeros((3, 50, 50)))          # 3 planes into a 1-channel network


def test_profile_and_launch_count(nuclei_model):
    rng = np.random.default_rng(5)
    with _engine(nuclei_model) as e:
        e.profile_enable(True)
        n0 = e.launch_count
        e.infer_image(rng.random((200, 200)))
        prof = e.profile_read()
        assert e.launch_count > n0
    names = [p["name"] for p in prof]
    assert "lu1.conv2" in names and "stitch_quantize" in names
    assert all(p["ms"] > 0 for p in prof if p["launches"] > 0)
