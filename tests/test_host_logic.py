"""Host-side logic that needs no GPU: tile geometry/band sharding, CLI flag mapping and naming,
pre/post-processing against the oracle restatement, TIFF round trips."""
import os

import numpy as np
import pytest

from oracle import pi2d_oracle, prepost_oracle
from unmicst_b200 import cli, prepost, tiffio
from unmicst_b200.engine import split_tile_rows, tile_geometry


@pytest.mark.parametrize("H,W,S", [(832, 960, 128), (20000, 20000, 64), (1, 1, 64), (4096, 4096, 128), (18432, 30720, 256)])
def test_tile_geometry_matches_pi2d(H, W, S):
    m, sub, npr, npc = tile_geometry(H, W, S)
    g = pi2d_oracle.tile_grid(H, W, S, int(S / 8))
    assert (m, sub, npr, npc) == (g.margin, g.sub, g.npr, g.npc)


def test_reference_tile_counts():
    """90 tiles @128 / 360 @64 for 832x960; 173 889 @64 for 20k^2; 1 849 @128 for 4096^2."""
    cnt = lambda H, W, S: (lambda t: t[2] * t[3])(tile_geometry(H, W, S))
    assert cnt(832, 960, 128) == 90 and cnt(832, 960, 64) == 360
    assert cnt(20000, 20000, 64) == 173889 and cnt(4096, 4096, 128) == 1849 and cnt(40000, 40000, 64) == 695556


@pytest.mark.parametrize("npr,parts", [(417, 8), (9, 2), (3, 8), (1, 4), (100, 1), (834, 8), (8, 8), (10, 8)])
def test_band_split_is_a_partition(npr, parts):
    for balance in (False, True):
        bands = split_tile_rows(npr, parts, balance_seams=balance)
        assert bands[0][0] == 0 and bands[-1][1] == npr
        assert all(a[1] == b[0] for a, b in zip(bands, bands[1:]))
        sizes = [b - a for a, b in bands]
        assert min(sizes) >= 1
        if balance:     # what a GPU computes = its rows + the recomputed seam row (all bands but the first)
            comp = [n + (1 if i else 0) for i, n in enumerate(sizes)]
            assert max(comp) - min(comp) <= 1
        else:
            assert max(sizes) - min(sizes) <= 1


def test_band_split_417_rows_on_8_gpus_computes_53_each():
    """20k x 20k solo = 417 tile rows; balanced with the seam row every GPU computes 53 (not 54)."""
    bands = split_tile_rows(417, 8)
    assert [b - a + (1 if i else 0) for i, (a, b) in enumerate(bands)] == [53] * 8


def test_wrapper_converts_to_zero_based():
    a = cli.wrapper_parser().parse_args(["img.ome.tif", "--channel", "3", "--classOrder", "1", "2", "3", "--GPU", "2",
                                         "--stackOutput", "--outputPath", "out"])
    tool, argv = cli.wrapper_to_tool_argv(a)
    assert tool == "unmicst-solo"
    t = cli.tool_parser(cli.TOOLS[tool]).parse_args(argv)
    assert t.channel == ["2"] and t.classOrder == [0, 1, 2] and t.GPU == 1 and t.stackOutput and t.outputPath == "out"
    assert t.model == "nucleiDAPI1-5" and t.mean == -1 and t.scalingFactor == 1


def test_wrapper_tool_dispatch_and_defaults():
    for tool, model in (("unmicst-duo", "nucleiDAPILAMIN"), ("unmicst-legacy", "nucleiDAPI"), ("UnMicstCyto2", "CytoplasmIncell2")):
        a = cli.wrapper_parser().parse_args(["x.tif", "--tool", tool, "--channel", "1", "2"])
        got, argv = cli.wrapper_to_tool_argv(a)
        assert got == tool
        t = cli.tool_parser(cli.TOOLS[tool]).parse_args(argv)
        assert t.model == model and t.GPU == -1
        if tool == "unmicst-duo":
            assert t.channel == ["0", "1"]
    a = cli.wrapper_parser().parse_args(["x.tif"])
    assert "--outputPath" not in cli.wrapper_to_tool_argv(a)[1]       # quirk Q7 fixed: no literal 'None'


def test_file_stem_rules():
    assert cli.split_name("exemplar-001-cycle6.ome.tif", True) == ("exemplar-001-cycle6", "ome.tif")
    assert cli.split_name("a.b.tif", True) == ("a.b", "tif")
    assert cli.split_name("a.b.tif", False) == ("a", "b.tif")          # legacy/duo/Cyto2 split at the first dot
    with pytest.raises(NotImplementedError):
        cli.split_name("noext", True)
    with pytest.raises(NotImplementedError):
        cli.read_channel("x.xyz", "xyz", 0)


def _apply(arr, pm):
    x = arr.astype(np.float64) * pm.in_scale
    if pm.rescale:
        x = np.clip(x, pm.imin, pm.imax)
        x = (x - pm.imin) / (pm.imax - pm.imin) * (pm.omax - pm.omin) + pm.omin
    return x


def test_network_input_matches_oracle_preprocessing(sample_raw):
    arr, pm, shape = prepost.network_input(sample_raw, 1.0, stretch=True)
    assert arr.dtype == np.uint16 and shape is None                    # integers go to the GPU untouched
    assert np.array_equal(_apply(arr, pm), prepost_oracle.prepare_rescaled(sample_raw))
    arr, pm, shape = prepost.network_input(sample_raw, 1.0, stretch=False)
    assert np.array_equal(_apply(arr, pm), prepost_oracle.prepare_solo(sample_raw))
    crop = sample_raw[:200, :300]
    for f in (2.0, 0.5):
        # the samples stay raw; the resize happens on the GPU (tests/test_gpu_resize.py); the stretch range is the
        # resized image's (UnMicst.py:627-631), here through the host path (no engine): percentile and plain max
        for outlier in (99.0, -1):
            arr, pm, shape = prepost.network_input(crop, f, stretch=True, outlier=outlier)
            assert arr is crop or np.array_equal(arr, crop)
            assert shape == (int(200 * f), int(300 * f))
            resized = prepost_oracle.resize(crop, shape)
            top = np.max(resized) if outlier == -1 else np.percentile(resized, outlier)
            assert abs(pm.imin - resized.min()) <= 1e-15 and abs(pm.imax - top) <= 1e-15 and pm.rescale
            want = prepost_oracle.prepare_rescaled(crop, f, outlier)
            x = np.clip(resized, pm.imin, pm.imax)
            assert np.allclose((x - pm.imin) / (pm.imax - pm.imin) * 0.983, want, atol=1e-12)


def test_probe_tiles_equal_the_oracle_tiler(sample_raw):
    """engine.sample_probe_tiles (calibration / bench parity probes) cuts the same network inputs as the reference's
    tile loop (PartitionOfImage.py:49-82 + UnMicst1-5.py:700), also through a --scalingFactor resize."""
    from unmicst_b200.engine import PreMap, probe_tile_indices, sample_probe_tiles, tile_geometry
    crop = sample_raw[:150, :333]
    for f in (1.0, 2.0):
        cells = prepost_oracle.prepare_solo(crop, f)
        g = pi2d_oracle.tile_grid(cells.shape[0], cells.shape[1], 64, 8)
        frame = pi2d_oracle.pad_frame(cells, g)
        _, _, npr, npc = tile_geometry(cells.shape[0], cells.shape[1], 64)
        idx = probe_tile_indices(npr, npc, 12)
        assert len(set(idx)) == len(idx) == min(12, npr * npc) and 0 in idx and npr * npc - 1 in idx
        got = sample_probe_tiles(crop, 64, 1, 0.34, 0.25, PreMap(in_scale=1.0 / 65535), indices=idx,
                                 infer_shape=None if f == 1.0 else cells.shape)
        want = np.stack([(pi2d_oracle.cut_tile(frame, g, i) - 0.34) / 0.25 for i in idx])[..., None].astype(np.float32)
        assert np.abs(got - want).max() <= (0 if f == 1.0 else 1e-6)


def test_precision_knapsack_over_correction_terms():
    """engine.choose_op_terms: one option per op (terms, error alone, time), fastest assignment whose errors fit the budget
    in quadrature; engine.tensor_op_sources: the K each concat source contributes (weights of its correction MMAs)."""
    from unmicst_b200 import modelzoo
    from unmicst_b200.engine import choose_op_terms, tensor_op_sources
    opts = {1: [(0, 1e-3, 1.0), (10, 5e-4, 1.5), (15, 0.0, 3.0)], 2: [(0, 2e-4, 1.0), (15, 0.0, 3.0)],
            3: [(0, 9e-4, 2.0), (2, 3e-4, 3.0), (15, 0.0, 6.0)]}
    assert choose_op_terms(opts, 1e-3) == {1: 10, 2: 0, 3: 2}          # 5e-4, 2e-4, 3e-4 -> 6.2e-4, time 5.5
    assert choose_op_terms(opts, 1e-5) == {1: 15, 2: 15, 3: 15}
    assert choose_op_terms(opts, 1.0) == {1: 0, 2: 0, 3: 0}
    assert choose_op_terms(opts, 6e-4) == {1: 15, 2: 0, 3: 2}
    with pytest.raises(ValueError):
        choose_op_terms({1: [(0, 1e-3, 1.0)]}, 1e-4)                    # no zero-error option, nothing fits
    src = {n: ks for _, n, ks in tensor_op_sources(modelzoo.synthetic_model("nucleiDAPI1-5"))}
    assert src["lu1.conv2"] == [9 * 80, 9 * 160] and src["lu0.conv2"] == [9 * 80, 9] and src["ld1.conv0"] == [9 * 80]
    legacy = {n: ks for _, n, ks in tensor_op_sources(modelzoo.synthetic_model("nucleiDAPI"))}
    assert legacy["ld0.conv1"] == [25 * 16]                            # its one-channel shortcut lives in the epilogue, not in K


def test_error_budgeted_layer_selection():
    """engine.choose_single_mask: the subset with the largest saving whose errors, added in quadrature, fit the budget."""
    from unmicst_b200.engine import choose_single_mask
    errs = {1: 1e-4, 2: 5e-4, 3: 9e-4, 5: 2e-3}
    costs = {1: 1.0, 2: 2.0, 3: 5.0, 5: 50.0}
    assert choose_single_mask(errs, costs, 1e-3) == (1 << 1) | (1 << 3)         # {2,3} = 1.03e-3 is over; 5 never fits
    assert choose_single_mask(errs, costs, 1e-5) == 0
    assert choose_single_mask(errs, costs, 1.0) == 0b101110
    assert choose_single_mask(errs, {1: 1.0, 2: 0.0, 3: 5.0, 5: 50.0}, 1.0) == 0b101010   # no saving, no bit


def test_requantisation_table_and_preview(sample_raw, sample_goldens):
    v = np.arange(256, dtype=np.uint8).reshape(16, 16)
    want = np.uint8(255 * prepost_oracle.resize(v, (16, 16)))
    assert np.array_equal(prepost.back_to_raw_size(v, (16, 16)), want)
    assert np.abs(want.astype(int) - v.astype(int)).max() <= 1
    assert np.array_equal(prepost.preview_page(sample_raw), sample_goldens["raw"])


def test_streamed_pages_fill_in_any_band_order(tmp_path):
    """The K class maps leave the GPU band by band: BigTiffWriter.begin_pages reserves every page, rows land in place."""
    rng = np.random.default_rng(3)
    pages = rng.integers(0, 255, (3, 100, 77), dtype=np.uint8)
    p = str(tmp_path / "bands.tif")
    w = tiffio.BigTiffWriter(p)
    w.begin_pages(3, 100, 77)
    for r0, r1 in ((40, 100), (0, 17), (17, 40)):
        for k in range(3):
            w.write_page_rows(k, r0, pages[k, r0:r1])
    w.end_pages()
    w.write_page(pages[0][:50])                 # the preview page follows
    w.close()
    assert tiffio.count_pages(p) == 4
    assert all(np.array_equal(tiffio.read_page(p, k), pages[k]) for k in range(3))
    assert np.array_equal(tiffio.read_page(p, 3), pages[0][:50])
    w = tiffio.BigTiffWriter(str(tmp_path / "short.tif"))
    w.begin_pages(1, 10, 5)
    w.write_page_rows(0, 0, np.zeros((4, 5), np.uint8))
    with pytest.raises(tiffio.TiffError):
        w.end_pages()                           # 6 rows never arrived


def test_tiff_roundtrip_append_and_golden_pages(tmp_path, golden_dir, sample_goldens):
    rng = np.random.default_rng(0)
    pages = [rng.integers(0, 255, (77, 131), dtype=np.uint8) for _ in range(3)]
    for big in (True, False):
        p = str(tmp_path / f"s{big}.tif")
        for i, pg in enumerate(pages):
            tiffio.imsave(p, pg, append=i > 0, bigtiff=big)
        assert tiffio.count_pages(p) == 3
        assert all(np.array_equal(tiffio.read_page(p, i), pages[i]) for i in range(3))
        from PIL import Image
        im = Image.open(p)
        assert im.n_frames == 3
    g = os.path.join(golden_dir, "sample", "105_ContoursPM_1.tif")
    assert tiffio.count_pages(g) == 2
    assert np.array_equal(tiffio.read_page(g, 0), sample_goldens["contours"])
    assert np.array_equal(tiffio.read_page(g, 1), sample_goldens["raw"])
    with pytest.raises(IndexError):
        tiffio.read_page(g, 2)
    x16 = rng.integers(0, 65535, (40, 50), dtype=np.uint16)
    p = str(tmp_path / "u16.tif")
    with tiffio.BigTiffWriter(p) as w:                                   # streamed page, band by band
        w.begin_page(40, 50, np.uint16)
        w.write_rows(x16[:13]); w.write_rows(x16[13:])
        w.end_page()
    assert np.array_equal(tiffio.read_page(p), x16)


def test_product_never_imports_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for d, _, files in os.walk(os.path.join(root, "unmicst_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(d, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_mixed_precision_mask_packing():
    """UMX_PREC_MIXED carries its 64-bit op mask in two int32 fields: bit patterns must survive the signed ctypes fields
    (the C side reassembles them as (uint32)lo | (uint64)(uint32)hi << 32, umx_api.cu umx_create)."""
    import ctypes as C
    from unmicst_b200.engine import mask_to_reserved
    from unmicst_b200 import _lib
    for mask in (0, 1, 0b1010, (1 << 31), (1 << 32) - 1, (1 << 32), (1 << 63) | 5, 2 ** 64 - 1, 2 ** 64 + 3):
        lo, hi = mask_to_reserved(mask)
        assert -(1 << 31) <= lo < (1 << 31) and -(1 << 31) <= hi < (1 << 31)
        d = _lib.umx_model_desc()
        d.reserved[0], d.reserved[1] = lo, hi
        back = (d.reserved[0] & 0xFFFFFFFF) | ((d.reserved[1] & 0xFFFFFFFF) << 32)
        assert back == mask & (2 ** 64 - 1)
    assert _lib.PRECISIONS["mixed"] == 4
    assert C.sizeof(_lib.umx_model_desc) == 16 * 4
