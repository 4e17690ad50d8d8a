"""TF-bundle reader, hp pickles, graph-variant detection and shape derivation against
the real on-disk files of every reference model folder (fixtures in tests/golden/models)."""
import os

import numpy as np
import pytest

from unmicst_b200 import modelzoo, tfbundle

ALL = ["nucleiDAPI", "nucleiDAPI1-5", "nucleiDAPILAMIN", "CytoplasmIncell2", "CytoplasmIncell",
       "CytoplasmZeissNikon", "mousenucleiDAPI"]


@pytest.mark.parametrize("name", ALL)
def test_index_matches_derived_shapes(golden_dir, name):
    d = os.path.join(golden_dir, "models", name)
    entries = tfbundle.read_index(os.path.join(d, "model.ckpt.index"))
    hp = modelzoo.load_pickle(os.path.join(d, "hp.data"))
    assert hp == modelzoo.KNOWN_HP[name]
    variant = modelzoo.detect_variant(entries.keys())
    assert variant == modelzoo.KNOWN_VARIANT[name]
    want = modelzoo.expected_tensors(hp, variant)
    have = {n: e.shape for n, e in entries.items() if not tfbundle.is_optimizer_slot(n)}
    assert have == want
    mean = modelzoo.load_pickle(os.path.join(d, "datasetMean.data"))
    std = modelzoo.load_pickle(os.path.join(d, "datasetStDev.data"))
    assert (mean, std) == modelzoo.KNOWN_NORM[name]


@pytest.mark.parametrize("name", ["nucleiDAPI", "CytoplasmIncell"])
def test_data_shard_is_tiled_exactly_by_entries(golden_dir, name):
    prefix = os.path.join(golden_dir, "models", name, "model.ckpt")
    entries = sorted(tfbundle.read_index(prefix + ".index").values(), key=lambda e: e.offset)
    pos = 0
    for e in entries:
        assert e.offset == pos
        pos += e.size
    assert pos == os.path.getsize(tfbundle.data_path(prefix))


def test_load_model_real_weights(nuclei_model):
    m = nuclei_model
    assert m.variant == "legacy" and not m.synthetic
    assert m.weights["downsampling/ld0/kernel1"].shape == (5, 5, 1, 16)
    assert all(v.dtype == np.float32 for v in m.weights.values())
    assert not any(tfbundle.is_optimizer_slot(n) for n in m.weights)
    assert sum(v.size for v in m.weights.values()) == 238368


def test_missing_shard_raises_unless_synthetic(golden_dir):
    d = os.path.join(golden_dir, "models", "nucleiDAPI1-5")
    with pytest.raises(FileNotFoundError):
        modelzoo.load_model(d)
    m = modelzoo.load_model(d, allow_synthetic=True)
    assert m.synthetic and m.variant == "v2"
    assert m.weights["lu3/kernel2"].shape == (3, 3, 960, 640)
    assert sum(v.size for v in m.weights.values()) == 29335532


def test_synthetic_weights_are_deterministic():
    a = modelzoo.synthetic_model("CytoplasmIncell2", seed=0)
    b = modelzoo.synthetic_model("CytoplasmIncell2", seed=0)
    c = modelzoo.synthetic_model("CytoplasmIncell2", seed=1)
    assert all(np.array_equal(a.weights[k], b.weights[k]) for k in a.weights)
    assert any(not np.array_equal(a.weights[k], c.weights[k]) for k in a.weights)


def test_hp_mismatch_is_reported(golden_dir):
    d = os.path.join(golden_dir, "models", "nucleiDAPI")
    with pytest.raises(ValueError, match="hp.data does not describe"):
        modelzoo.load_model(d, hp_override={"nOut0": 20})


def test_bad_magic(tmp_path):
    p = tmp_path / "x.index"
    p.write_bytes(b"\0" * 64)
    with pytest.raises(tfbundle.BundleError):
        tfbundle.read_index(str(p))
