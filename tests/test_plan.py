"""Plan construction (graph -> fused ops -> kernel families) through the host-only umx_describe_plan: no GPU.

What is checked is the reference's graph structure (UnMicst1-5.py:55-237, UnMicst.py:51-187) as the library
lowers it: op order, fusions, the raw-input rewrite, per-op precision, and the FLOP totals SURVEY.md §8d quotes."""
import re

import pytest

from unmicst_b200 import _lib, modelzoo
from unmicst_b200.engine import describe_plan


def _parse(lines):
    ops = []
    for l in lines:
        idx, kind, name, rest = l.split(" ", 3)
        d = {"idx": int(idx), "kind": kind, "name": name, "rest": rest}
        if kind == "conv":
            d["family"] = rest.split()[0]
            d["mode"] = int(re.search(r"mode=(\d+)", rest).group(1))
            d["planes"] = int(re.search(r"planes=(\d+)", rest).group(1))
            d["flops"] = float(re.search(r"flops=(\d+)", rest).group(1))
            d["terms"] = re.findall(r"\[k=(\d+) ([^\]]+)\]", rest)
            d["flags"] = set(rest.split(" out=")[0].split()[3:])
        ops.append(d)
    return ops


def test_solo_plan_structure_and_flops():
    m = modelzoo.synthetic_model("nucleiDAPI1-5", seed=0)
    ops = _parse(describe_plan(m))
    names = [o["name"] for o in ops]
    assert names == ["ld0.conv0", "ld1.conv0", "ld2.conv0", "ld3.conv0", "lb.conv", "lu3.convT", "lu3.conv2", "lu2.convT",
                     "lu2.conv2", "lu1.convT", "lu1.conv2", "lu0.convT", "input.taps3", "lu0.conv2", "lt.softmax"]
    convs = [o for o in ops if o["kind"] == "conv"]
    assert convs[0]["family"] == "first" and all(o["family"] == "tensor" for o in convs[1:])
    # v2 down layers: kernelD + shortcut merged into ONE 3x3 term, BN folded into weights + bias, pool fused
    assert convs[1]["terms"] == [("3", "ld0.conv0:80")] and {"pool", "bias"} <= convs[1]["flags"]
    # up path: concat [skip | up] as two sources of one term, never materialised
    assert convs[6]["name"] == "lu3.conv2" and convs[6]["terms"] == [("3", "ld2.conv0:320|lu3.convT:640")]
    # lu0.conv2 = conv3x3(up) + 1x1(tap-expanded raw input) joined at the centre tap
    lu0 = next(o for o in ops if o["name"] == "lu0.conv2")
    assert lu0["mode"] == 4 and lu0["terms"] == [("3", "lu0.convT:80"), ("1", "input.taps3:9")]
    assert all("transpose" in o["flags"] for o in convs if o["name"].endswith("convT"))
    total = sum(o["flops"] for o in convs) + 2 * 80 * 3 * 64 * 64
    assert abs(total - 4.4964e9) / 4.4964e9 < 1e-3           # 4.4964 GFLOP per tile


def test_fp32_plan_has_no_tensor_ops_and_no_rewrite():
    m = modelzoo.synthetic_model("nucleiDAPI1-5", seed=0)
    ops = _parse(describe_plan(m, "fp32"))
    assert not any(o["kind"] == "taps" for o in ops)
    assert all(o["family"] in ("first", "simt") for o in ops if o["kind"] == "conv")
    lu0 = next(o for o in ops if o["name"] == "lu0.conv2")
    assert lu0["terms"] == [("3", "input:1|lu0.convT:80")]   # the reference's concat order (UnMicst1-5.py:196)


@pytest.mark.parametrize("name,taps_c,k_classes,per_tile_gflop", [("nucleiDAPILAMIN", 18, 3, 4.6325), ("CytoplasmIncell2", 9, 2, 7.5104)])
def test_other_v2_models(name, taps_c, k_classes, per_tile_gflop):
    m = modelzoo.synthetic_model(name, seed=0)
    ops = _parse(describe_plan(m, "single"))
    taps = [o for o in ops if o["kind"] == "taps"]
    assert len(taps) == 1 and f"x{taps_c}" in taps[0]["rest"]
    assert ops[-1]["kind"] == "top" and f"k={k_classes}" in ops[-1]["rest"]
    convs = [o for o in ops if o["kind"] == "conv"]
    assert all(o["planes"] == 1 for o in convs if o["family"] == "tensor")
    S, n0 = m.hp["imSize"], m.hp["nOut0"]
    total = sum(o["flops"] for o in convs) + 2 * n0 * k_classes * S * S
    assert abs(total - per_tile_gflop * 1e9) / (per_tile_gflop * 1e9) < 1e-3


def test_legacy_plan(nuclei_model):
    ops = _parse(describe_plan(nuclei_model))
    by = {o["name"]: o for o in ops}
    # extra conv chain: conv0 alone, conv1 = 5x5 of conv0 + 1x1 shortcut of the layer input, BN after the ReLU (post)
    assert by["ld0.conv1"]["mode"] == 3 and by["ld0.conv1"]["terms"] == [("5", "ld0.conv0:16"), ("1", "input:1")]
    assert by["ld1.conv1"]["mode"] == 4 and {"pool", "post"} <= by["ld1.conv1"]["flags"]
    assert by["lu0.conv2"]["terms"] == [("5", "lu0.convT:16"), ("1", "input.taps5:25")]
    assert ops[-1]["rest"].startswith("src=lu0.extra0")
    total = sum(o["flops"] for o in ops if o["kind"] == "conv") + 2 * 16 * 3 * 128 * 128
    assert abs(total - 1.8151e9) / 1.8151e9 < 1e-3           # legacy nucleiDAPI 1.8151 GFLOP per tile


def test_mixed_mask_selects_planes_per_op():
    m = modelzoo.synthetic_model("nucleiDAPI1-5", seed=0)
    base = _parse(describe_plan(m, "split3"))
    tensor_idx = [o["idx"] for o in base if o["kind"] == "conv" and o["family"] == "tensor"]
    mask = sum(1 << i for i in tensor_idx[::2])
    ops = _parse(describe_plan(m, "mixed", mask))
    for o in ops:
        if o["kind"] == "conv" and o["family"] == "tensor":
            assert o["planes"] == (1 if (mask >> o["idx"]) & 1 else 2)


def test_bad_models_are_errors_not_crashes():
    m = modelzoo.synthetic_model("nucleiDAPI1-5", seed=0)
    bad = modelzoo.synthetic_model("nucleiDAPI1-5", seed=0)
    bad.hp = dict(m.hp, imSize=96)                            # not a power of two
    with pytest.raises(_lib.EngineError):
        describe_plan(bad)
    missing = modelzoo.synthetic_model("nucleiDAPI1-5", seed=0)
    missing.weights = {k: v for k, v in m.weights.items() if k != "lb/kernel1"}
    with pytest.raises(_lib.EngineError, match="lb/kernel1"):
        describe_plan(missing)
