"""Python handle over the C-ABI: one Engine = one model resident on one B200.

Engine.forward_tiles   <-> Session.run(UNet2D.nn, {tfData, tfTraining:0})   UnMicst1-5.py:704
Engine.infer_image     <-> UNet2D.singleImageInference for all classes        UnMicst1-5.py:687-710
                           (+ PI2D, toolbox/PartitionOfImage.py:23-122, + uint8 quantise :848)
MultiEngine            tile-row bands over several GPUs, one host thread per GPU,
                       no collective
"""
from __future__ import annotations

import ctypes as C
import threading
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import check, lib
from .modelzoo import LEGACY, V2, Model

_NP2UMX = {np.dtype(np.uint8): _lib.UMX_U8, np.dtype(np.uint16): _lib.UMX_U16,
           np.dtype(np.float32): _lib.UMX_F32, np.dtype(np.float64): _lib.UMX_F64}


@dataclass
class PreMap:
    """img_as_float scaling and the optional rescale_intensity stretch of the CLI
    scripts (UnMicst1-5.py:813-821), applied on the GPU in float64."""
    in_scale: float = 1.0
    rescale: bool = False
    imin: float = 0.0
    imax: float = 1.0
    omin: float = 0.0
    omax: float = 0.983

    def to_c(self) -> _lib.umx_premap:
        return _lib.umx_premap(self.in_scale, int(self.rescale), 0, self.imin, self.imax, self.omin, self.omax)


def device_count() -> int:
    return int(lib().umx_device_count())


def pick_gpu_most_free() -> int:
    """toolbox/GPUselect.py:4-22 — index of the GPU with the most free memory."""
    best, best_free = 0, -1
    for d in range(device_count()):
        free, total = C.c_int64(0), C.c_int64(0)
        check(lib().umx_device_free_mem(d, C.byref(free), C.byref(total)))
        if free.value > best_free:
            best, best_free = d, free.value
    return best


def tile_geometry(H: int, W: int, S: int) -> Tuple[int, int, int, int]:
    """(margin, sub, npr, npc) of PI2D.setup (PartitionOfImage.py:25-50) with margin=int(S/8)."""
    m = int(S / 8)
    sub = S - 2 * m
    return m, sub, -(-H // sub), -(-W // sub)


def split_tile_rows(npr: int, parts: int, balance_seams: bool = True) -> List[Tuple[int, int]]:
    """Contiguous tile-row bands; empty bands dropped.  Every band but the first also recomputes
    the tile row above its seam, so with ``balance_seams`` the bands are sized to equalise the tile rows each GPU
    *computes* (own rows + 1 for bands 1..), not the rows it owns."""
    parts = max(1, min(parts, npr))
    if not balance_seams or parts == 1:
        base, extra = divmod(npr, parts)
        sizes = [base + (1 if p < extra else 0) for p in range(parts)]
    else:
        base, extra = divmod(npr + parts - 1, parts)           # computed rows per band
        comp = [base + (1 if p >= parts - extra else 0) for p in range(parts)]
        sizes = [comp[0]] + [c - 1 for c in comp[1:]]
        if min(sizes) < 1:                                     # too few rows to pay for seams: plain split
            return split_tile_rows(npr, parts, balance_seams=False)
    out, lo = [], 0
    for n in sizes:
        out.append((lo, lo + n))
        lo += n
    assert lo == npr
    return out


def band_rows_py(H: int, S: int, tile_rows: Tuple[int, int]) -> Tuple[int, int]:
    """Image rows [r0, r1) owned by the tile-row band [a, b) — pure-Python twin of umx_band_rows:
    the band emits padded-frame rows [a*sub, b*sub) (to the frame end for the last band)."""
    m, sub, npr, _ = tile_geometry(H, 1, S)
    a, b = tile_rows
    b = npr if b <= 0 or b > npr else b
    p0 = a * sub
    p1 = npr * sub + 2 * m if b == npr else b * sub
    return min(H, max(0, p0 - m)), min(H, max(0, p1 - m))


def _resize_window(y: int, src: int, dst: int) -> Tuple[int, int]:
    """First / last source row the resize of destination row y touches (skimage.transform.resize: bilinear on the half-pixel
    grid, Gaussian of radius int(4 sigma + 0.5), sigma = (src/dst - 1)/2, when shrinking) — before mirroring."""
    if src == dst:
        return y, y
    zoom = src / dst
    r = int(4.0 * max(0.0, (zoom - 1.0) / 2.0) + 0.5) if dst < src else 0
    cc = abs((y + 0.5) * zoom - 0.5)
    s0 = int(np.floor(cc))
    return s0 - r, s0 + 1 + r


def band_out_rows_py(infer_h: int, raw_h: int, S: int, tile_rows: Tuple[int, int]) -> Tuple[int, int]:
    """Pure-Python twin of umx_band_out_rows: raw-grid rows [r0, r1) a tile-row band writes when the network runs at
    ``infer_h`` rows and the pages are resized back to ``raw_h`` rows.  The cut between two bands is the first raw row
    whose resize window reaches the inference rows the upper band does not emit."""
    if raw_h == infer_h:
        return band_rows_py(infer_h, S, tile_rows)
    m, sub, npr, _ = tile_geometry(infer_h, 1, S)
    a, b = tile_rows
    b = npr if b <= 0 or b > npr else b

    def cut(t: int) -> int:
        if t <= 0:
            return 0
        if t >= npr:
            return raw_h
        bound = t * sub - m
        y = 0
        while y < raw_h and _resize_window(y, infer_h, raw_h)[1] < bound:
            y += 1
        return y

    return cut(a), cut(b)


AUTO_TOLERANCE = 1e-3     # max|dp| budget of the chosen mode vs split3 on the probe tiles: half of the 2e-3 contract
AUTO_PROBE_TILES = 64     # probe tiles sampled from the image being processed (corners, edges, interior)


def _premap_apply(x: np.ndarray, premap: Optional["PreMap"]) -> np.ndarray:
    """The float64 sample map of gather_tiles_kernel (img_as_float scale, optional rescale_intensity stretch)."""
    x = x.astype(np.float64)
    if premap is None:
        return x
    x = x * premap.in_scale
    if premap.rescale:
        x = np.minimum(np.maximum(x, premap.imin), premap.imax)
        x = (x - premap.imin) / (premap.imax - premap.imin)
        x = x * (premap.omax - premap.omin) + premap.omin
    return x


def probe_tile_indices(npr: int, npc: int, n: int, seed: int = 0) -> List[int]:
    """Tile indices for calibration / parity probes: the four corner tiles, mid-edge tiles (these carry the zero
    padding that becomes -mean/std, PartitionOfImage.py:56-63), then interior tiles on a regular grid and seeded
    random picks.  Deterministic, no duplicates, at most n (fewer when the image has fewer tiles)."""
    total = npr * npc
    want = min(n, total)
    picks: List[int] = []

    def add(i: int, j: int):
        t = min(max(i, 0), npr - 1) * npc + min(max(j, 0), npc - 1)
        if t not in picks and len(picks) < want:
            picks.append(t)

    for i, j in ((0, 0), (0, npc - 1), (npr - 1, 0), (npr - 1, npc - 1), (0, npc // 2), (npr - 1, npc // 2),
                 (npr // 2, 0), (npr // 2, npc - 1)):
        add(i, j)
    g = max(1, int(np.sqrt(max(1, (want - len(picks)) * 3 // 4))))
    for a in range(g):
        for b in range(g):
            add((2 * a + 1) * npr // (2 * g), (2 * b + 1) * npc // (2 * g))
    rng = np.random.default_rng(seed)
    guard = 0
    while len(picks) < want and guard < 100 * want:
        add(int(rng.integers(npr)), int(rng.integers(npc)))
        guard += 1
    return picks


def _resize_host(a: np.ndarray, out_shape: Tuple[int, int]) -> np.ndarray:
    """skimage.transform.resize defaults on a float64 array (scipy.ndimage): only used to cut calibration tiles."""
    from scipy import ndimage as ndi
    fy, fx = a.shape[0] / out_shape[0], a.shape[1] / out_shape[1]
    if fy > 1 or fx > 1:
        a = ndi.gaussian_filter(a, (max(0.0, (fy - 1) / 2), max(0.0, (fx - 1) / 2)), mode="mirror")
    return ndi.zoom(a, (out_shape[0] / a.shape[0], out_shape[1] / a.shape[1]), order=1, mode="mirror", grid_mode=True)


def sample_probe_tiles(image: np.ndarray, S: int, C: int, mean: float, std: float, premap=None,
                       n: int = AUTO_PROBE_TILES, seed: int = 0, indices: Optional[Sequence[int]] = None,
                       infer_shape: Optional[Tuple[int, int]] = None) -> np.ndarray:
    """Network inputs [n,S,S,C] float32 of n PI2D tiles of ``image`` ([H,W] or [C,H,W]), built on the host as
    gather_tiles_kernel builds them on the device (PartitionOfImage.py:49-82 + UnMicst1-5.py:700): frame = premap(sample)
    inside the image and 0 outside, (frame - mean)/std in float64, rounded to float32.  ``premap``: one PreMap or one
    per channel.  With ``infer_shape`` the tiles are cut from the image resized to that shape (--scalingFactor); each
    tile is resized from its own source window, which differs from the whole-image resize only by rounding and by the
    window's mirror edge — good for calibration, not a parity reference."""
    img = np.asarray(image)
    planes = [img] if img.ndim == 2 else [img[c] for c in range(img.shape[0])]
    RH, RW = planes[0].shape
    H, W = (RH, RW) if infer_shape is None else (int(infer_shape[0]), int(infer_shape[1]))
    scaled = (H, W) != (RH, RW)
    m, sub, npr, npc = tile_geometry(H, W, S)
    idx = list(indices) if indices is not None else probe_tile_indices(npr, npc, n, seed)
    out = np.empty((len(idx), S, S, C), dtype=np.float32)
    for k, t in enumerate(idx):
        ti, tj = divmod(int(t), npc)
        r0, c0 = ti * sub - m, tj * sub - m                       # image coordinates of the tile's first pixel
        ra, rb, ca, cb = max(r0, 0), min(r0 + S, H), max(c0, 0), min(c0 + S, W)
        for c in range(C):
            frame = np.zeros((S, S), dtype=np.float64)
            if rb > ra and cb > ca:
                src = planes[0 if len(planes) == 1 else c]
                pm = premap[c] if isinstance(premap, (list, tuple)) else premap
                if not scaled:
                    frame[ra - r0:rb - r0, ca - c0:cb - c0] = _premap_apply(src[ra:rb, ca:cb], pm)
                else:
                    # source window of the tile (+ a margin for the interpolation taps), resized with the global zoom
                    zy, zx = RH / H, RW / W
                    pad = 8
                    sa, sb = max(0, int(ra * zy) - pad), min(RH, int(np.ceil(rb * zy)) + pad)
                    ta, tb = max(0, int(ca * zx) - pad), min(RW, int(np.ceil(cb * zx)) + pad)
                    oa, ob = int(round(sa / zy)), int(round(sb / zy))
                    pa, pb = int(round(ta / zx)), int(round(tb / zx))
                    win = src[sa:sb, ta:tb].astype(np.float64) * (pm.in_scale if pm is not None else 1.0)
                    up = _resize_host(win, (max(ob - oa, 1), max(pb - pa, 1)))
                    sub_img = up[np.clip(np.arange(ra, rb) - oa, 0, up.shape[0] - 1)][:, np.clip(np.arange(ca, cb) - pa, 0, up.shape[1] - 1)]
                    if pm is not None and pm.rescale:
                        sub_img = np.minimum(np.maximum(sub_img, pm.imin), pm.imax)
                        sub_img = (sub_img - pm.imin) / (pm.imax - pm.imin) * (pm.omax - pm.omin) + pm.omin
                    frame[ra - r0:rb - r0, ca - c0:cb - c0] = sub_img
            out[k, :, :, c] = ((frame - mean) / std).astype(np.float32)
    return out


def mask_to_reserved(mask: int) -> Tuple[int, int]:
    """64-bit op mask of UMX_PREC_MIXED -> the two int32 fields umx_model_desc.reserved[0] (low word) and [1] (high
    word) carry it in: same bit patterns, expressed as signed 32-bit values for ctypes."""
    mask &= 2 ** 64 - 1
    words = (mask & 0xFFFFFFFF, mask >> 32)
    return tuple(w - (1 << 32) if w >= 1 << 31 else w for w in words)


def describe_plan(model: Model, precision: str = "default", single_mask: int = 0) -> List[str]:
    """The op list the library would build for ``model`` (host-only ``umx_describe_plan``, no GPU needed): one line per
    op with its kernel family (tensor / first / simt), tensor-path mode, operand planes, sources and FLOPs per tile."""
    hp = model.hp
    desc = _lib.umx_model_desc()
    desc.abi_version = _lib.UMX_ABI_VERSION
    desc.graph = {LEGACY: _lib.UMX_GRAPH_LEGACY, V2: _lib.UMX_GRAPH_V2}[model.variant]
    desc.im_size, desc.n_channels, desc.n_classes = int(hp["imSize"]), int(hp["nChannels"]), int(hp["nClasses"])
    desc.n_out0, desc.n_layers = int(hp["nOut0"]), int(hp["nLayers"])
    desc.feat_maps_fact, desc.down_samp_fact = int(hp["featMapsFact"]), int(hp["downSampFact"])
    desc.ks, desc.n_extra_convs = int(hp["ks"]), int(hp["nExtraConvs"])
    desc.precision = _lib.PRECISIONS[precision]
    if precision == "mixed":
        desc.reserved[0], desc.reserved[1] = mask_to_reserved(single_mask)
    names = sorted(model.weights)
    arr = (_lib.umx_tensor * len(names))()
    keep = []
    for i, n in enumerate(names):
        a = np.ascontiguousarray(model.weights[n], dtype=np.float32)
        keep.append(a)
        arr[i].name = n.encode()
        arr[i].data = a.ctypes.data_as(C.POINTER(C.c_float))
        arr[i].ndim = a.ndim
        for d in range(a.ndim):
            arr[i].shape[d] = a.shape[d]
    buf = C.create_string_buffer(1 << 16)
    n = lib().umx_describe_plan(C.byref(desc), arr, len(names), buf, len(buf))
    if n < 0:
        check(int(n))
    return buf.value.decode().splitlines()


def tensor_ops(model: Model) -> List[Tuple[int, str]]:
    """(op index, name) of every op the tensor path runs in the split plan: the candidates of per-layer precision."""
    return [(i, n) for i, n, _ in tensor_op_sources(model)]


def tensor_op_sources(model: Model) -> List[Tuple[int, str, List[float]]]:
    """(op index, name, [K of source 0, K of source 1]) per tensor-path op, K = taps x channels the source contributes to
    the contraction (the cost weight of its correction terms).  Parsed from the host-only plan description."""
    out = []
    for line in describe_plan(model, "split3"):
        f = line.split()
        if len(f) > 3 and f[1] == "conv" and f[3] == "tensor":
            terms = line.split("terms=")[1].replace("]", "").split("[")[1:]
            ks: List[float] = []
            for t in terms:                                   # "k=3 a:80|b:160"  or  "k=3 up:80" + "k=1 taps:9"
                parts = t.split()
                k = int(parts[0][2:])
                for srcs in parts[1].split("|"):
                    ks.append(float(k * k * int(srcs.rsplit(":", 1)[1])))
            if "mode=3" in f:                                 # the second term is the fp32 one-channel shortcut of the epilogue
                ks = ks[:1]
            out.append((int(f[0]), f[2], ks[:2]))
    return out


def choose_single_mask(errs: Dict[int, float], costs: Dict[int, float], budget: float, exclude: Sequence[int] = ()) -> int:
    """Error-budgeted layer selection (binary form): the subset of ops to run with one MMA per product that saves the
    most time while the predicted max|dp| — the per-layer contributions added in quadrature (rounding errors of
    different layers are independent) — stays inside ``budget``.  Exact over all subsets up to 20 candidates."""
    ops = [i for i in errs if i not in exclude and costs.get(i, 0.0) > 0.0]
    if not ops:
        return 0
    b2 = budget * budget
    if len(ops) <= 20:
        best_mask, best_gain = 0, 0.0
        e2 = [errs[i] ** 2 for i in ops]
        cs = [costs[i] for i in ops]
        for sub in range(1, 1 << len(ops)):
            e, g = 0.0, 0.0
            for k in range(len(ops)):
                if sub >> k & 1:
                    e += e2[k]
                    g += cs[k]
            if e <= b2 and g > best_gain:
                best_gain, best_mask = g, sub
        return sum(1 << ops[k] for k in range(len(ops)) if best_mask >> k & 1)
    mask, e = 0, 0.0
    for i in sorted(ops, key=lambda i: errs[i] ** 2 / costs[i]):
        if e + errs[i] ** 2 <= b2:
            e += errs[i] ** 2
            mask |= 1 << i
    return mask


def choose_op_terms(options: Dict[int, List[Tuple[int, float, float]]], budget: float, bins: int = 2000) -> Dict[int, int]:
    """Multiple-choice knapsack behind `auto`: ``options[op]`` lists (terms, max|dp| when only this op runs with these
    terms, time); pick one option per op minimising the total time while the contributions, added in quadrature, stay
    inside ``budget``.  Dynamic programme over the squared-error budget in ``bins`` steps (errors rounded UP to a bin,
    so the answer is feasible; every op must offer an option with zero error: the full split)."""
    b2 = budget * budget
    step = b2 / bins
    INF = float("inf")
    best = [0.0] * (bins + 1)             # best[j]: least time with squared error <= j * step, over the ops seen so far
    choice: List[List[int]] = []
    ops = sorted(options)
    table = [best]
    for op in ops:
        prev = table[-1]
        cur = [INF] * (bins + 1)
        pick = [-1] * (bins + 1)
        for oi, (_, err, t) in enumerate(options[op]):
            w = 0 if err <= 0 else int(-(-(err * err) // step))
            if w > bins:
                continue
            for j in range(w, bins + 1):
                v = prev[j - w] + t
                if v < cur[j]:
                    cur[j], pick[j] = v, oi
        table.append(cur)
        choice.append(pick)
    if table[-1][bins] == INF:
        raise ValueError("no feasible precision assignment (every op needs a zero-error option)")
    out: Dict[int, int] = {}
    j = bins
    for k in range(len(ops) - 1, -1, -1):
        oi = choice[k][j]
        terms, err, _ = options[ops[k]][oi]
        out[ops[k]] = terms
        j -= 0 if err <= 0 else int(-(-(err * err) // step))
    return out


def calibrate(model: Model, device: int = 0, probe_tiles: Optional[np.ndarray] = None, budget: float = AUTO_TOLERANCE,
              verbose: bool = False) -> Tuple[str, Dict[int, int], Dict]:
    """Pick the cheapest arithmetic that keeps max|dp| vs the full hi/lo split within ``budget`` on ``probe_tiles``
    (tiles of the image about to be processed: sample_probe_tiles).  Returns (precision, op_terms, report).

    The full split adds two correction MMAs to every product: a_hi*w_lo (weight rounding, term bit 0) and a_lo*w_hi
    (activation rounding, bit 1).  They can be switched per layer and per concat source.
    1. all layers with one MMA per product within budget -> 'single'.
    2. otherwise, on ONE live split engine (umx_set_op_terms, no rebuilds): for every tensor-path layer and every term
       combination, the max|dp| this layer alone causes; time per option interpolated between the layer's profiled
       single and split times by the share of correction MMAs it issues; choose_op_terms picks the fastest assignment
       whose contributions fit the budget in quadrature; the chosen assignment is then measured as a whole on the same
       engine and the budget tightened until the measured error fits.
    """
    S, C = int(model.hp["imSize"]), int(model.hp["nChannels"])
    if probe_tiles is None:
        rng = np.random.default_rng(2024)
        probe_tiles = rng.normal(size=(16, S, S, C)).astype(np.float32)
    probe_tiles = np.ascontiguousarray(probe_tiles, dtype=np.float32)
    n_time = max(len(probe_tiles), min(1024, (1 << 22) // (S * S)))        # enough tiles for meaningful kernel times
    reps = -(-n_time // len(probe_tiles))
    timing_tiles = np.concatenate([probe_tiles] * reps)[:n_time] if reps > 1 else probe_tiles
    batch = max(64, n_time)

    def profile_of(e: "Engine") -> List[Dict]:
        e.forward_tiles(timing_tiles)                 # warm-up
        e.profile_enable(True)
        e.forward_tiles(timing_tiles)
        prof = e.profile_read()
        e.profile_enable(False)
        return prof

    srcs = [(i, n, ks) for i, n, ks in tensor_op_sources(model)]
    with Engine(model, device, "single", batch) as es:
        got_single = es.forward_tiles(probe_tiles)
        prof_single = profile_of(es)
        resident = {i for i, _, _ in srcs if es.op_info(i)["resident"]}       # weight-stationary layers (64 x 64 grids)
    with Engine(model, device, "split3", batch) as ref:
        want = ref.forward_tiles(probe_tiles)
        d_single = float(np.abs(got_single - want).max())
        report: Dict = {"probe_tiles": int(len(probe_tiles)), "budget": budget, "tolerance": budget,
                        "single_vs_split3_max_abs_dp": d_single}
        if d_single <= budget:
            report["chosen"] = "single"
            return "single", {}, report
        prof_split = profile_of(ref)
        # layers that multiply a_hi by [w_hi | w_lo] in one wide MMA (umx_op_info: resident == 2): the a_hi*w_lo term is free there
        ncat_ops = {i for i, _, _ in srcs if ref.op_info(i)["resident"] == 2}
        # per op: (terms, error alone, time by the linear model, time if hi-only weights stay resident)
        measured: Dict[int, List[Tuple[int, float, float, float]]] = {}
        layers = []
        for i, name, ks in srcs:
            two = len(ks) > 1
            k0, k1 = ks[0], (ks[1] if two else 0.0)
            t_single, t_split = prof_single[i]["ms"], max(prof_split[i]["ms"], prof_single[i]["ms"])
            opts: List[Tuple[int, float, float, float]] = []
            rec = {"op": i, "name": name, "split_ms": round(t_split, 4), "single_ms": round(t_single, 4), "dp": {}}
            for t0 in range(4):
                for t1 in (range(4) if two else (t0,)):
                    terms = t0 | (t1 << 2)
                    share = (k0 * bin(t0).count("1") + k1 * bin(t1).count("1")) / (2.0 * (k0 + k1))
                    t_lin = t_single + (t_split - t_single) * share
                    # a layer whose weights stay in shared memory in single mode keeps them there as long as no
                    # a_hi*w_lo term asks for the lo plane of the weights: its time then scales with the MMAs issued
                    # (measured on B200: 1.76x for twice the MMAs; a partial split that streams both weight planes costs
                    # ~20 % more than the MMA share suggests)
                    # (the one-tile slab of a 1x1 term - a tap-expanded raw input, K < 64 - may keep its lo plane too)
                    hi_only = not (t0 & 1) and (not two or not (t1 & 1) or k1 < 64)
                    t_res = t_single * (1.0 + 1.55 * share) if (i in resident and hi_only) else t_single + (t_split - t_single) * min(1.0, 1.2 * share)
                    if i in ncat_ops:
                        share_n = (k0 * ((t0 >> 1) & 1) + k1 * ((t1 >> 1) & 1)) / (k0 + k1)
                        t_lin = t_res = t_single + (t_split - t_single) * share_n
                    if terms == 15 or (not two and t0 == 3):
                        err = 0.0
                    else:
                        ref.set_op_terms(i, terms)
                        err = float(np.abs(ref.forward_tiles(probe_tiles) - want).max())
                        ref.set_op_terms(i, 15)
                    opts.append((terms if two else (t0 | t0 << 2), err, t_lin, t_res if terms not in (0, 15) and not (not two and t0 in (0, 3)) else t_lin))
                    rec["dp"][f"{t0}{t1}" if two else f"{t0}"] = err
            measured[i] = opts
            layers.append(rec)
            if verbose:
                print(f"[calibrate] {name:14s} split {t_split:.3f} ms single {t_single:.3f} ms  dp by terms {rec['dp']}", flush=True)

        def assign(col: int) -> Tuple[Dict[int, int], float, int]:
            options = {i: [(x[0], x[1], x[2 + col]) for x in opts] for i, opts in measured.items()}
            eff, rounds = 0.9 * budget, 0
            while True:
                rounds += 1
                pick = choose_op_terms(options, eff)
                for i, t in pick.items():
                    ref.set_op_terms(i, t)
                d = float(np.abs(ref.forward_tiles(probe_tiles) - want).max())
                if d <= budget or rounds >= 8:
                    break
                eff *= 0.85
            for i in pick:
                ref.set_op_terms(i, 15)
            if d > budget:
                return {i: 15 for i in pick}, 0.0, rounds
            return pick, d, rounds

        cands = [assign(0)]
        if resident:
            alt = assign(1)
            if alt[0] != cands[0][0]:
                cands.append(alt)
    # candidates are built for real and timed (the models above rank options, the clock decides between assignments)
    timed = []
    for pick, d, rounds in cands:
        if all(t == 15 for t in pick.values()) or len(cands) == 1:
            timed.append((None, pick, d, rounds))
            continue
        with Engine(model, device, "mixed", batch, op_terms=pick) as e:
            prof = profile_of(e)
        timed.append((sum(prof[i]["ms"] for i in pick), pick, d, rounds))
    best = min(timed, key=lambda c: (c[0] is None, c[0] or 0.0)) if any(c[0] is not None for c in timed) else timed[0]
    _, chosen, d_mixed, rounds = best
    options = {i: [(x[0], x[1], x[2]) for x in opts] for i, opts in measured.items()}
    names = {i: n for i, n, _ in srcs}
    pred = float(np.sqrt(sum(next(e for tt, e, _ in options[i] if tt == t) ** 2 for i, t in chosen.items())))
    est = sum(next(tm for tt, _, tm in options[i] if tt == t) for i, t in chosen.items())
    all_split = all(t == 15 for t in chosen.values())
    report.update({"layers": layers, "op_terms": {names[i]: [t & 3, t >> 2] for i, t in chosen.items()},
                   "single_layers": [names[i] for i, t in chosen.items() if t == 0],
                   "partial_layers": [names[i] for i, t in chosen.items() if t not in (0, 15)],
                   "mixed_vs_split3_max_abs_dp": d_mixed, "predicted_quadrature_dp": pred, "rounds": rounds,
                   "weight_stationary_layers": [names[i] for i in sorted(resident)],
                   "candidates_timed_ms": [None if c[0] is None else round(c[0], 4) for c in timed],
                   "est_tensor_ms": {"chosen": est, "single": sum(prof_single[i]["ms"] for i in chosen),
                                     "split3": sum(prof_split[i]["ms"] for i in chosen)},
                   "chosen": "split3" if all_split else "mixed"})
    return ("split3" if all_split else "mixed"), ({} if all_split else dict(chosen)), report


class Engine:
    """precision: 'default'/'split3' (fp16 hi/lo split, ~fp32 accurate, always within the 2e-3 contract),
    'single' (one fp16 MMA per product: ~3x less tensor work, accuracy depends on how steep the
    model's softmax is), 'mixed' (per layer: ``single_mask`` bit i = op i of ``profile_read`` runs single),
    'fp32' (CUDA cores only), or 'auto': ``calibrate`` on ``probe_tiles`` (tiles of the image to be processed, see
    ``sample_probe_tiles``; seeded noise tiles when none are given) against 'split3' with the AUTO_TOLERANCE budget —
    all-'single' if that fits, otherwise the error-budgeted per-layer 'mixed' selection, which may end at 'split3'."""

    def __init__(self, model: Model, device: int = 0, precision: str = "default", max_batch_tiles: int = 0,
                 probe_tiles: Optional[np.ndarray] = None, single_mask: int = 0, op_terms: Optional[Dict[int, int]] = None):
        self.auto_report = None
        self.single_mask = 0
        self.op_terms = dict(op_terms) if op_terms else {}
        if precision == "auto":
            chosen, terms, report = calibrate(model, device, probe_tiles)
            final = Engine(model, device, chosen, max_batch_tiles, op_terms=terms)
            self.__dict__.update(final.__dict__)
            final._h = None
            self.auto_report = report
            return
        L = lib()
        hp = model.hp
        self.model = model
        self.device = device
        self.precision = "split3" if precision == "default" else precision
        self.S, self.C, self.K = int(hp["imSize"]), int(hp["nChannels"]), int(hp["nClasses"])
        desc = _lib.umx_model_desc()
        desc.abi_version = _lib.UMX_ABI_VERSION
        desc.graph = {LEGACY: _lib.UMX_GRAPH_LEGACY, V2: _lib.UMX_GRAPH_V2}[model.variant]
        desc.im_size, desc.n_channels, desc.n_classes = self.S, self.C, self.K
        desc.n_out0, desc.n_layers = int(hp["nOut0"]), int(hp["nLayers"])
        desc.feat_maps_fact, desc.down_samp_fact = int(hp["featMapsFact"]), int(hp["downSampFact"])
        desc.ks, desc.n_extra_convs = int(hp["ks"]), int(hp["nExtraConvs"])
        desc.precision = _lib.PRECISIONS[precision]
        desc.max_batch_tiles = int(max_batch_tiles)
        if precision == "mixed":
            self.single_mask = int(single_mask) & (2 ** 64 - 1)
            desc.reserved[0], desc.reserved[1] = mask_to_reserved(self.single_mask)
        names = sorted(model.weights)
        arr = (_lib.umx_tensor * len(names))()
        keep = []
        for i, n in enumerate(names):
            a = np.ascontiguousarray(model.weights[n], dtype=np.float32)
            keep.append(a)
            arr[i].name = n.encode()
            arr[i].data = a.ctypes.data_as(C.POINTER(C.c_float))
            arr[i].ndim = a.ndim
            for d in range(a.ndim):
                arr[i].shape[d] = a.shape[d]
        h = C.c_void_p()
        if self.op_terms:               # per op and per source: which hi/lo correction terms (umx_create_ex)
            n_ops = max(self.op_terms) + 1
            ot = (C.c_int32 * n_ops)(*[int(self.op_terms.get(i, -1)) for i in range(n_ops)])
            check(L.umx_create_ex(C.byref(desc), arr, len(names), device, ot, n_ops, C.byref(h)))
        else:
            check(L.umx_create(C.byref(desc), arr, len(names), device, C.byref(h)))
        self._h = h
        self._lock = threading.Lock()

    # -- lifetime ---------------------------------------------------------------------------
    def close(self) -> None:
        if getattr(self, "_h", None):
            lib().umx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- Session.run ------------------------------------------------------------------------
    def forward_tiles(self, tiles: np.ndarray, precision: str = "default") -> np.ndarray:
        t = np.ascontiguousarray(tiles, dtype=np.float32)
        if t.ndim != 4 or t.shape[1:] != (self.S, self.S, self.C):
            raise ValueError(f"tiles must be [n,{self.S},{self.S},{self.C}], got {t.shape}")
        out = np.empty((t.shape[0], self.S, self.S, self.K), dtype=np.float32)
        with self._lock:
            check(lib().umx_forward_tiles(self._h, t.ctypes.data, t.shape[0], out.ctypes.data, _lib.PRECISIONS[precision]))
        return out

    # -- singleImageInference (all classes) ------------------------------------------------------
    def band_rows(self, H: int, tile_rows: Optional[Tuple[int, int]] = None) -> Tuple[int, int]:
        r0, r1 = C.c_int32(0), C.c_int32(0)
        tr0, tr1 = tile_rows if tile_rows else (0, 0)
        check(lib().umx_band_rows(self._h, H, tr0, tr1, C.byref(r0), C.byref(r1)))
        return r0.value, r1.value

    def band_out_rows(self, infer_h: int, raw_h: int, tile_rows: Optional[Tuple[int, int]] = None) -> Tuple[int, int]:
        """Raw-grid rows a tile-row band writes with ``cli_quant`` when the network runs at ``infer_h`` rows."""
        r0, r1 = C.c_int32(0), C.c_int32(0)
        tr0, tr1 = tile_rows if tile_rows else (0, 0)
        check(lib().umx_band_out_rows(self._h, infer_h, tr0, tr1, raw_h, C.byref(r0), C.byref(r1)))
        return r0.value, r1.value

    def resample_minmax(self, plane: np.ndarray, out_shape: Tuple[int, int], in_scale: float = 1.0) -> Tuple[float, float]:
        """(min, max) of skimage.transform.resize(img_as_float(plane), out_shape), computed on the GPU."""
        a = np.ascontiguousarray(plane)
        if a.ndim != 2 or a.dtype not in _NP2UMX:
            raise TypeError("plane must be a 2-D uint8/uint16/float32/float64 array")
        lo, hi = C.c_double(0), C.c_double(0)
        with self._lock:
            check(lib().umx_resample_minmax(self._h, a.ctypes.data, _NP2UMX[a.dtype], a.shape[0], a.shape[1],
                                            int(out_shape[0]), int(out_shape[1]), float(in_scale), C.byref(lo), C.byref(hi)))
        return lo.value, hi.value

    def infer_ptr(self, img_ptr: int, dtype: int, n_planes: int, H: int, W: int, plane_stride: int,
                  mean: float, std: float, out_u8_ptr: int = 0, out_f32_ptr: int = 0,
                  tile_rows: Optional[Tuple[int, int]] = None, premap: Optional[PreMap] = None,
                  out_plane_stride: int = 0, out_row_base: int = 0, precision: str = "default",
                  no_sync: bool = False, infer_shape: Optional[Tuple[int, int]] = None, cli_quant: bool = False,
                  continue_prev: bool = False, stitch_mode: str = "accumulate", fp16_quant: bool = False) -> None:
        """Raw-pointer form (host pageable / pinned or device memory — the library detects which).
        ``infer_shape``: run the network on the image resized to this (rows, cols) (--scalingFactor);
        ``cli_quant``: out_u8 = the page the reference CLI writes ([K,H,W] of the raw grid, quantised twice)."""
        o = _lib.umx_opts()
        if tile_rows:
            o.tile_row0, o.tile_row1 = int(tile_rows[0]), int(tile_rows[1])
        o.precision = _lib.PRECISIONS[precision]
        o.flags = ((_lib.UMX_F_NO_SYNC if no_sync else 0) | (_lib.UMX_F_CLI_QUANT if cli_quant else 0) |
                   (_lib.UMX_F_CONTINUE if continue_prev else 0) | (_lib.UMX_F_STITCH_REPLACE if stitch_mode == "replace" else 0) |
                   (_lib.UMX_F_FP16_QUANT if fp16_quant else 0))
        if stitch_mode not in ("accumulate", "replace"):
            raise ValueError("stitch_mode must be 'accumulate' or 'replace' (PartitionOfImage.py:92-100)")
        if infer_shape is not None:
            o.infer_h, o.infer_w = int(infer_shape[0]), int(infer_shape[1])
        if isinstance(premap, (list, tuple)):                 # one map per image plane (unmicst-duo)
            if len(premap) != self.C:
                raise ValueError(f"{len(premap)} premaps for a network with {self.C} input channels")
            pm = (_lib.umx_premap * len(premap))(*[q.to_c() for q in premap])
            o.premap = C.cast(pm, C.POINTER(_lib.umx_premap))
            o.flags |= _lib.UMX_F_PREMAP_PER_PLANE
        else:
            pm = premap.to_c() if premap else None
            o.premap = C.pointer(pm) if pm is not None else None
        o.out_plane_stride = int(out_plane_stride)
        o.out_row_base = int(out_row_base)
        with self._lock:
            check(lib().umx_infer_image(self._h, img_ptr, dtype, n_planes, H, W, plane_stride, float(mean), float(std),
                                        out_u8_ptr or None, out_f32_ptr or None, C.byref(o)))

    def infer_image(self, image: np.ndarray, mean: Optional[float] = None, std: Optional[float] = None,
                    premap: Optional[PreMap] = None, want_u8: bool = True, want_f32: bool = False,
                    tile_rows: Optional[Tuple[int, int]] = None, out_u8: Optional[np.ndarray] = None,
                    out_f32: Optional[np.ndarray] = None, precision: str = "default",
                    infer_shape: Optional[Tuple[int, int]] = None, cli_quant: bool = False, stitch_mode: str = "accumulate",
                    fp16_quant: bool = False):
        """image [H,W] or [C,H,W] (uint8/uint16/float32/float64) -> (u8 [K,h,w] | None, f32 [K,h,w] | None).
        ``fp16_quant``: quantise as the reference's float16 pipeline does (np.uint8(255 * float16), UMX_F_FP16_QUANT).

        With ``tile_rows`` only the image rows of that band are written (into full-size outputs).
        ``infer_shape`` = (rows, cols) the network runs at (the image is resized on the GPU with
        skimage.transform.resize semantics, UnMicst1-5.py:813-815); outputs then have that shape, unless
        ``cli_quant`` asks for the reference CLI's page: resized back to [K,H,W] and quantised twice (:848-853)."""
        img = np.ascontiguousarray(image)
        if img.dtype not in _NP2UMX:
            raise TypeError(f"unsupported image dtype {img.dtype}")
        if img.ndim == 2:
            planes, (H, W) = 1, img.shape
        elif img.ndim == 3:
            planes, H, W = img.shape
        else:
            raise ValueError("image must be [H,W] or [C,H,W]")
        mean = self.model.mean if mean is None else mean
        std = self.model.std if std is None else std
        oh, ow = (H, W) if (cli_quant or infer_shape is None) else (int(infer_shape[0]), int(infer_shape[1]))
        if cli_quant:
            want_u8, want_f32 = True, False
        if want_u8 and out_u8 is None:
            out_u8 = np.zeros((self.K, oh, ow), dtype=np.uint8)
        if want_f32 and out_f32 is None:
            out_f32 = np.zeros((self.K, oh, ow), dtype=np.float32)
        self.infer_ptr(img.ctypes.data, _NP2UMX[img.dtype], planes, H, W, H * W, mean, std,
                       out_u8.ctypes.data if out_u8 is not None else 0,
                       out_f32.ctypes.data if out_f32 is not None else 0,
                       tile_rows=tile_rows, premap=premap, precision=precision, infer_shape=infer_shape, cli_quant=cli_quant,
                       stitch_mode=stitch_mode, fp16_quant=fp16_quant)
        return out_u8, out_f32

    def stream_image(self, image: np.ndarray, mean: Optional[float] = None, std: Optional[float] = None,
                     premap=None, chunk_tile_rows: int = 0, infer_shape: Optional[Tuple[int, int]] = None,
                     cli_quant: bool = False):
        """Generator over row bands of the uint8 maps: yields (row0, row1, u8 [K, row1-row0, w]) as each band of
        ``chunk_tile_rows`` tile rows leaves the GPU, so a consumer (the BigTIFF writer, UnMicst1-5.py:852-862) works on
        band i while band i+1 is computed.  Bands continue each other on the device (UMX_F_CONTINUE): nothing is
        recomputed and the bytes equal one whole-image call."""
        img = np.ascontiguousarray(image)
        if img.dtype not in _NP2UMX:
            raise TypeError(f"unsupported image dtype {img.dtype}")
        planes, (H, W) = (1, img.shape) if img.ndim == 2 else (img.shape[0], img.shape[1:])
        mean = self.model.mean if mean is None else mean
        std = self.model.std if std is None else std
        ih, iw = (H, W) if infer_shape is None else (int(infer_shape[0]), int(infer_shape[1]))
        _, _, npr, npc = tile_geometry(ih, iw, self.S)
        if chunk_tile_rows <= 0:
            chunk_tile_rows = max(1, min(npr, -(-8192 // npc)))           # ~8k tiles per band
        ow = W if cli_quant else iw
        for a in range(0, npr, chunk_tile_rows):
            b = min(npr, a + chunk_tile_rows)
            r0, r1 = self.band_out_rows(ih, H, (a, b)) if cli_quant else self.band_rows(ih, (a, b))
            buf = np.empty((self.K, max(r1 - r0, 0), ow), dtype=np.uint8)
            self.infer_ptr(img.ctypes.data, _NP2UMX[img.dtype], planes, H, W, H * W, mean, std, out_u8_ptr=buf.ctypes.data,
                           tile_rows=(a, b), premap=premap, out_plane_stride=buf.shape[1] * ow, out_row_base=r0,
                           infer_shape=infer_shape, cli_quant=cli_quant, continue_prev=a > 0)
            if r1 > r0:
                yield r0, r1, buf

    # -- many small images per launch (TMA cores) ---------------------------------------------------------
    def infer_images(self, images: Sequence[np.ndarray], premaps: Optional[Sequence[Optional[PreMap]]] = None,
                     mean: Optional[float] = None, std: Optional[float] = None, want_f32: bool = False,
                     cli_quant: bool = False) -> List[np.ndarray]:
        """``umx_infer_images``: every image ([H,W] or [C,H,W]) through the resident model, tiles of several images
        sharing each network launch (the dearray loop of batchUNet2DTMACycif.py:539-569).  Returns one uint8 [K,H,W]
        (or float32 with ``want_f32``) per image, equal to ``infer_image`` on each."""
        n = len(images)
        arr = (_lib.umx_image * n)()
        keep, outs = [], []
        for i, im in enumerate(images):
            a = np.ascontiguousarray(im)
            if a.dtype not in _NP2UMX:
                raise TypeError(f"unsupported image dtype {a.dtype}")
            planes, (H, W) = (1, a.shape) if a.ndim == 2 else (a.shape[0], a.shape[1:])
            out = np.zeros((self.K, H, W), dtype=np.float32 if want_f32 else np.uint8)
            pm = premaps[i].to_c() if premaps is not None and premaps[i] is not None else None
            keep.append((a, pm))
            outs.append(out)
            arr[i].img, arr[i].dtype, arr[i].n_planes, arr[i].H, arr[i].W = a.ctypes.data, _NP2UMX[a.dtype], planes, H, W
            arr[i].plane_stride = H * W
            arr[i].premap = C.pointer(pm) if pm is not None else None
            if want_f32:
                arr[i].out_f32 = out.ctypes.data
            else:
                arr[i].out_u8 = out.ctypes.data
        mean = self.model.mean if mean is None else mean
        std = self.model.std if std is None else std
        with self._lock:
            check(lib().umx_infer_images(self._h, arr, n, float(mean), float(std), _lib.UMX_F_CLI_QUANT if cli_quant else 0))
        return outs

    OP_INFO_FIELDS = ("tensor", "halo", "pair", "stages", "b_stages", "gb", "resident", "merge_px", "planes_a", "planes_b", "terms0", "terms1")

    def op_info(self, op_index: int) -> Dict[str, int]:
        """How op ``op_index`` was lowered (umx_op_info)."""
        buf = (C.c_int32 * 12)()
        check(lib().umx_op_info(self._h, int(op_index), buf, 12))
        return dict(zip(self.OP_INFO_FIELDS, [int(v) for v in buf]))

    def set_op_terms(self, op_index: int, terms: int) -> None:
        """Calibration aid (split3 engines only): correction terms of one op, t0 | t1 << 2."""
        check(lib().umx_set_op_terms(self._h, int(op_index), int(terms)))

    # -- instrumentation -----------------------------------------------------------------------
    def set_stream(self, cuda_stream: int) -> None:
        check(lib().umx_set_stream(self._h, int(cuda_stream)))

    def profile_enable(self, on: bool = True) -> None:
        check(lib().umx_profile_enable(self._h, int(on)))

    def profile_read(self, reset: bool = True) -> List[Dict]:
        cap = 128
        buf = (_lib.umx_prof_entry * cap)()
        n = lib().umx_profile_read(self._h, buf, cap, int(reset))
        if n < 0:
            check(n)
        return [dict(name=buf[i].name.decode(), launches=int(buf[i].launches), ms=float(buf[i].ms_total),
                     flops=float(buf[i].flops), bytes=float(buf[i].bytes)) for i in range(min(n, cap))]

    def debug_buffer(self, name: str, n_tiles: int, shape_hwc: Tuple[int, int, int]) -> np.ndarray:
        """Activation written by op ``name`` during the last forward, as fp32 [n,h,w,c] (bring-up aid)."""
        hh, ww, cc = shape_hwc
        out = np.empty((n_tiles, hh, ww, cc), dtype=np.float32)
        r = lib().umx_debug_buffer(self._h, name.encode(), n_tiles, out.ctypes.data, out.size)
        if r < 0:
            check(int(r))
        if r != hh * ww * cc:
            raise ValueError(f"buffer {name} has {r} elements per tile, expected {hh * ww * cc}")
        return out

    @property
    def launch_count(self) -> int:
        return int(lib().umx_launch_count(self._h))


class MultiEngine:
    """The same model on several GPUs of one box; an image is cut into contiguous tile-row
    bands, one per GPU, each driven by its own host thread (ctypes releases the GIL).
    Every band recomputes the single tile row above its seam, so band outputs are
    bit-identical to the 1-GPU result and no device-to-device exchange exists.
    precision 'auto' calibrates once (on the first device) and builds every engine with that result, so all bands
    run the same arithmetic."""

    def __init__(self, model: Model, devices: Sequence[int], precision: str = "default", max_batch_tiles: int = 0,
                 probe_tiles: Optional[np.ndarray] = None, single_mask: int = 0, op_terms: Optional[Dict[int, int]] = None):
        if not devices:
            raise ValueError("MultiEngine needs at least one device")
        self.auto_report = None
        if precision == "auto":
            precision, op_terms, self.auto_report = calibrate(model, devices[0], probe_tiles)
        self.engines = [Engine(model, d, precision, max_batch_tiles, single_mask=single_mask, op_terms=op_terms) for d in devices]
        self.model = model
        self.precision, self.single_mask, self.op_terms = self.engines[0].precision, self.engines[0].single_mask, self.engines[0].op_terms
        self.S, self.C, self.K = self.engines[0].S, self.engines[0].C, self.engines[0].K

    def close(self) -> None:
        for e in self.engines:
            e.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def launch_count(self) -> int:
        return sum(e.launch_count for e in self.engines)

    def resample_minmax(self, plane, out_shape, in_scale: float = 1.0):
        return self.engines[0].resample_minmax(plane, out_shape, in_scale)

    def infer_image(self, image: np.ndarray, mean=None, std=None, premap=None, want_u8=True, want_f32=False,
                    precision: str = "default", infer_shape: Optional[Tuple[int, int]] = None, cli_quant: bool = False,
                    stitch_mode: str = "accumulate"):
        img = np.ascontiguousarray(image)
        H, W = img.shape[-2:]
        ih, iw = (H, W) if infer_shape is None else (int(infer_shape[0]), int(infer_shape[1]))
        _, _, npr, _ = tile_geometry(ih, iw, self.S)
        bands = split_tile_rows(npr, len(self.engines))
        oh, ow = (H, W) if cli_quant else (ih, iw)
        if cli_quant:
            want_u8, want_f32 = True, False
        out_u8 = np.zeros((self.K, oh, ow), dtype=np.uint8) if want_u8 else None
        out_f32 = np.zeros((self.K, oh, ow), dtype=np.float32) if want_f32 else None
        errs: List[BaseException] = []

        def work(eng: Engine, band):
            try:
                eng.infer_image(img, mean, std, premap, want_u8, want_f32, tile_rows=band, out_u8=out_u8,
                                out_f32=out_f32, precision=precision, infer_shape=infer_shape, cli_quant=cli_quant,
                                stitch_mode=stitch_mode)
            except BaseException as ex:  # re-raised on the caller's thread
                errs.append(ex)

        threads = [threading.Thread(target=work, args=(e, b)) for e, b in zip(self.engines, bands)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errs:
            raise errs[0]
        return out_u8, out_f32
