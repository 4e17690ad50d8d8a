#!/usr/bin/env python
"""Benchmark of the UnMicst probability-map hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload solo20k|solo40k|duo4k|cyto2tma|legacy20k|sample105|solo105] [--size PX]
                    [--precision auto|split3|single|fp32] [--configs all|none|a,b,...]

A "step" = one pass of the hot path (gather/normalise -> UNet -> stitch -> uint8) over one synthetic slide.
Default workload = BASELINE.json configs[2]: unmicst-solo graph (nucleiDAPI1-5 shapes), 20 000 x 20 000 px
synthetic DNA image, tile rows sharded over the N GPUs (strong scaling, no collective).  Prints ONE JSON line (rank 0).

Weights: the real checkpoint where the repository ships one (the legacy fixtures); otherwise seeded stand-ins whose
last linear map is scaled until max|logit| on tiles of the benchmark image is ~20, the steepness of the real models
 - the regime in which reduced-precision operands are NOT automatically inside the
north_star tolerance.  `--precision auto` (default) therefore calibrates on 64 tiles of the image being processed and
picks, per layer, one fp16 MMA per product or the hi/lo split (engine.calibrate), with half the contract as budget.

  value        megapixels/s with the image and the outputs resident in HBM
  e2e          same metric through the public API with pinned HOST buffers (H2D + D2H inside)
  parity       the timed engine vs the fp32 ORACLE on those 64 tiles (max|dp|, argmax) and, for the stitched uint8 maps,
               a crop of the timed output vs the oracle pipeline on the same crop
  modes        the same slide in plain `single` and `split3`, for reference
  roofline     dominant kernel, algorithmic FLOPs / CUDA-event time vs MEASURED_PEAKS.json; `step` = whole step vs the
               summed per-layer roofline
  cpu_baseline the oracle port (torch-CPU fp32 + Python PI2D loop) on a fixed tile sample of the same image
  configs      one compact entry per other BASELINE.json config, run the same way at this N
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TARGET_MAX_LOGIT = 20.0       # steepness of the stand-in weights (real nucleiDAPI: logits up to |22.5|)
SAMPLE_TIF = os.path.join(ROOT, "tests", "golden", "sample", "105.tif")

WORKLOADS = {
    # name: model, raw H x W, image planes, scalingFactor, BASELINE.json config
    "solo20k": dict(model="nucleiDAPI1-5", H=20000, W=20000, planes=1, scale=1.0, cfg="configs[2] unmicst-solo whole-slide 20k x 20k"),
    "solo40k": dict(model="nucleiDAPI1-5", H=20000, W=20000, planes=1, scale=2.0,
                    cfg="configs[3] unmicst-solo --scalingFactor 2.0: raw 20k x 20k, resized on the GPU, inference at 40k x 40k, pages resized back"),
    "duo4k": dict(model="nucleiDAPILAMIN", H=4096, W=4096, planes=2, scale=1.0, cfg="configs[1] unmicst-duo 4k x 4k"),
    "cyto2tma": dict(model="CytoplasmIncell2", H=18432, W=30720, planes=1, scale=1.0, cfg="configs[4] UnMicstCyto2 60-core TMA montage"),
    "legacy20k": dict(model="nucleiDAPI", H=20000, W=20000, planes=1, scale=1.0, cfg="unmicst-legacy 20k x 20k (not a BASELINE config)"),
    "sample105": dict(model="nucleiDAPI", H=832, W=960, planes=1, scale=1.0, cfg="configs[0] sample image 105.tif, legacy graph, real checkpoint", file=SAMPLE_TIF),
    "solo105": dict(model="nucleiDAPI1-5", H=832, W=960, planes=1, scale=1.0, cfg="configs[0] sample image 105.tif, unmicst-solo graph", file=SAMPLE_TIF),
}
DEFAULT_CONFIGS = ["sample105", "solo105", "duo4k", "solo40k", "cyto2tma", "cyto2cores"]


def synthetic_dna(H: int, W: int, seed: int = 1234, lamin: bool = False) -> np.ndarray:
    """Deterministic synthetic DNA-channel image: background N(800,60), anisotropic
    Gaussian nuclei ~1 per 22x22 px, sigma U(3,6), peak logN(9000, 0.5), shot noise, uint16.
    Built from a 2200 x 2200 periodic block so 400 MP take seconds, not minutes."""
    rng = np.random.default_rng(seed)
    B = 2200
    block = np.zeros((B, B), dtype=np.float32)
    g = 22
    ys, xs = np.meshgrid(np.arange(g // 2, B, g), np.arange(g // 2, B, g), indexing="ij")
    cy = (ys + rng.uniform(-7, 7, ys.shape)).ravel()
    cx = (xs + rng.uniform(-7, 7, xs.shape)).ravel()
    keep = rng.random(cy.size) < 0.85
    cy, cx = cy[keep], cx[keep]
    sy, sx = rng.uniform(3, 6, cy.size), rng.uniform(3, 6, cy.size)
    peak = np.exp(rng.normal(np.log(9000), 0.5, cy.size))
    R = 18
    yy, xx = np.mgrid[-R:R + 1, -R:R + 1].astype(np.float32)
    for i in range(cy.size):
        y0, x0 = int(round(cy[i])), int(round(cx[i]))
        blob = peak[i] * np.exp(-0.5 * (((yy - (cy[i] - y0)) / sy[i]) ** 2 + ((xx - (cx[i] - x0)) / sx[i]) ** 2))
        if lamin:
            blob = peak[i] * 0.6 * np.exp(-0.5 * ((np.sqrt(((yy - (cy[i] - y0)) / sy[i]) ** 2 +
                                                         ((xx - (cx[i] - x0)) / sx[i]) ** 2) - 1.6) / 0.35) ** 2)
        rows = (np.arange(y0 - R, y0 + R + 1) % B)[:, None]
        cols = (np.arange(x0 - R, x0 + R + 1) % B)[None, :]
        block[rows, cols] += blob
    out = np.empty((H, W), dtype=np.uint16)
    reps_x = -(-W // B)
    row_block = np.tile(block, (1, reps_x))[:, :W]
    step = 1000
    for r in range(0, H, step):
        n = min(step, H - r)
        idx = (np.arange(r, r + n) % B)
        clean = row_block[idx] + 800.0
        noise = rng.standard_normal((n, W), dtype=np.float32)
        v = clean + noise * np.sqrt(clean) + rng.standard_normal((n, W), dtype=np.float32) * 60.0
        out[r:r + n] = np.clip(v, 0, 65535).astype(np.uint16)
    return out


def make_image(workload: str, H: int, W: int) -> np.ndarray:
    wl = WORKLOADS[workload]
    if wl.get("file") and os.path.exists(wl["file"]) and (H, W) == (wl["H"], wl["W"]):
        from unmicst_b200 import tiffio
        return np.ascontiguousarray(tiffio.read_page(wl["file"], 0).astype(np.uint16))
    if wl["planes"] == 1:
        return synthetic_dna(H, W)
    return np.stack([synthetic_dna(H, W), synthetic_dna(H, W, lamin=True)])


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tc_burst=d["bf16_tflops"], tc_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured")
    return dict(hbm=6650.0, tc_burst=1590.0, tc_sustained=1400.0, source="fallback")


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int, enabled: bool = True):
        self.index = index
        self.enabled = enabled
        self.samples = []
        self._stop = threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                r = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                    "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5)
                parts = [p.strip() for p in r.stdout.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        if self.enabled:
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.enabled:
            self._t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------------------------
# the checker: oracle/ is only ever used to judge results (parity) and as the timed CPU baseline
# ------------------------------------------------------------------------------------------------------------------
def oracle_forward(model, taps=None):
    from oracle import unet_oracle
    return lambda x: unet_oracle.forward(model.weights, model.hp, model.variant, x, taps=taps)


def host_threads(world: int = 1) -> int:
    import torch
    n = max(1, os.cpu_count() or 1)
    torch.set_num_threads(n)                 # torchrun pins OMP_NUM_THREADS=1; only rank 0 runs the checker, the other ranks wait
    return n


def cpu_baseline(model, image, scale: float, n_side: int):
    """The oracle port - torch-CPU fp32 graph with the reference's batch size and its Python PI2D loop (one pass for
    all classes: flatters the reference up to 3x) - on a FIXED sample: the top-left
    n_side x n_side tiles of the same image (64 x 64 = the 4 096 tiles of BASELINE.md section 5 when the budget allows).
    Returns MP/s of output pixels."""
    import torch
    from oracle import pi2d_oracle, prepost_oracle
    torch.set_num_threads(os.cpu_count() or 1)
    S, C, B = model.hp["imSize"], model.hp["nChannels"], model.hp["batchSize"]
    sub = S - 2 * (S // 8)
    fw = oracle_forward(model)
    h = min(image.shape[-2], int(np.ceil(n_side * sub / scale)))
    w = min(image.shape[-1], int(np.ceil(n_side * sub / scale)))
    crop = image[..., :h, :w]
    fw(np.zeros((B, S, S, C), np.float32))                                   # warm-up
    t = time.perf_counter()
    if scale != 1.0:
        cells = prepost_oracle.resize(crop, (int(h * scale), int(w * scale)))
    else:
        cells = crop.astype(np.float64) * (1.0 / 65535)
    pm = pi2d_oracle.infer_image(cells, fw, S, C, model.mean, model.std, B, accum_dtype=np.float16)
    u8 = pi2d_oracle.quantize_u8(pm)
    if scale != 1.0:
        for k in range(u8.shape[0]):
            np.uint8(255 * prepost_oracle.resize(u8[k], (h, w)))
    dt = time.perf_counter() - t
    ih, iw = cells.shape[-2:]
    tiles = (-(-ih // sub)) * (-(-iw // sub))
    return {"value": ih * iw / dt / 1e6, "unit": "MP/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"top-left {h}x{w} px of the same image = {tiles} tiles (batch {B}), torch-CPU fp32 oracle + Python PI2D loop"
                      f"{' + scipy resize both ways' if scale != 1.0 else ''}, {dt:.1f} s; TensorFlow itself is not installable here",
            "tiles": tiles, "seconds": dt}


def tiles_side_for_budget(model, budget_s: float, cap: int = 64) -> int:
    """Largest n (<= cap) such that n x n tiles of the oracle fit the budget (one timed probe batch)."""
    S, C, B = model.hp["imSize"], model.hp["nChannels"], model.hp["batchSize"]
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    fw = oracle_forward(model)
    x = np.zeros((B, S, S, C), np.float32)
    fw(x)
    t = time.perf_counter(); fw(x); per_tile = (time.perf_counter() - t) / B
    return int(max(2, min(cap, np.sqrt(budget_s / max(per_tile * 1.15, 1e-9)))))


def premap_for(model_name: str, img: np.ndarray):
    from unmicst_b200.engine import PreMap
    if model_name == "nucleiDAPI1-5":
        return PreMap(in_scale=1.0 / 65535)          # solo feeds img_as_float(u16) un-stretched (UnMicst1-5.py:816)
    planes = [img] if img.ndim == 2 else list(img)   # the other tools stretch each channel to (0, 0.983) (UnMicst.py:627-631)
    pms = [PreMap(in_scale=1.0 / 65535, rescale=True, imin=float(p.min()) / 65535, imax=float(p.max()) / 65535) for p in planes]
    return pms[0] if len(pms) == 1 else pms


def choose_model(name: str, probe_of):
    """Real checkpoint when shipped; otherwise stand-in weights at max|logit| ~ TARGET_MAX_LOGIT on the probe tiles."""
    from unmicst_b200 import modelzoo
    d = os.path.join(ROOT, "tests", "golden", "models", name)
    if os.path.exists(os.path.join(d, "model.ckpt.data-00000-of-00001")):
        return modelzoo.load_model(d), None
    base = modelzoo.synthetic_model(name, seed=0, logit_gain=1.0)
    taps = {}
    oracle_forward(base, taps)(probe_of(base))             # all probe tiles: the steepness is that of the hardest one
    gain = float(TARGET_MAX_LOGIT / max(1e-6, np.abs(taps["logits"]).max()))
    return modelzoo.synthetic_model(name, seed=0, logit_gain=gain), gain


def stitched_crop_check(model, img, premap, scale: float, got_u8_crop_fn, max_tiles: int = 150, rows_available: int = 0):
    """A crop of the TIMED uint8 output vs the oracle pipeline (prepare -> PI2D tile loop -> fp32 UNet -> stitch ->
    uint8 [-> resize back -> uint8]) on the same crop.  Tiles that lie fully inside the crop are identical to the
    whole-slide tiles, so output pixels covered only by such tiles must agree within 1 level."""
    from oracle import pi2d_oracle, prepost_oracle
    S, C = model.hp["imSize"], model.hp["nChannels"]
    m = S // 8
    sub = S - 2 * m
    H, W = img.shape[-2:]
    n_c = int(max(3, min(np.sqrt(max_tiles), (min(H, W) * scale) // sub)))
    rc = int(min(min(H, W), np.ceil(n_c * sub / scale)))
    crop = img[..., :rc, :rc]
    planes = [crop] if crop.ndim == 2 else list(crop)
    pms = premap if isinstance(premap, (list, tuple)) else [premap] * len(planes)
    cells = []
    for p, pm in zip(planes, pms):
        x = prepost_oracle.resize(p, (int(rc * scale), int(rc * scale))) if scale != 1.0 else p.astype(np.float64) * pm.in_scale
        if pm.rescale:
            x = prepost_oracle.rescale_intensity(x, (pm.imin, pm.imax), (pm.omin, pm.omax))
        cells.append(x)
    cells = cells[0] if len(cells) == 1 else np.stack(cells)
    pmaps = pi2d_oracle.infer_image(cells, oracle_forward(model), S, C, model.mean, model.std, model.hp["batchSize"], accum_dtype=np.float64)
    want = np.uint8(255 * pmaps)
    full = rc >= min(H, W)                                   # the crop is the whole image: every pixel is comparable
    if scale != 1.0:
        want = np.stack([np.uint8(255 * prepost_oracle.resize(w, (rc, rc))) for w in want])
        v = rc if full else int(((n_c - 1) * sub - m) / scale) - 8
    else:
        v = rc if full else (n_c - 1) * sub - m
    v = max(1, min(v, rc))
    if rows_available:                                       # several GPUs: rank 0 only holds the rows of its own band
        v = max(1, min(v, rows_available))
    got = got_u8_crop_fn(v)
    d = np.abs(got.astype(np.int16) - want[:, :v, :v].astype(np.int16))
    return {"region": f"top-left {v}x{v} px of the timed output", "max_abs_u8_diff": int(d.max()),
            "frac_off_by_one": float((d == 1).mean()), "argmax_agreement": float((got.argmax(0) == want[:, :v, :v].argmax(0)).mean()),
            "oracle_tiles": int((-(-cells.shape[-1] // sub)) ** 2)}


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def bcast(obj, world):
    if world == 1:
        return obj
    import torch.distributed as dist
    box = [obj]
    dist.broadcast_object_list(box, src=0)
    return box[0]


def gather_objs(obj, world):
    if world == 1:
        return [obj]
    import torch.distributed as dist
    out = [None] * world
    dist.all_gather_object(out, obj)
    return out


# ------------------------------------------------------------------------------------------------------------------
def run_workload(name: str, args, steps: int, warmup: int, main: bool, reuse=None):
    import torch
    from unmicst_b200._lib import UMX_U16
    from unmicst_b200.engine import Engine, calibrate, sample_probe_tiles, split_tile_rows, tile_geometry

    rank, world, local = dist_env()
    wl = WORKLOADS[name]
    model_name, H, W, planes, scale, cfg = wl["model"], wl["H"], wl["W"], wl["planes"], wl["scale"], wl["cfg"]
    if args.size and main:
        H = W = args.size
    IH, IW = (int(H * scale), int(W * scale)) if scale != 1.0 else (H, W)
    infer_shape = (IH, IW) if scale != 1.0 else None
    img = make_image(name, H, W)
    premap = premap_for(model_name, img)
    from unmicst_b200.modelzoo import KNOWN_HP
    S, C = KNOWN_HP[model_name]["imSize"], KNOWN_HP[model_name]["nChannels"]
    probe = None

    def probe_of(model):
        nonlocal probe
        if probe is None:
            probe = sample_probe_tiles(img, S, C, model.mean, model.std, premap, n=64, infer_shape=infer_shape)
        return probe

    # ---- rank 0 picks the weights' steepness and the arithmetic; every rank then builds the identical engine
    plan = None
    if rank == 0:
      try:
        host_threads(world)
        if reuse and reuse.get("model_name") == model_name:
            model, gain, prec, terms, auto = reuse["model"], reuse["gain"], reuse["prec"], reuse["terms"], reuse["auto"]
            probe_of(model)
        else:
            model, gain = choose_model(model_name, probe_of)
            prec, terms, auto = args.precision, {}, None
            if prec == "auto":
                prec, terms, auto = calibrate(model, local, probe_of(model))
            elif prec == "default":
                prec = "split3"
        plan = dict(gain=gain, prec=prec, terms=terms)
      except Exception as ex:              # every rank must leave this workload together
        plan = {"error": f"{type(ex).__name__}: {ex}"[:300]}
    plan = bcast(plan, world)
    if "error" in plan:
        raise RuntimeError(plan["error"])
    if rank != 0:
        from unmicst_b200 import modelzoo
        d = os.path.join(ROOT, "tests", "golden", "models", model_name)
        model = modelzoo.load_model(d) if plan["gain"] is None else modelzoo.synthetic_model(model_name, seed=0, logit_gain=plan["gain"])
        auto = None
    gain, prec, terms = plan["gain"], plan["prec"], plan["terms"]
    eng = Engine(model, device=local, precision=prec, max_batch_tiles=args.max_batch, op_terms=terms)
    K = eng.K

    # ---- parity of the timed engine against the oracle on tiles of this image
    parity, want = None, None
    if rank == 0:
      try:
        taps = {}
        want = oracle_forward(model, taps)(probe)
        got = eng.forward_tiles(probe)
        parity = {"vs": "oracle.unet_oracle.forward (torch-CPU fp32) on 64 tiles of this image (corners/edges/interior)",
                  "max_abs_dp": float(np.abs(got - want).max()), "argmax_agreement": float((got.argmax(-1) == want.argmax(-1)).mean()),
                  "max_abs_logit": float(np.abs(taps["logits"]).max()), "tolerance": 2e-3, "argmax_required": 0.999,
                  "within_contract": bool(np.abs(got - want).max() <= 2e-3 and (got.argmax(-1) == want.argmax(-1)).mean() >= 0.999),
                  "auto": auto}
      except Exception as ex:              # rank-0-only work must not strand the other ranks at the next barrier
        parity = {"error": f"{type(ex).__name__}: {ex}"[:300], "max_abs_dp": None, "argmax_agreement": None, "max_abs_logit": None,
                  "within_contract": None, "auto": auto}

    _, sub, npr, npc = tile_geometry(IH, IW, S)
    bands = split_tile_rows(npr, world)
    band = bands[rank] if rank < len(bands) else None
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)
    cli_quant = scale != 1.0                       # pages back at the raw size, quantised twice (UnMicst1-5.py:848-853)
    OH, OW = (H, W)
    if not cli_quant:
        OH, OW = IH, IW
    d_img = torch.from_numpy(img.view(np.int16)).cuda()
    d_out = torch.zeros((K, OH, OW), dtype=torch.uint8, device="cuda")
    h_img = torch.from_numpy(img.view(np.int16)).pin_memory()
    h_out = torch.zeros((K, OH, OW), dtype=torch.uint8).pin_memory()
    if band:
        r0, r1 = eng.band_out_rows(IH, H, band) if cli_quant else eng.band_rows(IH, band)
    else:
        r0 = r1 = 0

    def step(img_ptr, out_ptr, e=eng, b=band):
        if b:
            e.infer_ptr(img_ptr, UMX_U16, planes, H, W, H * W, model.mean, model.std, out_u8_ptr=out_ptr, tile_rows=b,
                        premap=premap, infer_shape=infer_shape, cli_quant=cli_quant)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        """n steps between barriers; returns (max over ranks, this rank's own) milliseconds."""
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(n):
            fn()
        b.record(stream)
        torch.cuda.synchronize()
        own = a.elapsed_time(b)
        barrier()
        ms = own
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([own], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, own

    resident = lambda: step(d_img.data_ptr(), d_out.data_ptr())
    e2e = lambda: step(h_img.data_ptr(), h_out.data_ptr())
    for _ in range(warmup):
        resident()
    n_before = eng.launch_count
    torch.cuda.nvtx.range_push("timed_resident")          # ncu --nvtx --nvtx-include "timed_resident/" profiles exactly these launches
    with ClockSampler(local, enabled=rank == 0) as clk:
        ms, own = timed(resident, steps)
    torch.cuda.nvtx.range_pop()
    launches = eng.launch_count - n_before
    ms_per_step = ms / steps
    value = IH * IW / 1e6 / (ms_per_step / 1e3)
    tile_rows_computed = (band[1] - band[0] + (1 if band[0] > 0 else 0)) if band else 0
    per_rank = gather_objs({"rank": rank, "ms_per_step": own / steps, "tile_rows": tile_rows_computed,
                            "tiles": tile_rows_computed * npc, "seam_tiles": npc if band and band[0] > 0 else 0,
                            "launch_groups": -(-tile_rows_computed // max(1, (args.max_batch or 4096) // npc)) if band else 0}, world)

    # ---- end to end through host buffers
    e2e()
    e2e_steps = max(1, min(steps, 3))
    ms_e2e = timed(e2e, e2e_steps)[0] / e2e_steps
    in_rows = 0
    if band:
        lo_t = max(band[0] - 1, 0)
        in_rows = min(IH, (band[1] - 1) * sub + S) - max(0, lo_t * sub - S // 8)
        in_rows = min(H, int(np.ceil(in_rows / scale)) + (4 if scale != 1.0 else 0))
    h2d = planes * in_rows * W * 2
    d2h = K * (r1 - r0) * OW

    # ---- stitched uint8: crop of the timed (resident) output vs the oracle pipeline
    crop_check = None
    if rank == 0 and args.crop_check:
        try:
            crop_check = stitched_crop_check(model, img, premap, scale, lambda v: d_out[:, :v, :v].cpu().numpy(),
                                             rows_available=r1 if world > 1 else 0)
        except Exception as ex:
            crop_check = {"error": f"{type(ex).__name__}: {ex}"[:300]}

    # ---- multi-GPU: every band of the e2e output, bit for bit, against ONE GPU doing the whole slide
    bands_ok = None
    if world > 1:
        mine = hashlib.sha1(h_out[:, r0:r1].numpy().tobytes()).hexdigest() if band else None
        rows_all = gather_objs((r0, r1, mine), world)
        if rank == 0:
            d_full = torch.zeros((K, OH, OW), dtype=torch.uint8, device="cuda")
            eng.infer_ptr(d_img.data_ptr(), UMX_U16, planes, H, W, H * W, model.mean, model.std, out_u8_ptr=d_full.data_ptr(),
                          premap=premap, infer_shape=infer_shape, cli_quant=cli_quant)
            full = d_full.cpu().numpy()
            del d_full
            covered = np.zeros(OH, dtype=np.int32)
            bands_ok = True
            for a, b, hsh in rows_all:
                if hsh is None:
                    continue
                covered[a:b] += 1
                bands_ok = bands_ok and hashlib.sha1(full[:, a:b].tobytes()).hexdigest() == hsh
            bands_ok = bool(bands_ok and (covered == 1).all())

    # ---- roofline (separate profiled pass: CUDA events around every launch on the launching stream)
    eng.profile_enable(True)
    resident()
    prof = eng.profile_read()
    eng.profile_enable(False)
    peaks = load_peaks()
    total_ms = sum(p["ms"] for p in prof) or 1.0
    dom = max(prof, key=lambda p: p["ms"])
    if dom["flops"] > 0:
        bound, achieved, peak, unit = "tensor", dom["flops"] / (dom["ms"] * 1e-3) / 1e12, peaks["tc_sustained"], "TFLOP/s"
    else:
        bound, achieved, peak, unit = "hbm", dom["bytes"] / (dom["ms"] * 1e-3) / 1e9, peaks["hbm"], "GB/s"
    by_hbm = dom["bytes"] / (dom["ms"] * 1e-3) / 1e9 / peaks["hbm"] if dom["ms"] else 0
    if bound == "tensor" and by_hbm > achieved / peak:      # the binding roofline term is HBM for this layer
        bound, achieved, peak, unit = "hbm", by_hbm * peaks["hbm"], peaks["hbm"], "GB/s"
    traffic = None          # DRAM bytes per launch of the dominant kernel: ncu per-tile figure x this run's tiles per launch
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        ent = json.load(open(tp)).get(dom["name"])
        if isinstance(ent, dict) and dom["launches"]:
            traffic = ent["dram_bytes_per_tile"] * tile_rows_computed * npc / dom["launches"]
    # per-layer roofline of the whole step: sum over kernels of max(FLOPs / tensor peak, bytes / HBM peak)
    roof_ms = sum(max(p["flops"] / (peaks["tc_sustained"] * 1e12), p["bytes"] / (peaks["hbm"] * 1e9)) * 1e3 for p in prof)
    # MMAs issued per algorithmic MMA: 1 + the correction terms of each source, weighted by its share of the contraction
    # (3 for the full hi/lo split): `frac` charges algorithmic FLOPs only, `frac_issued` what the tensor pipe executes
    from unmicst_b200.engine import tensor_op_sources
    factor = {}
    for _, nm, ks in tensor_op_sources(model):
        t = {"single": [0, 0], "split3": [3, 3], "fp32": [0, 0]}.get(prec) or (auto or {}).get("op_terms", {}).get(nm, [3, 3])
        k0, k1 = ks[0], (ks[1] if len(ks) > 1 else 0.0)
        factor[nm] = 1.0 + (k0 * bin(t[0]).count("1") + k1 * bin(t[1] if len(ks) > 1 else 0).count("1")) / (k0 + k1)
    fac = lambda name: factor.get(name.split("+")[0], 1.0)
    roof_issued_ms = sum(max(p["flops"] * fac(p["name"]) / (peaks["tc_sustained"] * 1e12), p["bytes"] / (peaks["hbm"] * 1e9)) * 1e3 for p in prof)
    roofline = {"kernel": dom["name"], "bound": bound, "achieved": achieved, "peak": peak, "unit": unit,
                "frac": achieved / peak, "mma_issue_factor": fac(dom["name"]) if bound == "tensor" else None,
                "frac_issued": achieved * fac(dom["name"]) / peak if bound == "tensor" else achieved / peak,
                "traffic": traffic, "peak_source": peaks["source"],
                "step": {"per_layer_roofline_ms": roof_ms, "measured_ms": total_ms, "frac": roof_ms / total_ms if total_ms else None,
                         "per_layer_roofline_issued_ms": roof_issued_ms, "frac_issued": roof_issued_ms / total_ms if total_ms else None},
                "avg_launch_ms": dom["ms"] / max(1, dom["launches"]), "share_of_step": dom["ms"] / total_ms,
                "kernels": [{"name": p["name"], "ms": round(p["ms"], 3), "launches": p["launches"], "mma_x": round(fac(p["name"]), 3),
                             "tflops": round(p["flops"] / (p["ms"] * 1e-3) / 1e12, 2) if p["ms"] else 0,
                             "gbs": round(p["bytes"] / (p["ms"] * 1e-3) / 1e9, 1) if p["ms"] else 0} for p in prof if p["launches"]]}

    # ---- the same slide in the two plain modes
    modes, full_check = None, None
    if main and args.modes:
        modes = {}
        d_alt = torch.zeros_like(d_out)
        for mode in ("single", "split3"):
            if mode == prec:
                modes[mode] = {"value": value, "ms_per_step": ms_per_step}
                continue
            e2 = Engine(model, device=local, precision=mode, max_batch_tiles=args.max_batch)
            e2.set_stream(stream.cuda_stream)
            fn = lambda: step(d_img.data_ptr(), d_alt.data_ptr(), e=e2)
            fn()
            t_ms = timed(fn, 2)[0] / 2
            entry = {"value": IH * IW / 1e6 / (t_ms / 1e3), "ms_per_step": t_ms}
            if rank == 0 and want is not None:
                g2 = e2.forward_tiles(probe)
                entry["max_abs_dp_vs_oracle"] = float(np.abs(g2 - want).max())
            modes[mode] = entry
            e2.close()
            if mode == "split3" and band:
                # the WHOLE timed output against the full hi/lo split (itself within ~5e-5 of the oracle on the probe
                # tiles): every stitched uint8 value of this rank's band, compared on the device
                a, b = d_out[:, r0:r1].to(torch.int16), d_alt[:, r0:r1].to(torch.int16)
                diff = (a - b).abs()
                mine = (int(diff.max().item()), int((diff > 0).sum().item()), int(diff.numel()))
                del a, b, diff
            elif mode == "split3":
                mine = (0, 0, 0)
            if mode == "split3":
                allr = gather_objs(mine, world)
                if rank == 0:
                    n = max(1, sum(x[2] for x in allr))
                    full_check = {"vs": "the split3 engine (3 MMAs per product everywhere) over the whole slide, all ranks",
                                  "max_abs_u8_diff": max(x[0] for x in allr), "frac_differing": sum(x[1] for x in allr) / n,
                                  "values_compared": n, "split3_max_abs_dp_vs_oracle_on_probe": entry.get("max_abs_dp_vs_oracle")}
        del d_alt

    out = None
    if rank == 0:
        dtype = {"fp32": "f32", "split3": "f16 hi/lo split x3 MMA, f32 accumulate", "single": "f16, f32 accumulate",
                 "mixed": "f16 (per layer and concat source: 1 MMA, or hi/lo split with 1-2 correction MMAs), f32 accumulate"}[prec]
        out = {
            "metric": "megapixels/sec of K-class probability map", "value": value, "unit": "MP/s",
            "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
            "config": {"workload": cfg, "model": model_name,
                       "weights": (f"synthetic seed 0, last linear map x{gain:.2f} so that max|logit| = {parity['max_abs_logit'] or float('nan'):.1f} on this image "
                                   "(real checkpoint not shipped)") if gain is not None else "real checkpoint (tests/golden/models)",
                       "H": IH, "W": IW, "raw_H": H, "raw_W": W, "scaling_factor": scale, "tiles": npr * npc, "tile": S,
                       "precision": prec, "precision_requested": args.precision,
                       "single_mma_layers": (auto or {}).get("single_layers") if prec == "mixed" else None,
                       "partial_split_layers": {k: v for k, v in (auto or {}).get("op_terms", {}).items() if v not in ([0, 0], [3, 3])} if prec == "mixed" else None,
                       "l2": "inputs + activations per step >> 126 MB L2, no explicit flush",
                       "parallelism": f"tile-row bands x{world} balanced by computed rows (own + seam), no collective"},
            "clocks": clk.summary(),
            "e2e": {"value": IH * IW / 1e6 / (ms_e2e / 1e3), "unit": "MP/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e},
            "gpu_launches": launches,
            "roofline": roofline,
            "parity": parity,
            "stitched_u8": crop_check,
            "stitched_u8_whole_slide": full_check,
            "per_rank": per_rank,
            "bands_bit_exact": bands_ok,
            "modes": modes,
        }
        if scale != 1.0:
            out["value_raw_px"] = H * W / 1e6 / (ms_per_step / 1e3)
            out["e2e"]["value_raw_px"] = H * W / 1e6 / (ms_e2e / 1e3)
    eng.close()
    del d_img, d_out, h_img, h_out
    torch.cuda.empty_cache()
    ctx = dict(model_name=model_name, model=model, gain=gain, prec=prec, terms=terms, auto=auto) if rank == 0 else None
    return out, ctx, img


def run_cores(args, reuse):
    """configs[4] as the reference's batch script sees it (batchUNet2DTMACycif.py:539-569): 60 separate TMA cores of
    3072 x 3072 px, here through umx_infer_images (tiles of several cores share every network launch), cores dealt
    round-robin to the ranks.  Host buffers in and out: the number is end to end."""
    import torch
    from unmicst_b200 import modelzoo
    from unmicst_b200.engine import Engine, PreMap
    rank, world, local = dist_env()
    n_cores, side, distinct = 60, 3072, 4
    plan = bcast(dict(gain=reuse["gain"], prec=reuse["prec"], terms=reuse["terms"]) if rank == 0 else None, world)
    model = reuse["model"] if rank == 0 else modelzoo.synthetic_model("CytoplasmIncell2", seed=0, logit_gain=plan["gain"])
    base = [synthetic_dna(side, side, seed=100 + i) for i in range(distinct)]
    mine = [base[i % distinct] for i in range(rank, n_cores, world)]
    pms = [PreMap(in_scale=1.0 / 65535, rescale=True, imin=float(c.min()) / 65535, imax=float(c.max()) / 65535) for c in mine]
    eng = Engine(model, device=local, precision=plan["prec"], max_batch_tiles=args.max_batch, op_terms=plan["terms"])

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    eng.infer_images(mine[:2], pms[:2])                       # warm-up
    n0 = eng.launch_count
    barrier()
    t0 = time.perf_counter()
    outs = eng.infer_images(mine, pms)
    own = time.perf_counter() - t0
    barrier()
    launches = eng.launch_count - n0
    secs = max(gather_objs(own, world))
    same = None
    if rank == 0:
        one, _ = eng.infer_image(mine[0], premap=pms[0])
        same = bool(np.array_equal(one, outs[0]))
    eng.close()
    if rank != 0:
        return None
    S = model.hp["imSize"]
    sub = S - 2 * (S // 8)
    tiles_per_core = (-(-side // sub)) ** 2
    return {"workload": "configs[4] as 60 separate 3072 x 3072 TMA cores through umx_infer_images (many images per launch), end to end from host buffers",
            "model": "CytoplasmIncell2", "n_gpus": world, "precision": plan["prec"], "cores": n_cores, "tiles_per_core": tiles_per_core,
            "e2e": {"value": n_cores * side * side / 1e6 / secs, "unit": "MP/s", "seconds": secs,
                    "h2d_bytes_per_step": n_cores * side * side * 2, "d2h_bytes_per_step": n_cores * side * side * model.hp["nClasses"]},
            "value": n_cores * side * side / 1e6 / secs, "unit": "MP/s",
            "gpu_launches_rank0": launches,
            "equals_single_image_call_bit_for_bit": same}


def compact(o):
    r = o["roofline"]
    return {"workload": o["config"]["workload"], "model": o["config"]["model"], "weights": o["config"]["weights"],
            "H": o["config"]["H"], "W": o["config"]["W"], "tiles": o["config"]["tiles"], "n_gpus": o["n_gpus"], "steps": o["steps"],
            "precision": o["config"]["precision"], "single_mma_layers": o["config"]["single_mma_layers"],
            "partial_split_layers": o["config"]["partial_split_layers"],
            "value": o["value"], "unit": "MP/s", "ms_per_step": o["ms_per_step"], "value_raw_px": o.get("value_raw_px"),
            "e2e": o["e2e"], "roofline": {"kernel": r["kernel"], "bound": r["bound"], "frac": r["frac"], "step": r["step"]},
            "parity": {k: o["parity"][k] for k in ("max_abs_dp", "argmax_agreement", "max_abs_logit", "within_contract")},
            "stitched_u8": o["stitched_u8"], "bands_bit_exact": o["bands_bit_exact"], "per_rank": o["per_rank"]}


def run_ours(args):
    import torch
    rank, world, local = dist_env()
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    out, ctx, img = run_workload(args.workload, args, args.steps, args.warmup, main=True)
    if rank == 0 and args.cpu_budget > 0:
        try:
            wl = WORKLOADS[args.workload]
            n_side = tiles_side_for_budget(ctx["model"], args.cpu_budget)
            out["cpu_baseline"] = cpu_baseline(ctx["model"], img, wl["scale"], n_side)
        except Exception as ex:
            out["cpu_baseline"] = {"error": f"{type(ex).__name__}: {ex}"[:300]}
    del img
    names = [] if args.configs == "none" else (DEFAULT_CONFIGS if args.configs == "all" else args.configs.split(","))
    configs = []
    cyto_ctx = None
    for name in names:
        if name == args.workload or (name not in WORKLOADS and name != "cyto2cores"):
            continue
        t0 = time.perf_counter()
        try:
            if name == "cyto2cores":
                have = bcast(cyto_ctx is not None if rank == 0 else None, world)
                if not have:
                    continue                                  # needs the calibration of the cyto2tma entry
                c = run_cores(args, cyto_ctx)
                if rank == 0:
                    c["wall_s"] = round(time.perf_counter() - t0, 1)
                    configs.append(c)
                continue
            o, c_ctx, _ = run_workload(name, args, max(1, min(3, args.steps)), 1, main=False, reuse=ctx)
            if name == "cyto2tma":
                cyto_ctx = c_ctx
            if rank == 0:
                c = compact(o)
                c["wall_s"] = round(time.perf_counter() - t0, 1)
                configs.append(c)
        except Exception as ex:          # a side config must not take the headline down with it
            if rank == 0:
                configs.append({"workload": WORKLOADS[name]["cfg"], "error": f"{type(ex).__name__}: {ex}"[:300]})
    if rank == 0:
        out["configs"] = configs
        print(json.dumps(out))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def run_reference(args):
    """--impl reference: the reference's own CPU path for this metric, i.e. (TensorFlow being uninstallable) the oracle
    port on all host cores, one bounded sample of the workload per step - the same fixed top-left tile sample as the
    cpu_baseline leg of the main arm, sized so that K + W steps end within a few minutes."""
    rank, world, local = dist_env()
    if rank != 0:
        return
    from unmicst_b200.engine import sample_probe_tiles
    wl = WORKLOADS[args.workload]
    model_name, H, W, scale = wl["model"], wl["H"], wl["W"], wl["scale"]
    if args.size:
        H = W = args.size
    n_steps = max(1, args.steps + args.warmup)
    budget = max(2.0, min(20.0, 170.0 / n_steps))
    from unmicst_b200.modelzoo import KNOWN_HP
    hp = KNOWN_HP[model_name]
    sub = hp["imSize"] - 2 * (hp["imSize"] // 8)
    side_px = int(min(H, np.ceil(64 * sub / scale)))
    img = make_image(args.workload, min(H, side_px), min(W, side_px)) if not wl.get("file") else make_image(args.workload, H, W)
    premap = premap_for(model_name, img)
    probe = lambda m: sample_probe_tiles(img, hp["imSize"], hp["nChannels"], m.mean, m.std, premap, n=16,
                                         infer_shape=(int(img.shape[-2] * scale), int(img.shape[-1] * scale)) if scale != 1.0 else None)
    model, gain = choose_model(model_name, probe)
    n_side = tiles_side_for_budget(model, budget)
    vals, times, last = [], [], None
    for i in range(n_steps):
        t0 = time.perf_counter()
        last = cpu_baseline(model, img, scale, n_side)
        if i >= args.warmup:
            vals.append(last["value"])
            times.append(time.perf_counter() - t0)
    v = float(np.mean(vals)) if vals else last["value"]
    last["value"] = v
    print(json.dumps({
        "impl": "reference", "metric": "megapixels/sec of K-class probability map", "value": v, "unit": "MP/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * float(np.mean(times)) if times else None, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["cfg"], "model": model_name, "H": int(H * scale), "W": int(W * scale),
                   "weights": "real checkpoint" if gain is None else f"synthetic seed 0, last linear map x{gain:.2f}",
                   "sample_tiles": last["tiles"], "note": "per step: a fixed top-left tile sample of the workload, extrapolated linearly in pixels"},
        "cpu_baseline": last, "e2e": {"value": v, "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="solo20k", choices=sorted(WORKLOADS))
    ap.add_argument("--size", type=int, default=0, help="override H=W of the main workload (debug)")
    ap.add_argument("--precision", default="auto", choices=["auto", "default", "fp32", "split3", "single"])
    ap.add_argument("--max-batch", type=int, default=0)
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU-baseline work (0 = skip)")
    ap.add_argument("--configs", default="all", help="other BASELINE configs to append: all | none | comma list")
    ap.add_argument("--no-modes", dest="modes", action="store_false", help="skip the plain single / split3 passes")
    ap.add_argument("--no-crop-check", dest="crop_check", action="store_false", help="skip the stitched-uint8 oracle check")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
