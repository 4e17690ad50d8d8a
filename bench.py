#!/usr/bin/env python
"""Benchmark of the UnMicst probability-map hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload solo20k|duo4k|cyto2tma|legacy20k] [--size PX] [--precision P]

A "step" = one pass of the hot path (gather/normalise -> UNet -> stitch -> uint8) over one
synthetic slide.  Default workload = BASELINE.json configs[2]: unmicst-solo graph
(nucleiDAPI1-5 shapes, seeded synthetic weights — the real checkpoint is not shipped),
20 000 x 20 000 px synthetic DNA image, tile rows sharded over the N GPUs (strong scaling,
no collective).  Prints ONE JSON line (rank 0).
  value     megapixels/s with the image and the outputs resident in HBM
  e2e       same metric through the public API with pinned HOST buffers (H2D + D2H inside)
  roofline  dominant kernel, algorithmic FLOPs / CUDA-event time vs MEASURED_PEAKS.json
  cpu_baseline  the oracle port (torch-CPU fp32 + Python PI2D loop) on a bounded tile sample
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (model, H, W, planes, BASELINE.json config)
    "solo20k": ("nucleiDAPI1-5", 20000, 20000, 1, "configs[2] unmicst-solo whole-slide 20k x 20k"),
    "solo40k": ("nucleiDAPI1-5", 40000, 40000, 1, "configs[3] unmicst-solo scalingFactor 2.0 (inference at 40k x 40k)"),
    "duo4k": ("nucleiDAPILAMIN", 4096, 4096, 2, "configs[1] unmicst-duo 4k x 4k"),
    "cyto2tma": ("CytoplasmIncell2", 18432, 30720, 1, "configs[4] UnMicstCyto2 60-core TMA montage"),
    "legacy20k": ("nucleiDAPI", 20000, 20000, 1, "unmicst-legacy 20k x 20k (not a BASELINE config)"),
    "sample105": ("nucleiDAPI", 832, 960, 1, "configs[0] sample-sized image, legacy graph"),
}


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
    planes = WORKLOADS[workload][3]
    if planes == 1:
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

    def __init__(self, index: int):
        self.index = index
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
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(s[0]) for s in self.samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "samples": len(self.samples)}


def oracle_forward(model):
    from oracle import unet_oracle
    return lambda x: unet_oracle.forward(model.weights, model.hp, model.variant, x)


def cpu_baseline(model, image, budget_s: float = 15.0):
    """The oracle port — torch-CPU fp32 graph with the reference's batch size and its Python PI2D
    loop (one pass for all classes: flatters the reference up to 3x) — on a crop
    of the same image sized to ~budget_s seconds.  Returns MP/s of valid output pixels."""
    import torch
    from oracle import pi2d_oracle
    torch.set_num_threads(os.cpu_count() or 1)       # torchrun pins OMP_NUM_THREADS=1; the baseline gets every host core
    S, C, B = model.hp["imSize"], model.hp["nChannels"], model.hp["batchSize"]
    sub = S - 2 * (S // 8)
    fw = oracle_forward(model)
    x = np.zeros((B, S, S, C), np.float32)
    fw(x)                                      # warm-up
    t = time.perf_counter(); fw(x); per_batch = time.perf_counter() - t
    n_batches = max(1, int(budget_s / per_batch))
    side = max(1, int(np.sqrt(n_batches * B)))
    h = min(image.shape[-2], side * sub); w = min(image.shape[-1], side * sub)
    crop = image[..., :h, :w].astype(np.float64) * (1.0 / 65535)
    t = time.perf_counter()
    pm = pi2d_oracle.infer_image(crop, fw, S, C, model.mean, model.std, B, accum_dtype=np.float16)
    pi2d_oracle.quantize_u8(pm)
    dt = time.perf_counter() - t
    tiles = (-(-h // sub)) * (-(-w // sub))
    return {"value": h * w / dt / 1e6, "unit": "MP/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{h}x{w} px crop ({tiles} tiles, batch {B}) of the same image, torch-CPU fp32 oracle + Python PI2D loop, "
                      f"{dt:.1f} s; TensorFlow itself is not installable here"}


def load_bench_model(name: str):
    """Real checkpoint when the repository ships it (the two legacy fixtures), seeded stand-ins otherwise."""
    from unmicst_b200 import modelzoo
    d = os.path.join(ROOT, "tests", "golden", "models", name)
    if os.path.exists(os.path.join(d, "model.ckpt.data-00000-of-00001")):
        return modelzoo.load_model(d)
    return modelzoo.synthetic_model(name, seed=0)


def dist_setup(n_gpus: int):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def run_reference(args):
    rank, world, local = dist_setup(args.gpus)
    if rank != 0:
        return
    from unmicst_b200 import modelzoo
    model_name, H, W, planes, cfg = WORKLOADS[args.workload]
    if args.size:
        H = W = args.size
    model = load_bench_model(model_name)
    side = 2000
    img = make_image(args.workload, min(H, side), min(W, side))
    budget = max(5.0, min(60.0, 150.0 / max(1, args.steps + args.warmup)))
    vals, times = [], []
    last = None
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        last = cpu_baseline(model, img, budget_s=budget)
        if i >= args.warmup:
            vals.append(last["value"])
            times.append(time.perf_counter() - t0)
    v = float(np.mean(vals)) if vals else last["value"]
    last["value"] = v
    ms_step = 1e3 * float(np.mean(times)) if times else None
    print(json.dumps({
        "impl": "reference", "metric": "megapixels/sec of K-class probability map", "value": v, "unit": "MP/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": cfg, "model": model_name, "H": H, "W": W,
                   "weights": "synthetic seed 0" if model.synthetic else "real checkpoint"},
        "cpu_baseline": last, "e2e": {"value": v, "unit": "MP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_ours(args):
    import torch
    from unmicst_b200 import modelzoo
    from unmicst_b200._lib import UMX_U16
    from unmicst_b200.engine import Engine, split_tile_rows, tile_geometry

    rank, world, local = dist_setup(args.gpus)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    model_name, H, W, planes, cfg = WORKLOADS[args.workload]
    if args.size:
        H = W = args.size
    model = load_bench_model(model_name)
    torch.cuda.set_device(local)
    eng = Engine(model, device=local, precision=args.precision, max_batch_tiles=args.max_batch)
    precision = eng.precision
    S, K = eng.S, eng.K
    # parity of the timed configuration against this library's own fp32 CUDA-core path (itself checked
    # against the oracle in tests/): same seeded weights, random normal tiles
    parity = None
    if rank == 0:
        rng = np.random.default_rng(99)
        probe = rng.normal(size=(8, S, S, eng.C)).astype(np.float32)
        ref32 = Engine(model, device=local, precision="fp32", max_batch_tiles=64)
        a, b = eng.forward_tiles(probe), ref32.forward_tiles(probe)
        ref32.close()
        parity = {"max_abs_dp_vs_fp32_path": float(np.abs(a - b).max()),
                  "argmax_agreement": float((a.argmax(-1) == b.argmax(-1)).mean()), "tolerance": 2e-3,
                  "auto": eng.auto_report}
    _, sub, npr, npc = tile_geometry(H, W, S)
    band = split_tile_rows(npr, world)[rank] if rank < min(world, npr) else None
    img = make_image(args.workload, H, W)
    premap = None
    from unmicst_b200.engine import PreMap
    if model_name == "nucleiDAPI1-5":
        premap = PreMap(in_scale=1.0 / 65535)    # solo feeds img_as_float(u16) un-stretched (UnMicst1-5.py:816)
    else:                                        # the other tools stretch to (0, 0.983) (UnMicst.py:627-631)
        premap = PreMap(in_scale=1.0 / 65535, rescale=True, imin=float(img.min()) / 65535, imax=float(img.max()) / 65535)

    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)
    # ---- device-resident buffers (value) and pinned host buffers (e2e)
    d_img = torch.from_numpy(img.view(np.int16)).cuda()
    d_out = torch.empty((K, H, W), dtype=torch.uint8, device="cuda")
    h_img = torch.from_numpy(img.view(np.int16)).pin_memory()
    h_out = torch.empty((K, H, W), dtype=torch.uint8).pin_memory()
    r0, r1 = eng.band_rows(H, band) if band else (0, 0)

    def step_resident():
        if band:
            eng.infer_ptr(d_img.data_ptr(), UMX_U16, planes, H, W, H * W, model.mean, model.std,
                          out_u8_ptr=d_out.data_ptr(), tile_rows=band, premap=premap)

    def step_e2e():
        if band:
            eng.infer_ptr(h_img.data_ptr(), UMX_U16, planes, H, W, H * W, model.mean, model.std,
                          out_u8_ptr=h_out.data_ptr(), tile_rows=band, premap=premap)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(steps):
            fn()
        b.record(stream)
        barrier()
        ms = a.elapsed_time(b)
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    for _ in range(args.warmup):
        step_resident()
    n_before = eng.launch_count
    with ClockSampler(local) as clk:
        ms = timed(step_resident, args.steps)
    launches = eng.launch_count - n_before
    ms_per_step = ms / args.steps
    value = H * W / 1e6 / (ms_per_step / 1e3)

    # ---- end to end through host buffers
    step_e2e()
    e2e_steps = max(1, min(args.steps, 3))
    ms_e2e = timed(step_e2e, e2e_steps) / e2e_steps
    tile_rows_in = (band[1] - max(band[0] - 1, 0)) if band else 0
    in_rows = min(H, tile_rows_in * sub + S) if band else 0
    h2d = planes * in_rows * W * 2
    d2h = K * (r1 - r0) * W

    # ---- roofline of the dominant kernel (separate profiled pass: CUDA events around every launch)
    eng.profile_enable(True)
    step_resident()
    prof = eng.profile_read()
    eng.profile_enable(False)
    peaks = load_peaks()
    total_ms = sum(p["ms"] for p in prof) or 1.0
    dom = max(prof, key=lambda p: p["ms"])
    if dom["flops"] > 0:
        bound = "tensor"
        achieved = dom["flops"] / (dom["ms"] * 1e-3) / 1e12
        peak = peaks["tc_sustained"]
        unit = "TFLOP/s"
    else:
        bound = "hbm"
        achieved = dom["bytes"] / (dom["ms"] * 1e-3) / 1e9
        peak = peaks["hbm"]
        unit = "GB/s"
    by_hbm = dom["bytes"] / (dom["ms"] * 1e-3) / 1e9 / peaks["hbm"] if dom["ms"] else 0
    if bound == "tensor" and by_hbm > achieved / peak:      # the binding roofline term is HBM for this layer
        bound, achieved, peak, unit = "hbm", by_hbm * peaks["hbm"], peaks["hbm"], "GB/s"
    traffic = None          # DRAM bytes per launch of the dominant kernel: ncu per-tile figure x this run's tiles per launch
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        ent = json.load(open(tp)).get(dom["name"])
        if isinstance(ent, dict) and dom["launches"]:
            my_tiles = (band[1] - max(band[0] - 1, 0)) * npc if band else 0
            traffic = ent["dram_bytes_per_tile"] * my_tiles / dom["launches"]
    # per-layer roofline of the whole step: sum over kernels of max(FLOPs / tensor peak, bytes / HBM peak)
    # with the algorithmic FLOPs and 4-byte activation bytes the library accounts per launch
    roof_ms = sum(max(p["flops"] / (peaks["tc_sustained"] * 1e12), p["bytes"] / (peaks["hbm"] * 1e9)) * 1e3 for p in prof)
    roofline = {"kernel": dom["name"], "bound": bound, "achieved": achieved, "peak": peak, "unit": unit,
                "frac": achieved / peak, "traffic": traffic, "peak_source": peaks["source"],
                "step": {"per_layer_roofline_ms": roof_ms, "measured_ms": total_ms, "frac": roof_ms / total_ms if total_ms else None},
                "avg_launch_ms": dom["ms"] / max(1, dom["launches"]), "share_of_step": dom["ms"] / total_ms,
                "kernels": [{"name": p["name"], "ms": round(p["ms"], 3), "launches": p["launches"],
                             "tflops": round(p["flops"] / (p["ms"] * 1e-3) / 1e12, 2) if p["ms"] else 0,
                             "gbs": round(p["bytes"] / (p["ms"] * 1e-3) / 1e9, 1) if p["ms"] else 0} for p in prof]}

    out = None
    if rank == 0:
        cpu = cpu_baseline(model, img if planes == 1 else img, budget_s=args.cpu_budget) if args.cpu_budget > 0 else None
        out = {
            "metric": "megapixels/sec of K-class probability map", "value": value, "unit": "MP/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": {"fp32": "f32", "split3": "f16 hi/lo split x3 MMA, f32 accumulate", "single": "f16, f32 accumulate",
                      "mixed": "f16 (per layer: 1 MMA or hi/lo split x3), f32 accumulate"}[precision],
            "data": "synthetic",
            "config": {"workload": cfg, "model": model_name,
                       "weights": "synthetic seed 0 (real checkpoint not shipped)" if model.synthetic else "real checkpoint (tests/golden/models)",
                       "H": H, "W": W, "tiles": npr * npc, "tile": S, "precision": precision, "precision_requested": args.precision,
                       "l2": "inputs + activations per step >> 126 MB L2, no explicit flush",
                       "parallelism": f"tile-row bands x{world}, no collective"},
            "clocks": clk.summary(),
            "e2e": {"value": H * W / 1e6 / (ms_e2e / 1e3), "unit": "MP/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e},
            "gpu_launches": launches,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "parity": parity,
        }
        print(json.dumps(out))
    eng.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="solo20k", choices=sorted(WORKLOADS))
    ap.add_argument("--size", type=int, default=0, help="override H=W (debug)")
    ap.add_argument("--precision", default="auto", choices=["auto", "default", "fp32", "split3", "single"])
    ap.add_argument("--max-batch", type=int, default=0)
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU-baseline work (0 = skip)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
