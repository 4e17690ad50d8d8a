"""world_size-2 gloo test (CPU): the tile-row band decomposition used for multi-GPU runs.
Each rank stitches only the tiles of its band (plus the one seam tile row above it) with the
oracle primitives; the gathered bands must partition the image rows and reproduce the
whole-image result exactly — no data exchange between ranks other than the final gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import pi2d_oracle
from unmicst_b200.engine import _resize_window, band_out_rows_py, band_rows_py, split_tile_rows, tile_geometry

S = 64
H, W = 333, 250


def _fake_network(x):
    """A deterministic per-tile 'network' with 3 outputs that depends on position inside the tile."""
    yy = np.linspace(0, 1, S, dtype=np.float32)[None, :, None, None]
    a = 1 / (1 + np.exp(-x * (1 + yy)))
    b = 1 / (1 + np.exp(x[:, ::-1] * 0.5))
    p = np.concatenate([a, b, 2 - a - b], axis=-1)
    return (p / p.sum(-1, keepdims=True)).astype(np.float32)


def _band(image, band):
    g = pi2d_oracle.tile_grid(H, W, S, S // 8)
    frame = pi2d_oracle.pad_frame(image, g)
    w = pi2d_oracle.ramp_weight(S, g.margin)
    num = np.zeros((3, g.frame_rows, g.frame_cols))
    cnt = np.zeros((g.frame_rows, g.frame_cols))
    first = max(band[0] - 1, 0)
    for t in range(first * g.npc, band[1] * g.npc):
        r0, c0 = g.origin(t)
        tile = ((pi2d_oracle.cut_tile(frame, g, t) - 0.3) / 0.2).astype(np.float32)[None, :, :, None]
        p = _fake_network(tile)[0]
        cnt[r0:r0 + S, c0:c0 + S] += w
        for k in range(3):
            num[k, r0:r0 + S, c0:c0 + S] += p[:, :, k] * w
    r0, r1 = band_rows_py(H, S, band)
    m = g.margin
    return r0, r1, num[:, m + r0:m + r1, m:m + W] / cnt[m + r0:m + r1, m:m + W]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    image = rng.random((H, W))
    _, _, npr, _ = tile_geometry(H, W, S)
    band = split_tile_rows(npr, world)[rank]
    r0, r1, out = _band(image, band)
    rows = torch.tensor([r0, r1])
    all_rows = [torch.zeros(2, dtype=torch.long) for _ in range(world)]
    dist.all_gather(all_rows, rows)
    full = torch.zeros((3, H, W), dtype=torch.float64)
    full[:, r0:r1] = torch.from_numpy(out)
    dist.all_reduce(full)                      # bands are disjoint: the sum is the concatenation
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)   # the max-over-ranks timing reduction bench.py uses
    if rank == 0:
        q.put(([tuple(int(v) for v in r) for r in all_rows], full.numpy(), float(t.item())))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_bands_reproduce_whole_image():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    rows, full, tmax = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert rows[0][0] == 0 and rows[0][1] == rows[1][0] and rows[1][1] == H
    assert tmax == 2.0
    rng = np.random.default_rng(5)
    image = rng.random((H, W))
    whole = pi2d_oracle.infer_image(image, _fake_network, S, 1, 0.3, 0.2, 7, accum_dtype=np.float64)
    assert np.allclose(full, whole, atol=1e-12)


@pytest.mark.parametrize("H_,S_", [(832, 128), (20000, 64), (100, 64), (18432, 256)])
def test_band_rows_partition_every_split(H_, S_):
    _, _, npr, _ = tile_geometry(H_, 1, S_)
    for parts in (1, 2, 3, 8):
        bands = split_tile_rows(npr, parts)
        rows = [band_rows_py(H_, S_, b) for b in bands]
        assert rows[0][0] == 0 and rows[-1][1] == H_
        assert all(a[1] == b[0] for a, b in zip(rows, rows[1:]))


@pytest.mark.parametrize("raw_h,infer_h,S_", [(20000, 40000, 64), (4000, 2000, 64), (832, 1248, 128), (1000, 370, 64), (3072, 3072, 256)])
def test_resized_band_rows_partition_and_stay_inside_the_band(raw_h, infer_h, S_):
    """--scalingFactor with several GPUs: a band owns the raw-grid rows whose resize window (bilinear taps + Gaussian on
    shrink) lies inside the inference rows the band has: its own rows plus the halo above that the recomputed seam
    tile row completes.  Bands must tile the raw rows and never look below what they emit."""
    m, sub, npr, _ = tile_geometry(infer_h, 1, S_)
    for parts in (1, 2, 3, 8):
        bands = split_tile_rows(npr, parts)
        rows = [band_out_rows_py(infer_h, raw_h, S_, b) for b in bands]
        assert rows[0][0] == 0 and rows[-1][1] == raw_h
        assert all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
        for (ta, tb), (r0, r1) in zip(bands, rows):
            if r1 <= r0:
                continue
            have_lo = 0 if ta == 0 else (ta - 1) * sub + m               # rows completed by tile rows >= ta - 1
            have_hi = infer_h if tb == npr else tb * sub - m
            lo = min(_resize_window(y, infer_h, raw_h)[0] for y in (r0, r1 - 1))
            hi = max(_resize_window(y, infer_h, raw_h)[1] for y in (r0, r1 - 1))
            assert lo >= have_lo or ta == 0
            assert hi < have_hi or tb == npr
