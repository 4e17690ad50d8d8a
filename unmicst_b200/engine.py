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


def split_tile_rows(npr: int, parts: int) -> List[Tuple[int, int]]:
    """Contiguous tile-row bands balanced by tile count; empty bands dropped."""
    parts = max(1, min(parts, npr))
    base, extra = divmod(npr, parts)
    out, lo = [], 0
    for p in range(parts):
        hi = lo + base + (1 if p < extra else 0)
        out.append((lo, hi))
        lo = hi
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


AUTO_TOLERANCE = 5e-4     # max|dp| of fp16 single-pass vs split3 on the probe tiles (4x under the 2e-3 contract)


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


class Engine:
    """precision: 'default'/'split3' (fp16 hi/lo split, ~fp32 accurate, always within the 2e-3 contract),
    'single' (one fp16 MMA per product: ~3x less tensor work, accuracy depends on how steep the
    model's softmax is), 'mixed' (per layer: ``single_mask`` bit i = op i of ``profile_read`` runs single),
    'fp32' (CUDA cores only), or 'auto': calibrate on probe tiles against 'split3' — keep 'single' if it
    stays within AUTO_TOLERANCE, otherwise switch layers to single greedily (most expensive first) for as
    long as the probe stays within the tolerance ('mixed'), which may end at plain 'split3'."""

    def __init__(self, model: Model, device: int = 0, precision: str = "default", max_batch_tiles: int = 0,
                 probe_tiles: Optional[np.ndarray] = None, single_mask: int = 0):
        self.auto_report = None
        self.single_mask = 0
        if precision == "auto":
            probe_batch = 64
            ref = Engine(model, device, "split3", probe_batch)
            if probe_tiles is None:
                rng = np.random.default_rng(2024)
                probe_tiles = rng.normal(size=(16, ref.S, ref.S, ref.C)).astype(np.float32)
            want = ref.forward_tiles(probe_tiles)

            def err_of(prec: str, mask: int = 0) -> float:
                with Engine(model, device, prec, probe_batch, single_mask=mask) as e:
                    return float(np.abs(e.forward_tiles(probe_tiles) - want).max())

            d_single = err_of("single")
            report = {"single_vs_split3_max_abs_dp": d_single, "tolerance": AUTO_TOLERANCE}
            if d_single <= AUTO_TOLERANCE:
                chosen, mask = "single", 0
            else:
                # cost per op from one profiled pass of the split engine; candidates = tensor-path layers
                ref.profile_enable(True)
                ref.forward_tiles(probe_tiles)
                prof = ref.profile_read()
                ref.profile_enable(False)
                order = sorted((i for i, p in enumerate(prof) if p["flops"] > 0 and p["launches"] > 0 and i < 64),
                               key=lambda i: -prof[i]["ms"])
                mask, kept, d_mixed = 0, [], 0.0
                for i in order:
                    d = err_of("mixed", mask | (1 << i))
                    if d == d_mixed:
                        continue                              # the bit changes nothing: not a tensor-path layer
                    if d <= AUTO_TOLERANCE:
                        mask |= 1 << i
                        kept.append(prof[i]["name"])
                        d_mixed = d
                chosen = "mixed" if mask else "split3"
                report.update({"single_layers": kept, "mixed_vs_split3_max_abs_dp": d_mixed})
            ref.close()
            final = Engine(model, device, chosen, max_batch_tiles, single_mask=mask)
            self.__dict__.update(final.__dict__)
            final._h = None
            report["chosen"] = chosen
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

    def infer_ptr(self, img_ptr: int, dtype: int, n_planes: int, H: int, W: int, plane_stride: int,
                  mean: float, std: float, out_u8_ptr: int = 0, out_f32_ptr: int = 0,
                  tile_rows: Optional[Tuple[int, int]] = None, premap: Optional[PreMap] = None,
                  out_plane_stride: int = 0, out_row_base: int = 0, precision: str = "default",
                  no_sync: bool = False) -> None:
        """Raw-pointer form (host pageable / pinned or device memory — the library detects which)."""
        o = _lib.umx_opts()
        if tile_rows:
            o.tile_row0, o.tile_row1 = int(tile_rows[0]), int(tile_rows[1])
        o.precision = _lib.PRECISIONS[precision]
        o.flags = _lib.UMX_F_NO_SYNC if no_sync else 0
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
                    out_f32: Optional[np.ndarray] = None, precision: str = "default"):
        """image [H,W] or [C,H,W] (uint8/uint16/float32/float64) -> (u8 [K,H,W] | None, f32 [K,H,W] | None).

        With ``tile_rows`` only the image rows of that band are written (into full-size outputs)."""
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
        if want_u8 and out_u8 is None:
            out_u8 = np.zeros((self.K, H, W), dtype=np.uint8)
        if want_f32 and out_f32 is None:
            out_f32 = np.zeros((self.K, H, W), dtype=np.float32)
        self.infer_ptr(img.ctypes.data, _NP2UMX[img.dtype], planes, H, W, H * W, mean, std,
                       out_u8.ctypes.data if out_u8 is not None else 0,
                       out_f32.ctypes.data if out_f32 is not None else 0,
                       tile_rows=tile_rows, premap=premap, precision=precision)
        return out_u8, out_f32

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
    bit-identical to the 1-GPU result and no device-to-device exchange exists."""

    def __init__(self, model: Model, devices: Sequence[int], precision: str = "default", max_batch_tiles: int = 0):
        if not devices:
            raise ValueError("MultiEngine needs at least one device")
        self.engines = [Engine(model, d, precision, max_batch_tiles) for d in devices]
        self.model = model
        self.S, self.C, self.K = self.engines[0].S, self.engines[0].C, self.engines[0].K

    def close(self) -> None:
        for e in self.engines:
            e.close()

    def infer_image(self, image: np.ndarray, mean=None, std=None, premap=None, want_u8=True, want_f32=False,
                    precision: str = "default"):
        img = np.ascontiguousarray(image)
        H, W = img.shape[-2:]
        _, _, npr, _ = tile_geometry(H, W, self.S)
        bands = split_tile_rows(npr, len(self.engines))
        out_u8 = np.zeros((self.K, H, W), dtype=np.uint8) if want_u8 else None
        out_f32 = np.zeros((self.K, H, W), dtype=np.float32) if want_f32 else None
        errs: List[BaseException] = []

        def work(eng: Engine, band):
            try:
                eng.infer_image(img, mean, std, premap, want_u8, want_f32, tile_rows=band, out_u8=out_u8,
                                out_f32=out_f32, precision=precision)
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
