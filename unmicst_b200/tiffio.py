"""Minimal TIFF / BigTIFF page reader and streaming BigTIFF writer.

Replaces, for the hot path's file formats only, what the reference gets from tifffile /
scikit-image: reading ONE channel page of a (OME-)TIFF (UnMicst1-5.py:794-797,
``skio.imread(img_num=...)`` / ``tifffile.imread(key=...)``: the page index in the main IFD
chain, SubIFD pyramids ignored) and writing uncompressed uint8 pages into a BigTIFF with
append semantics (``skimage.io.imsave(..., bigtiff=True, append=...)``, UnMicst1-5.py:852-862).
Strips and tiles, uncompressed / Deflate / LZW, with or without the horizontal predictor, are decoded here
(Deflate by zlib, LZW by the library's host-side decoder); anything else (JPEG, multi-sample) goes through PIL.
"""
from __future__ import annotations

import os
import struct
import zlib
from typing import Dict, List, Optional, Tuple

import numpy as np

_TYPE_FMT = {1: "B", 2: "c", 3: "H", 4: "I", 5: "II", 6: "b", 7: "B", 8: "h", 9: "i", 10: "ii", 11: "f", 12: "d",
             16: "Q", 17: "q", 18: "Q"}
_TYPE_SIZE = {1: 1, 2: 1, 3: 2, 4: 4, 5: 8, 6: 1, 7: 1, 8: 2, 9: 4, 10: 8, 11: 4, 12: 8, 16: 8, 17: 8, 18: 8}

T_WIDTH, T_LENGTH, T_BITS, T_COMPRESSION, T_PHOTOMETRIC, T_DESCRIPTION = 256, 257, 258, 259, 262, 270
T_STRIP_OFFSETS, T_SPP, T_ROWS_PER_STRIP, T_STRIP_COUNTS, T_PLANAR = 273, 277, 278, 279, 284
T_TILE_W, T_TILE_L, T_TILE_OFFSETS, T_TILE_COUNTS, T_SAMPLE_FORMAT = 322, 323, 324, 325, 339
T_PREDICTOR = 317
C_NONE, C_LZW, C_DEFLATE, C_DEFLATE_OLD = 1, 5, 8, 32946


class TiffError(ValueError):
    pass


class _Reader:
    def __init__(self, path: str):
        self.path = path
        self.f = open(path, "rb")
        head = self.f.read(16)
        if head[:2] == b"II":
            self.e = "<"
        elif head[:2] == b"MM":
            self.e = ">"
        else:
            raise TiffError(f"{path}: not a TIFF file")
        magic = struct.unpack(self.e + "H", head[2:4])[0]
        if magic == 42:
            self.big = False
            self.first_ifd = struct.unpack(self.e + "I", head[4:8])[0]
        elif magic == 43:
            self.big = True
            self.first_ifd = struct.unpack(self.e + "Q", head[8:16])[0]
        else:
            raise TiffError(f"{path}: bad TIFF magic {magic}")

    def close(self):
        self.f.close()

    def _read_ifd(self, off: int) -> Tuple[Dict[int, tuple], int]:
        f, e = self.f, self.e
        f.seek(off)
        if self.big:
            n = struct.unpack(e + "Q", f.read(8))[0]
            raw = f.read(n * 20 + 8)
            esz, cnt_fmt, val_len = 20, "Q", 8
        else:
            n = struct.unpack(e + "H", f.read(2))[0]
            raw = f.read(n * 12 + 4)
            esz, cnt_fmt, val_len = 12, "I", 4
        tags: Dict[int, tuple] = {}
        for i in range(n):
            ent = raw[i * esz:(i + 1) * esz]
            tag, typ = struct.unpack(e + "HH", ent[:4])
            count = struct.unpack(e + cnt_fmt, ent[4:4 + val_len])[0]
            if typ not in _TYPE_SIZE:
                continue
            nbytes = _TYPE_SIZE[typ] * count
            if nbytes <= val_len:
                data = ent[4 + val_len:4 + val_len + nbytes]
            else:
                ptr = struct.unpack(e + cnt_fmt, ent[4 + val_len:4 + 2 * val_len])[0]
                here = f.tell()
                f.seek(ptr)
                data = f.read(nbytes)
                f.seek(here)
            if typ == 2:
                tags[tag] = (data.rstrip(b"\0").decode("latin-1"),)
            elif typ in (5, 10):
                vals = struct.unpack(e + _TYPE_FMT[typ][0] * (2 * count), data)
                tags[tag] = tuple(vals[2 * j] / max(1, vals[2 * j + 1]) for j in range(count))
            else:
                tags[tag] = struct.unpack(e + _TYPE_FMT[typ] * count, data)
        nxt = struct.unpack(e + ("Q" if self.big else "I"), raw[n * esz:n * esz + (8 if self.big else 4)])[0]
        return tags, nxt

    def ifds(self):
        off, seen = self.first_ifd, set()
        while off and off not in seen:
            seen.add(off)
            tags, nxt = self._read_ifd(off)
            yield tags
            off = nxt

    def page_tags(self, page: int) -> Dict[int, tuple]:
        for i, tags in enumerate(self.ifds()):
            if i == page:
                return tags
        raise IndexError(f"{self.path}: page {page} out of range")

    def _segment(self, off: int, nbytes: int, comp: int, rows: int, cols: int, dt: np.dtype, predictor: int) -> np.ndarray:
        """One strip or tile -> [rows, cols] samples in native byte order."""
        self.f.seek(off)
        raw = self.f.read(nbytes)
        want = rows * cols * dt.itemsize
        if comp in (C_DEFLATE, C_DEFLATE_OLD):
            raw = zlib.decompress(raw)
        elif comp == C_LZW:
            from ._lib import lib
            out = np.empty(want, dtype=np.uint8)
            src = np.frombuffer(raw, dtype=np.uint8)
            n = lib().umx_tiff_lzw_decode(src.ctypes.data, src.size, out.ctypes.data, out.size)
            if n < 0:
                raise TiffError(f"{self.path}: corrupt LZW data")
            raw = out[:n].tobytes()
        if len(raw) < want:
            raise TiffError(f"{self.path}: segment holds {len(raw)} of {want} bytes")
        a = np.frombuffer(raw, dtype=dt, count=rows * cols).reshape(rows, cols).astype(dt.newbyteorder("="))
        if predictor == 2:       # horizontal differencing, modulo the sample width
            a = np.cumsum(a, axis=1, dtype=a.dtype)
        return a

    def read_page(self, page: int) -> Optional[np.ndarray]:
        """Decode a single-sample page (strips or tiles; uncompressed, Deflate or LZW; optional horizontal
        predictor), or None if the layout needs a full codec."""
        t = self.page_tags(page)
        w, h = t[T_WIDTH][0], t[T_LENGTH][0]
        bits = t.get(T_BITS, (1,))[0]
        spp = t.get(T_SPP, (1,))[0]
        comp = t.get(T_COMPRESSION, (1,))[0]
        fmt = t.get(T_SAMPLE_FORMAT, (1,))[0]
        predictor = t.get(T_PREDICTOR, (1,))[0]
        if comp not in (C_NONE, C_LZW, C_DEFLATE, C_DEFLATE_OLD) or spp != 1 or bits not in (8, 16, 32, 64):
            return None
        kind = {1: "u", 2: "i", 3: "f"}.get(fmt)
        if kind is None or (kind == "f" and bits < 32):
            return None
        if predictor not in (1, 2) or (predictor == 2 and kind == "f"):
            return None
        dt = np.dtype(f"{self.e}{kind}{bits // 8}")
        out = np.empty((h, w), dtype=dt.newbyteorder("="))
        if T_TILE_OFFSETS in t:
            tw, tl = t[T_TILE_W][0], t[T_TILE_L][0]
            offs, counts = t[T_TILE_OFFSETS], t.get(T_TILE_COUNTS)
            across = -(-w // tw)
            for i, off in enumerate(offs):
                r0, c0 = (i // across) * tl, (i % across) * tw
                if r0 >= h:
                    break
                nbytes = counts[i] if counts else tw * tl * dt.itemsize
                if nbytes == 0:                      # sparse (never written) tile
                    out[r0:r0 + tl, c0:c0 + tw] = 0
                    continue
                tile = self._segment(off, nbytes, comp, tl, tw, dt, predictor)
                rr, cc = min(tl, h - r0), min(tw, w - c0)
                out[r0:r0 + rr, c0:c0 + cc] = tile[:rr, :cc]
        else:
            rps = min(t.get(T_ROWS_PER_STRIP, (h,))[0], h)
            offs, counts = t[T_STRIP_OFFSETS], t.get(T_STRIP_COUNTS)
            for i, off in enumerate(offs):
                r0 = i * rps
                rr = min(rps, h - r0)
                nbytes = counts[i] if counts else rr * w * dt.itemsize
                out[r0:r0 + rr] = self._segment(off, nbytes, comp, rr, w, dt, predictor)
        return out


def count_pages(path: str) -> int:
    r = _Reader(path)
    try:
        return sum(1 for _ in r.ifds())
    finally:
        r.close()


def read_page(path: str, page: int = 0) -> np.ndarray:
    """One page (channel) of a TIFF/BigTIFF/OME-TIFF as a 2-D array in its stored dtype."""
    r = _Reader(path)
    try:
        arr = r.read_page(page)
    finally:
        r.close()
    if arr is not None:
        return arr
    from PIL import Image   # compressed / exotic layouts
    Image.MAX_IMAGE_PIXELS = None
    with Image.open(path) as im:
        im.seek(page)
        a = np.array(im)
    if a.dtype == np.int32 and im.mode.startswith("I;16"):
        a = a.astype(np.uint16)
    return a


class BigTiffWriter:
    """Uncompressed little-endian (Big)TIFF, one IFD per ``write_page`` call, rows written as they
    arrive (``write_rows``) so a page can be streamed band by band from the GPU."""

    def __init__(self, path: str, append: bool = False, bigtiff: bool = True):
        self.path = path
        self.big = bigtiff
        self._link_pos = None
        if append and os.path.exists(path) and os.path.getsize(path) > 16:
            self.f = open(path, "r+b")
            r = _Reader(path)
            if r.e != "<" or r.big != bigtiff:
                r.close()
                raise TiffError("append needs a little-endian file of the same TIFF flavour")
            off = r.first_ifd
            link = 8 if bigtiff else 4
            while off:
                r.f.seek(off)
                n = struct.unpack("<Q" if bigtiff else "<H", r.f.read(8 if bigtiff else 2))[0]
                link = off + (8 if bigtiff else 2) + n * (20 if bigtiff else 12)
                r.f.seek(link)
                off = struct.unpack("<Q" if bigtiff else "<I", r.f.read(8 if bigtiff else 4))[0]
            r.close()
            self._link_pos = link
            self.f.seek(0, os.SEEK_END)
        else:
            self.f = open(path, "wb")
            if bigtiff:
                self.f.write(struct.pack("<2sHHHQ", b"II", 43, 8, 0, 0))
                self._link_pos = 8
            else:
                self.f.write(struct.pack("<2sHI", b"II", 42, 0))
                self._link_pos = 4
        self._page = None

    def begin_page(self, height: int, width: int, dtype=np.uint8, description: Optional[str] = None) -> None:
        if self._page is not None:
            raise TiffError("previous page not finished")
        pos = self.f.tell()
        if pos % 16:
            self.f.write(b"\0" * (16 - pos % 16))
        self._page = dict(h=height, w=width, dt=np.dtype(dtype), data_off=self.f.tell(), rows=0, desc=description)

    def write_rows(self, rows: np.ndarray) -> None:
        p = self._page
        a = np.ascontiguousarray(rows, dtype=p["dt"])
        if a.ndim != 2 or a.shape[1] != p["w"] or p["rows"] + a.shape[0] > p["h"]:
            raise TiffError("rows do not fit the page")
        self.f.write(a.tobytes() if a.dtype.byteorder != ">" else a.byteswap().tobytes())
        p["rows"] += a.shape[0]

    def end_page(self) -> None:
        p = self._page
        if p["rows"] != p["h"]:
            raise TiffError(f"page has {p['rows']} of {p['h']} rows")
        self._write_ifd(p)
        self._page = None

    def _write_ifd(self, p: Dict) -> None:
        nbytes = p["h"] * p["w"] * p["dt"].itemsize
        fmt = {"u": 1, "i": 2, "f": 3}[p["dt"].kind]
        tags: List[Tuple[int, int, int, object]] = [
            (T_WIDTH, 4, 1, p["w"]), (T_LENGTH, 4, 1, p["h"]), (T_BITS, 3, 1, p["dt"].itemsize * 8),
            (T_COMPRESSION, 3, 1, 1), (T_PHOTOMETRIC, 3, 1, 1),
            (T_STRIP_OFFSETS, 16 if self.big else 4, 1, p["data_off"]), (T_SPP, 3, 1, 1),
            (T_ROWS_PER_STRIP, 4, 1, p["h"]), (T_STRIP_COUNTS, 16 if self.big else 4, 1, nbytes),
            (T_SAMPLE_FORMAT, 3, 1, fmt)]
        if not self.big and (p["data_off"] + nbytes) >= 2 ** 32:
            raise TiffError("classic TIFF cannot exceed 4 GiB; use bigtiff")
        extra = b""
        self.f.seek(0, os.SEEK_END)
        pos = self.f.tell()
        if pos % 16:
            self.f.write(b"\0" * (16 - pos % 16))
        ifd_off = self.f.tell()
        if p["desc"]:
            d = p["desc"].encode("latin-1") + b"\0"
            tags.append((T_DESCRIPTION, 2, len(d), d))
        tags.sort(key=lambda t: t[0])
        n = len(tags)
        esz, hdr = (20, 8) if self.big else (12, 2)
        extra_off = ifd_off + hdr + n * esz + (8 if self.big else 4)
        body = struct.pack("<Q" if self.big else "<H", n)
        for tag, typ, cnt, val in tags:
            inline = 8 if self.big else 4
            if typ == 2:
                if len(val) <= inline:
                    payload = val.ljust(inline, b"\0")
                else:
                    payload = struct.pack("<Q" if self.big else "<I", extra_off + len(extra))
                    extra += val
            else:
                payload = struct.pack("<" + _TYPE_FMT[typ], val).ljust(inline, b"\0")
            body += struct.pack("<HH", tag, typ) + struct.pack("<Q" if self.big else "<I", cnt) + payload
        body += struct.pack("<Q" if self.big else "<I", 0)
        self.f.write(body + extra)
        end = self.f.tell()
        self.f.seek(self._link_pos)
        self.f.write(struct.pack("<Q" if self.big else "<I", ifd_off))
        self.f.seek(end)
        self._link_pos = ifd_off + hdr + n * esz

    # -- several pages streamed side by side: the K class maps of one image arrive band by band (rows r0..r1 of every
    # page at once), so the data areas of all pages are reserved first and rows are written in place as they come
    def begin_pages(self, n_pages: int, height: int, width: int, dtype=np.uint8) -> None:
        if self._page is not None or getattr(self, "_pages", None):
            raise TiffError("previous page not finished")
        self.f.seek(0, os.SEEK_END)
        pos = self.f.tell()
        if pos % 16:
            self.f.write(b"\0" * (16 - pos % 16))
        dt = np.dtype(dtype)
        nbytes = height * width * dt.itemsize
        stride = (nbytes + 15) & ~15
        base = self.f.tell()
        self._pages = [dict(h=height, w=width, dt=dt, data_off=base + i * stride, rows=0, desc=None, seen=np.zeros(height, bool))
                       for i in range(n_pages)]
        self.f.truncate(base + n_pages * stride)          # reserves (sparse) space; rows land by seek + write

    def write_page_rows(self, page: int, row0: int, rows: np.ndarray) -> None:
        p = self._pages[page]
        a = np.ascontiguousarray(rows, dtype=p["dt"])
        if a.ndim != 2 or a.shape[1] != p["w"] or row0 < 0 or row0 + a.shape[0] > p["h"]:
            raise TiffError("rows do not fit the page")
        self.f.seek(p["data_off"] + row0 * p["w"] * p["dt"].itemsize)
        self.f.write(a.tobytes())
        p["seen"][row0:row0 + a.shape[0]] = True

    def end_pages(self) -> None:
        for p in self._pages:
            if not p["seen"].all():
                raise TiffError(f"page is missing {int((~p['seen']).sum())} of {p['h']} rows")
        for p in self._pages:
            self._write_ifd(p)
        self._pages = None

    def write_page(self, array: np.ndarray, description: Optional[str] = None) -> None:
        a = np.asarray(array)
        self.begin_page(a.shape[0], a.shape[1], a.dtype, description)
        step = max(1, (64 << 20) // max(1, a.shape[1] * a.dtype.itemsize))
        for r in range(0, a.shape[0], step):
            self.write_rows(a[r:r + step])
        self.end_page()

    def close(self) -> None:
        if self._page is not None or getattr(self, "_pages", None):
            raise TiffError("page not finished")
        self.f.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.f.close()


def imsave(path: str, array: np.ndarray, append: bool = False, bigtiff: bool = True) -> None:
    """skimage.io.imsave(path, array, bigtiff=True, append=...) for a 2-D page or a [pages,H,W] stack."""
    a = np.asarray(array)
    with BigTiffWriter(path, append=append, bigtiff=bigtiff) as w:
        if a.ndim == 2:
            w.write_page(a)
        elif a.ndim == 3:
            for pg in a:
                w.write_page(pg)
        else:
            raise TiffError("expected [H,W] or [pages,H,W]")
