"""Minimal Zeiss CZI (ZISRAW) channel-plane reader.

Replaces, for the hot path's needs only, ``czifile.CziFile(path).asarray()[0, 0, channel, 0, 0, :, :, 0]``
(UnMicst1-5.py:797-800; the package is not installable here): one 2-D channel plane, assembled from the
file's sub-blocks.  Supported: uncompressed Gray8 / Gray16 / Gray32Float sub-blocks, single- or multi-tile
(mosaic) scenes, pyramid levels skipped.  Compressed sub-blocks (JPEG, JPEG-XR, zstd) and colour pixel types raise.

Layout restated from the published ZISRAW segment structure (the same one czifile implements):
  segment   = 16-byte id, int64 allocated size, int64 used size, then the payload
  ZISRAWFILE payload: major, minor, 2 reserved int32, two 16-byte GUIDs, file part, int64 directory position, ...
  ZISRAWDIRECTORY payload: int32 entry count, 124 reserved bytes, then DirectoryEntryDV records
  DirectoryEntryDV: 'DV', int32 pixel type, int64 file position of the sub-block segment, int32 file part,
                    int32 compression, uint8 pyramid type, 5 reserved bytes, int32 dimension count, then per
                    dimension: 4-char name, int32 start, int32 size, float32 start coordinate, int32 stored size
  ZISRAWSUBBLOCK payload: int32 metadata size, int32 attachment size, int64 data size, a DirectoryEntryDV copy padded
                    to at least 240 bytes, the metadata, then the pixel data
No CZI file exists in the reference tree: the reader is validated against files written by tests/ own encoder of
this layout — parity with czifile itself is unpinned.
"""
from __future__ import annotations

import struct
from typing import Dict, List, Tuple

import numpy as np

_PIXEL = {0: np.dtype("<u1"), 1: np.dtype("<u2"), 2: np.dtype("<f4")}          # Gray8, Gray16, Gray32Float
_SEG = struct.Struct("<16sqq")
_ENTRY = struct.Struct("<2siqiiB5si")
_DIM = struct.Struct("<4siifi")


class CziError(ValueError):
    pass


def _segment(f, pos: int) -> Tuple[bytes, int]:
    f.seek(pos)
    sid, _alloc, used = _SEG.unpack(f.read(32))
    return sid.rstrip(b"\0"), used


def _entry(buf: bytes, off: int):
    schema, ptype, fpos, _part, comp, pyramid, _r, ndim = _ENTRY.unpack_from(buf, off)
    if schema != b"DV":
        raise CziError(f"unsupported directory entry schema {schema!r}")
    off += _ENTRY.size
    dims: Dict[str, Tuple[int, int, int]] = {}
    for _ in range(ndim):
        name, start, size, _coord, stored = _DIM.unpack_from(buf, off)
        dims[name.rstrip(b"\0").decode("ascii")] = (start, size, stored or size)
        off += _DIM.size
    return dict(pixel=ptype, pos=fpos, comp=comp, pyramid=pyramid, dims=dims), off


def directory(path: str) -> List[dict]:
    with open(path, "rb") as f:
        sid, _ = _segment(f, 0)
        if sid != b"ZISRAWFILE":
            raise CziError(f"{path}: not a CZI file")
        head = f.read(80)
        dir_pos = struct.unpack_from("<q", head, 52)[0]
        sid, used = _segment(f, dir_pos)
        if sid != b"ZISRAWDIRECTORY":
            raise CziError(f"{path}: sub-block directory not found")
        buf = f.read(used)
    count = struct.unpack_from("<i", buf, 0)[0]
    off, out = 128, []
    for _ in range(count):
        e, off = _entry(buf, off)
        out.append(e)
    return out


def read_channel(path: str, channel: int) -> np.ndarray:
    """The 2-D plane of ``channel`` at the first index of every other non-spatial dimension."""
    entries = [e for e in directory(path) if all(d[1] == d[2] for d in e["dims"].values())]      # full resolution only
    if not entries:
        raise CziError(f"{path}: no full-resolution sub-blocks")
    other = sorted({k for e in entries for k in e["dims"]} - {"X", "Y", "C", "M"})
    first = {k: min(e["dims"].get(k, (0, 1, 1))[0] for e in entries) for k in other}
    chans = sorted({e["dims"].get("C", (0, 1, 1))[0] for e in entries})
    if channel < 0 or channel >= len(chans):
        raise IndexError(f"{path}: channel {channel} of {len(chans)}")
    pick = [e for e in entries if e["dims"].get("C", (0, 1, 1))[0] == chans[channel]
            and all(e["dims"].get(k, (first[k], 1, 1))[0] == first[k] for k in other)]
    x0 = min(e["dims"]["X"][0] for e in pick)
    y0 = min(e["dims"]["Y"][0] for e in pick)
    x1 = max(e["dims"]["X"][0] + e["dims"]["X"][1] for e in pick)
    y1 = max(e["dims"]["Y"][0] + e["dims"]["Y"][1] for e in pick)
    if pick[0]["pixel"] not in _PIXEL:
        raise NotImplementedError(f"{path}: CZI pixel type {pick[0]['pixel']} (colour / complex) is not supported")
    dt = _PIXEL[pick[0]["pixel"]]
    out = np.zeros((y1 - y0, x1 - x0), dtype=dt.newbyteorder("="))
    with open(path, "rb") as f:
        for e in pick:
            if e["comp"] != 0:
                raise NotImplementedError(f"{path}: compressed CZI sub-blocks (compression {e['comp']}) are not supported; "
                                          f"export uncompressed or convert to (OME-)TIFF")
            sid, _ = _segment(f, e["pos"])
            if sid != b"ZISRAWSUBBLOCK":
                raise CziError(f"{path}: sub-block segment expected at {e['pos']}")
            meta_size, _att, data_size = struct.unpack("<iiq", f.read(16))
            sub, end = _entry(f.read(_ENTRY.size + 20 * 16), 0)
            f.seek(e["pos"] + 32 + 16 + max(240, end) + meta_size)
            h, w = sub["dims"]["Y"][1], sub["dims"]["X"][1]
            if data_size < h * w * dt.itemsize:
                raise CziError(f"{path}: sub-block holds {data_size} bytes, {h}x{w} {dt} needs {h * w * dt.itemsize}")
            tile = np.frombuffer(f.read(h * w * dt.itemsize), dtype=dt).reshape(h, w)
            ys, xs = sub["dims"]["Y"][0] - y0, sub["dims"]["X"][0] - x0
            out[ys:ys + h, xs:xs + w] = tile
    return out
