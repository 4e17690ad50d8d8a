"""The CZI channel reader (czifile.CziFile(...).asarray()[0, 0, c, 0, 0, :, :, 0], UnMicst1-5.py:797-800) against files
written by a test-only encoder of the published ZISRAW segment layout (no CZI file ships with the reference and czifile
is not installable here: parity with czifile itself is unpinned)."""
import struct

import numpy as np
import pytest

from unmicst_b200 import cli, czi


def _seg(sid: bytes, payload: bytes) -> bytes:
    pad = (-len(payload)) % 32
    return struct.pack("<16sqq", sid, len(payload) + pad, len(payload)) + payload + b"\0" * pad


def _entry(pixel, pos, comp, dims, pyramid=0):
    b = struct.pack("<2siqiiB5si", b"DV", pixel, pos, 0, comp, pyramid, b"", len(dims))
    for name, start, size, stored in dims:
        b += struct.pack("<4siifi", name.encode(), start, size, float(start), stored)
    return b


def write_czi(path, planes, tile=None, comp=0, with_pyramid=False):
    """planes: [C][H][W]; optional mosaic of tile x tile sub-blocks; optional 2x-downsampled pyramid sub-blocks."""
    C, H, W = planes.shape
    ptype = {np.dtype("uint8"): 0, np.dtype("uint16"): 1, np.dtype("float32"): 2}[planes.dtype]
    tile = tile or max(H, W)
    blobs, entries = [], []
    pos = 32 + 512                                            # file header segment: 32-byte header + 512-byte payload
    m = 0
    for c in range(C):
        for y in range(0, H, tile):
            for x in range(0, W, tile):
                t = np.ascontiguousarray(planes[c, y:y + tile, x:x + tile])
                dims = [("X", x + 1000, t.shape[1], t.shape[1]), ("Y", y - 50, t.shape[0], t.shape[0]), ("C", c, 1, 1), ("Z", 0, 1, 1),
                        ("T", 0, 1, 1), ("M", m, 1, 1), ("B", 0, 1, 1)]
                e = _entry(ptype, pos, comp, dims)
                meta = b"<METADATA/>"
                payload = struct.pack("<iiq", len(meta), 0, t.nbytes) + e + b"\0" * max(0, 240 - len(e)) + meta + t.tobytes()
                seg = _seg(b"ZISRAWSUBBLOCK", payload)
                blobs.append(seg); entries.append(e); pos += len(seg); m += 1
        if with_pyramid:                                      # a half-size copy: stored size != size, must be ignored
            t = np.ascontiguousarray(planes[c, ::2, ::2])
            dims = [("X", 1000, W, t.shape[1]), ("Y", -50, H, t.shape[0]), ("C", c, 1, 1), ("Z", 0, 1, 1), ("T", 0, 1, 1), ("B", 0, 1, 1)]
            e = _entry(ptype, pos, comp, dims, pyramid=1)
            payload = struct.pack("<iiq", 0, 0, t.nbytes) + e + b"\0" * max(0, 240 - len(e)) + t.tobytes()
            seg = _seg(b"ZISRAWSUBBLOCK", payload)
            blobs.append(seg); entries.append(e); pos += len(seg)
    directory = _seg(b"ZISRAWDIRECTORY", struct.pack("<i", len(entries)) + b"\0" * 124 + b"".join(entries))
    header = struct.pack("<iiii16s16siqqiq", 1, 0, 0, 0, b"g" * 16, b"g" * 16, 0, pos, 0, 0, 0).ljust(512, b"\0")
    with open(path, "wb") as f:
        f.write(struct.pack("<16sqq", b"ZISRAWFILE", 512, 512) + header)
        for b in blobs:
            f.write(b)
        f.write(directory)


@pytest.mark.parametrize("dtype,tile,pyr", [(np.uint16, None, False), (np.uint16, 70, True), (np.uint8, 64, False), (np.float32, None, True)])
def test_czi_channel_planes_round_trip(tmp_path, dtype, tile, pyr):
    rng = np.random.default_rng(0)
    planes = (rng.random((3, 150, 201)) * (255 if dtype == np.uint8 else 60000)).astype(dtype)
    p = str(tmp_path / "img.czi")
    write_czi(p, planes, tile=tile, with_pyramid=pyr)
    assert len(czi.directory(p)) >= 3
    for c in range(3):
        got = czi.read_channel(p, c)
        assert got.dtype == dtype and np.array_equal(got, planes[c])
    assert np.array_equal(cli.read_channel(p, "czi", 1), planes[1])            # the CLI's dispatch on the extension
    with pytest.raises(IndexError):
        czi.read_channel(p, 3)


def test_czi_unsupported_content_is_reported(tmp_path):
    planes = np.zeros((1, 16, 16), np.uint16)
    p = str(tmp_path / "jxr.czi")
    write_czi(p, planes, comp=4)                              # JPEG-XR
    with pytest.raises(NotImplementedError):
        czi.read_channel(p, 0)
    q = str(tmp_path / "not.czi")
    open(q, "wb").write(b"II*\0" + b"\0" * 100)
    with pytest.raises(czi.CziError):
        czi.read_channel(q, 0)
    with pytest.raises(NotImplementedError):
        cli.read_channel("x.nd2", "nd2", 0)                   # nd2reader is absent and the format unpublished
