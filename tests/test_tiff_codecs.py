"""Channel-page reader on compressed layouts: strips and tiles, Deflate and LZW, horizontal
predictor, both byte orders — against independent encoders (Pillow/libtiff, zlib, a reference LZW encoder
written here) and without touching a GPU."""
import struct
import zlib

import numpy as np
import pytest

from unmicst_b200 import _lib, tiffio


def _image(h, w, dtype, seed=0):
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w]
    smooth = (np.sin(yy / 7.0) + np.cos(xx / 11.0) + 2) * (0.2 * np.iinfo(dtype).max)
    noise = rng.integers(0, max(2, np.iinfo(dtype).max // 50), size=(h, w))
    return (smooth + noise).astype(dtype)


def _native(path, page=0):
    r = tiffio._Reader(str(path))
    try:
        return r.read_page(page)       # None would mean "fell back to PIL"
    finally:
        r.close()


@pytest.mark.parametrize("dtype,mode", [(np.uint8, "L"), (np.uint16, "I;16")])
@pytest.mark.parametrize("compression", ["tiff_lzw", "tiff_adobe_deflate", "raw"])
def test_pillow_written_files_decode_natively(tmp_path, dtype, mode, compression):
    from PIL import Image
    a = _image(211, 333, dtype, seed=3)
    p = tmp_path / f"{mode.replace(';', '')}_{compression}.tif"
    im = Image.fromarray(a)
    assert im.mode == mode
    im.save(p, compression=compression)
    got = _native(p)
    assert got is not None and got.dtype == dtype
    assert np.array_equal(got, a)
    assert np.array_equal(tiffio.read_page(str(p)), a)


def test_pillow_lzw_with_horizontal_predictor(tmp_path):
    from PIL import Image
    a = _image(97, 260, np.uint8, seed=4)
    p = tmp_path / "pred.tif"
    Image.fromarray(a).save(p, compression="tiff_lzw", tiffinfo={317: 2})
    r = tiffio._Reader(str(p))
    tags = r.page_tags(0)
    r.close()
    if tags.get(tiffio.T_PREDICTOR, (1,))[0] != 2:
        pytest.skip("this Pillow build does not write the predictor tag")
    assert np.array_equal(_native(p), a)


def _bigtiff_tiled(path, a, tile, comp, predictor, endian="<"):
    """Hand-written BigTIFF with one tiled page (the layout ASHLAR / bioformats2raw produce)."""
    h, w = a.shape
    e = endian
    dt = a.dtype.newbyteorder(e)
    tiles = []
    for r0 in range(0, h, tile):
        for c0 in range(0, w, tile):
            t = np.zeros((tile, tile), a.dtype)
            blk = a[r0:r0 + tile, c0:c0 + tile]
            t[:blk.shape[0], :blk.shape[1]] = blk
            if predictor == 2:
                t = np.concatenate([t[:, :1], np.diff(t, axis=1)], axis=1).astype(a.dtype)
            raw = t.astype(dt).tobytes()
            tiles.append(zlib.compress(raw) if comp == 8 else raw)
    body = b"".join(tiles)
    offs, pos = [], 16
    for t in tiles:
        offs.append(pos); pos += len(t)
    n = len(tiles)
    ifd_off = 16 + len(body)
    entries = [(256, 4, 1, w), (257, 4, 1, h), (258, 3, 1, a.dtype.itemsize * 8), (259, 3, 1, comp), (262, 3, 1, 1),
               (277, 3, 1, 1), (317, 3, 1, predictor), (322, 3, 1, tile), (323, 3, 1, tile), (339, 3, 1, 1)]
    arrays = struct.pack(e + "Q" * n, *offs) + struct.pack(e + "Q" * n, *[len(t) for t in tiles])
    arr_off = ifd_off + 8 + 20 * (len(entries) + 2) + 8
    ifd = struct.pack(e + "Q", len(entries) + 2)
    allent = entries + [(324, 16, n, None), (325, 16, n, None)]
    for tag, typ, cnt, val in sorted(allent):
        if val is None:
            o = arr_off if tag == 324 else arr_off + 8 * n
            v = struct.pack(e + "Q", o) if n > 1 else struct.pack(e + "Q", offs[0] if tag == 324 else len(tiles[0]))
        else:
            v = struct.pack(e + {3: "H", 4: "I"}[typ], val).ljust(8, b"\0")
        ifd += struct.pack(e + "HHQ", tag, typ, cnt) + v
    ifd += struct.pack(e + "Q", 0)
    with open(path, "wb") as f:
        f.write((b"II" if e == "<" else b"MM") + struct.pack(e + "HHHQ", 43, 8, 0, ifd_off))
        f.write(body); f.write(ifd); f.write(arrays)


@pytest.mark.parametrize("endian", ["<", ">"])
@pytest.mark.parametrize("comp,predictor", [(1, 1), (8, 1), (8, 2)])
def test_bigtiff_tiles_deflate_predictor_both_byte_orders(tmp_path, comp, predictor, endian):
    a = _image(300, 421, np.uint16, seed=5)
    p = tmp_path / "tiled.tif"
    _bigtiff_tiled(p, a, 128, comp, predictor, endian)
    got = _native(p)
    assert got is not None and got.shape == a.shape
    assert np.array_equal(got, a)


def _lzw_encode(data: bytes) -> bytes:
    """Reference TIFF-LZW encoder (MSB first, early change), independent of the decoder under test."""
    out, acc, nb = bytearray(), 0, 0

    def put(code, width):
        nonlocal acc, nb
        acc = (acc << width) | code; nb += width
        while nb >= 8:
            out.append((acc >> (nb - 8)) & 0xFF); nb -= 8

    table = {bytes([i]): i for i in range(256)}
    nxt, width = 258, 9
    put(256, width)
    w = b""
    for byte in data:
        wc = w + bytes([byte])
        if wc in table:
            w = wc
            continue
        put(table[w], width)
        table[wc] = nxt; nxt += 1
        if nxt == (1 << width) - 1 + 1 and width < 12:      # the decoder's table is one entry behind the encoder's
            width += 1
        if nxt == 4094 + 1:
            put(256, width)
            table = {bytes([i]): i for i in range(256)}
            nxt, width = 258, 9
        w = bytes([byte])
    if w:
        put(table[w], width)
    put(257, width)
    if nb:
        out.append((acc << (8 - nb)) & 0xFF)
    return bytes(out)


@pytest.mark.parametrize("n,kind", [(0, "zeros"), (1, "zeros"), (5000, "zeros"), (70000, "text"), (200000, "random")])
def test_lzw_decoder_against_reference_encoder(n, kind):
    rng = np.random.default_rng(n)
    if kind == "zeros":
        data = bytes(n)
    elif kind == "text":
        data = (b"the quick brown fox jumps over the lazy dog " * (n // 44 + 1))[:n]
    else:
        data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()       # incompressible: many table resets
    enc = np.frombuffer(_lzw_encode(data), dtype=np.uint8)
    out = np.empty(max(n, 1), dtype=np.uint8)
    got = _lib.lib().umx_tiff_lzw_decode(enc.ctypes.data, enc.size, out.ctypes.data, n)
    assert got == n
    assert out[:n].tobytes() == data
    # a too-small destination is filled, never overrun
    if n > 10:
        small = np.full(n // 2 + 8, 0xAB, dtype=np.uint8)
        got = _lib.lib().umx_tiff_lzw_decode(enc.ctypes.data, enc.size, small.ctypes.data, n // 2)
        assert got == n // 2 and small[:n // 2].tobytes() == data[:n // 2] and (small[n // 2:] == 0xAB).all()


def test_lzw_decoder_rejects_garbage():
    bad = np.array([0xFF] * 64, dtype=np.uint8)            # starts with code 511: not a valid first code
    out = np.empty(256, dtype=np.uint8)
    assert _lib.lib().umx_tiff_lzw_decode(bad.ctypes.data, bad.size, out.ctypes.data, out.size) < 0
