"""TensorFlow checkpoint-V2 ("tensor bundle") reader that needs no TensorFlow.

Replaces ``tf.train.Saver().restore(sess, modelPath + '/model.ckpt')`` of the
reference (UnMicst1-5.py:677-681, UnMicst.py:507-512) for the inference path:
the named fp32 variables are pulled straight out of ``<prefix>.index`` (a
LevelDB-style sorted string table) and ``<prefix>.data-00000-of-00001`` (raw
little-endian tensors).  Optimiser slots saved next to the weights
(``*/Adam``, ``*/Adam_1``, ``*/Momentum``, ``optim/beta*_power``, the
``Variable`` global step) are recognised and skipped.

On-disk layout:
  footer (last 48 bytes)  varint64 metaindex{off,size}, index{off,size}, pad,
                          magic 0xdb4775248b80fb57 (LE)
  block                   prefix-compressed entries + uint32 restarts[n] + n,
                          followed by 1 byte compression type and 4 byte CRC
  index block values      (varint off, varint size) handles of data blocks
  data block values       BundleEntryProto {1:dtype 2:shape 3:shard 4:offset
                          5:size 6:crc32c}; key "" holds the BundleHeaderProto
"""
from __future__ import annotations

import os
import struct
from dataclasses import dataclass
from typing import Dict, Iterator, List, Tuple

import numpy as np

_TABLE_MAGIC = 0xDB4775248B80FB57
_FOOTER_LEN = 48

# TensorFlow DataType enum values that occur in the shipped checkpoints.
_DTYPES = {1: np.dtype("<f4"), 2: np.dtype("<f8"), 3: np.dtype("<i4"), 9: np.dtype("<i8")}

_SLOT_SUFFIXES = ("/Adam", "/Adam_1", "/Momentum")
_SLOT_NAMES = ("Variable", "optim/beta1_power", "optim/beta2_power", "beta1_power", "beta2_power")


class BundleError(ValueError):
    """Raised when the .index/.data pair is not a readable tensor bundle."""


@dataclass(frozen=True)
class BundleEntry:
    name: str
    dtype: np.dtype
    shape: Tuple[int, ...]
    shard: int
    offset: int
    size: int
    crc32c: int


def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    out = 0
    shift = 0
    while True:
        if pos >= len(buf):
            raise BundleError("truncated varint")
        b = buf[pos]
        pos += 1
        out |= (b & 0x7F) << shift
        if b < 0x80:
            return out, pos
        shift += 7
        if shift > 63:
            raise BundleError("varint too long")


def _proto_fields(buf: bytes) -> Iterator[Tuple[int, int, object]]:
    """Yield (field number, wire type, value) of one protobuf message."""
    pos = 0
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        field, wire = tag >> 3, tag & 7
        if wire == 0:
            val, pos = _varint(buf, pos)
        elif wire == 1:
            val = buf[pos:pos + 8]
            pos += 8
        elif wire == 2:
            n, pos = _varint(buf, pos)
            val = buf[pos:pos + n]
            pos += n
        elif wire == 5:
            val = buf[pos:pos + 4]
            pos += 4
        else:
            raise BundleError(f"unsupported protobuf wire type {wire}")
        yield field, wire, val


def _parse_shape(buf: bytes) -> Tuple[int, ...]:
    dims: List[int] = []
    for field, wire, val in _proto_fields(buf):
        if field == 2 and wire == 2:  # repeated Dim
            size = 0
            for f2, w2, v2 in _proto_fields(val):
                if f2 == 1 and w2 == 0:
                    size = v2 if v2 < (1 << 63) else v2 - (1 << 64)
            dims.append(int(size))
    return tuple(dims)


def _parse_entry(name: str, buf: bytes) -> BundleEntry:
    dtype_code, shape, shard, offset, size, crc = 0, (), 0, 0, 0, 0
    for field, wire, val in _proto_fields(buf):
        if field == 1 and wire == 0:
            dtype_code = val
        elif field == 2 and wire == 2:
            shape = _parse_shape(val)
        elif field == 3 and wire == 0:
            shard = val
        elif field == 4 and wire == 0:
            offset = val
        elif field == 5 and wire == 0:
            size = val
        elif field == 6 and wire == 5:
            crc = struct.unpack("<I", val)[0]
    if dtype_code not in _DTYPES:
        raise BundleError(f"tensor {name!r}: unsupported TF dtype code {dtype_code}")
    return BundleEntry(name, _DTYPES[dtype_code], shape, int(shard), int(offset), int(size), crc)


def _block_entries(table: bytes, off: int, size: int) -> Iterator[Tuple[bytes, bytes]]:
    if off + size + 5 > len(table):
        raise BundleError("block handle past end of file")
    if table[off + size] != 0:
        raise BundleError("compressed table blocks are not supported")
    block = table[off:off + size]
    if len(block) < 4:
        raise BundleError("block too small")
    n_restarts = struct.unpack("<I", block[-4:])[0]
    end = len(block) - 4 - 4 * n_restarts
    if end < 0:
        raise BundleError("corrupt restart array")
    pos = 0
    key = b""
    while pos < end:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        yield key, block[pos:pos + vlen]
        pos += vlen


def read_index(index_path: str) -> Dict[str, BundleEntry]:
    """Parse ``<prefix>.index`` and return {tensor name: BundleEntry} in key order."""
    with open(index_path, "rb") as f:
        table = f.read()
    if len(table) < _FOOTER_LEN:
        raise BundleError(f"{index_path}: too short for an SSTable")
    footer = table[-_FOOTER_LEN:]
    if struct.unpack("<Q", footer[-8:])[0] != _TABLE_MAGIC:
        raise BundleError(f"{index_path}: bad table magic")
    pos = 0
    _, pos = _varint(footer, pos)   # metaindex offset
    _, pos = _varint(footer, pos)   # metaindex size
    idx_off, pos = _varint(footer, pos)
    idx_size, pos = _varint(footer, pos)
    entries: Dict[str, BundleEntry] = {}
    for _, handle in _block_entries(table, idx_off, idx_size):
        boff, p = _varint(handle, 0)
        bsize, _ = _varint(handle, p)
        for key, val in _block_entries(table, boff, bsize):
            if key == b"":
                continue  # BundleHeaderProto
            name = key.decode("utf-8")
            entries[name] = _parse_entry(name, val)
    return entries


def is_optimizer_slot(name: str) -> bool:
    return name in _SLOT_NAMES or name.endswith(_SLOT_SUFFIXES)


def data_path(prefix: str, shard: int = 0, num_shards: int = 1) -> str:
    return f"{prefix}.data-{shard:05d}-of-{num_shards:05d}"


def load_bundle(prefix: str, skip_slots: bool = True) -> Dict[str, np.ndarray]:
    """Load every (non-optimiser) tensor of the bundle ``prefix`` as a numpy array."""
    entries = read_index(prefix + ".index")
    dpath = data_path(prefix)
    if not os.path.exists(dpath):
        raise FileNotFoundError(f"{dpath}: tensor data shard missing (only .index/.meta shipped?)")
    fsize = os.path.getsize(dpath)
    out: Dict[str, np.ndarray] = {}
    with open(dpath, "rb") as f:
        for name, e in entries.items():
            if skip_slots and is_optimizer_slot(name):
                continue
            if e.shard != 0:
                raise BundleError(f"{name}: multi-shard bundles are not supported")
            want = int(np.prod(e.shape, dtype=np.int64)) * e.dtype.itemsize
            if want != e.size or e.offset + e.size > fsize:
                raise BundleError(f"{name}: entry size/offset inconsistent with shape or data file")
            f.seek(e.offset)
            raw = f.read(e.size)
            out[name] = np.frombuffer(raw, dtype=e.dtype).reshape(e.shape).copy()
    return out
