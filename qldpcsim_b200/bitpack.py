"""Bit-packed shot I/O: row r of a 0/1 matrix becomes ceil(cols/32) little-endian uint32 words, bit j of
word w = column 32*w + j.  This is the syndrome / error-estimate layout of the C-ABI (include/qldpc_b200.h)."""
from __future__ import annotations

import numpy as np


def words(nbits: int) -> int:
    return (int(nbits) + 31) // 32


def pack_rows(bits: np.ndarray) -> np.ndarray:
    bits = np.ascontiguousarray(np.asarray(bits).astype(np.uint8) & 1)
    if bits.ndim == 1:
        bits = bits[None, :]
    rows, cols = bits.shape
    W = words(cols)
    padded = np.zeros((rows, W * 32), dtype=np.uint8)
    padded[:, :cols] = bits
    by = np.packbits(padded, axis=1, bitorder="little")
    return np.ascontiguousarray(by).view(np.uint32).reshape(rows, W)


def unpack_rows(packed: np.ndarray, cols: int) -> np.ndarray:
    packed = np.ascontiguousarray(np.asarray(packed, dtype=np.uint32))
    if packed.ndim == 1:
        packed = packed[None, :]
    by = packed.view(np.uint8)
    return np.unpackbits(by, axis=1, bitorder="little")[:, :cols]
