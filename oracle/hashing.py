"""Oracle (test infrastructure): string hash -> bucket.

Restates ``/root/reference/torchctr/utils.py:103-119``::

    hash_bucket(v, buckets, seed) == murmurhash3_32(str(v), seed, positive=True) % buckets

``murmurhash3_32`` is scikit-learn's Cython wrapper of Austin Appleby's public
domain MurmurHash3_x86_32 (``sklearn/utils/murmurhash.pyx`` +
``src/MurmurHash3.cpp``); it is not vendored in the reference, so the published
algorithm is restated here: 4-byte little-endian blocks mixed with
c1=0xcc9e2d51 / c2=0x1b873593, rotl 15 / 13, ``h = h*5 + 0xe6546b64``, the
1..3 byte tail, ``h ^= len`` and the 0x85ebca6b / 0xc2b2ae35 finaliser.

The value that reaches the hash in the reference is the *canonical string* of a
category (``transformer.py:367-401``): anything that parses as a number is cast
to Int32 and printed in decimal, so integer ids hash as their decimal ASCII.
``hash_bucket_ids`` is that case, vectorised with numpy.
"""
from __future__ import annotations

import numpy as np

_C1 = 0xCC9E2D51
_C2 = 0x1B873593
_M32 = 0xFFFFFFFF


def _rotl(x: int, r: int) -> int:
    return ((x << r) | (x >> (32 - r))) & _M32


def murmur3_32(data: bytes, seed: int = 0) -> int:
    """MurmurHash3_x86_32 of ``data``; unsigned result (sklearn ``positive=True``)."""
    if not 0 <= seed <= _M32:
        raise OverflowError("seed must fit uint32 (sklearn raises OverflowError too)")
    h = seed
    n = len(data)
    for off in range(0, n - (n % 4), 4):
        k = int.from_bytes(data[off:off + 4], "little")
        k = (k * _C1) & _M32
        k = _rotl(k, 15)
        k = (k * _C2) & _M32
        h ^= k
        h = _rotl(h, 13)
        h = (h * 5 + 0xE6546B64) & _M32
    rem = n % 4
    if rem:
        k = int.from_bytes(data[n - rem:], "little")
        k = (k * _C1) & _M32
        k = _rotl(k, 15)
        k = (k * _C2) & _M32
        h ^= k
    h ^= n
    h ^= h >> 16
    h = (h * 0x85EBCA6B) & _M32
    h ^= h >> 13
    h = (h * 0xC2B2AE35) & _M32
    h ^= h >> 16
    return h


def hash_bucket(v, buckets: int, seed: int = 0) -> int:
    """``utils.py:113,119``: murmur3 of ``str(v)`` (utf-8), unsigned, modulo ``buckets``."""
    return murmur3_32(str(v).encode("utf-8"), seed) % buckets


def _murmur3_words(words: np.ndarray, nbytes: np.ndarray, seed: int) -> np.ndarray:
    """Vectorised murmur3 over rows of little-endian u32 words (unused bytes zero)."""
    with np.errstate(over="ignore"):
        h = np.full(words.shape[0], seed, dtype=np.uint32)
        c1, c2 = np.uint32(_C1), np.uint32(_C2)
        nblocks = nbytes // 4
        for w in range(words.shape[1]):
            k = words[:, w] * c1
            k = (k << np.uint32(15)) | (k >> np.uint32(17))
            k = k * c2
            full = nblocks > w
            tail = (nblocks == w) & ((nbytes % 4) != 0)
            hk = h ^ k
            h_full = (hk << np.uint32(13)) | (hk >> np.uint32(19))
            h_full = h_full * np.uint32(5) + np.uint32(0xE6546B64)
            h = np.where(full, h_full, np.where(tail, hk, h))
        h = h ^ nbytes.astype(np.uint32)
        h ^= h >> np.uint32(16)
        h = h * np.uint32(0x85EBCA6B)
        h ^= h >> np.uint32(13)
        h = h * np.uint32(0xC2B2AE35)
        h ^= h >> np.uint32(16)
    return h


def decimal_ascii_words(ids: np.ndarray):
    """Decimal ASCII of signed 64-bit ints packed into 5 little-endian u32 words.

    Mirrors ``transformer.py:371-381``: a numeric category becomes the decimal
    string of its integer value before it is hashed.
    """
    ids = np.asarray(ids, dtype=np.int64).ravel()
    n = ids.shape[0]
    neg = ids < 0
    mag = np.where(neg, -(ids + 1), ids).astype(np.uint64) + neg.astype(np.uint64)
    digits = np.zeros((n, 20), dtype=np.uint8)
    ndig = np.zeros(n, dtype=np.int64)
    m = mag.copy()
    for p in range(20):
        digits[:, p] = (m % np.uint64(10)).astype(np.uint8)
        live = (m > 0) | (p == 0)
        ndig = np.where(live, p + 1, ndig)
        m = m // np.uint64(10)
    length = ndig + neg
    buf = np.zeros((n, 20), dtype=np.uint8)
    for p in range(20):
        # character p of the string: '-' first when negative, then most significant digit
        dpos = ndig - 1 - (p - neg.astype(np.int64))
        is_digit = (dpos >= 0) & (p >= neg.astype(np.int64))
        d = digits[np.arange(n), np.clip(dpos, 0, 19)]
        ch = np.where(is_digit, d + ord("0"), 0).astype(np.uint8)
        if p == 0:
            ch = np.where(neg, ord("-"), ch).astype(np.uint8)
        buf[:, p] = ch
    words = buf.view("<u4").reshape(n, 5).astype(np.uint32)
    return words, length.astype(np.int64)


def murmur3_decimal(ids: np.ndarray, seed: int = 0) -> np.ndarray:
    """murmur3_32(str(int(id))) for an array of signed 64-bit ids -> uint32 array."""
    words, length = decimal_ascii_words(ids)
    return _murmur3_words(words, length, seed)


def hash_bucket_ids(ids: np.ndarray, buckets: int, seed: int = 0) -> np.ndarray:
    """Vectorised ``hash_bucket(int(id), buckets, seed)``; int32 like ``transformer.py:490``."""
    shape = np.asarray(ids).shape
    h = murmur3_decimal(ids, seed)
    return (h % np.uint32(buckets)).astype(np.int32).reshape(shape)
