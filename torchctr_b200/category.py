"""Raw category values -> what the reference hashes / looks up (host side of the category transform).

``FeatureTransformer._process_category`` (``torchctr/transformer.py:367-401``) turns a column into *canonical strings* before
``hash_bucket`` (``:487-490``) or the vocabulary (``:492-498``) sees it:

  1. whatever parses as a number and fits Int32 becomes the decimal string of that integer (``"12.0"`` -> ``"12"``, ``7.9`` ->
     ``"7"``: the Float64 -> Int32 cast truncates), NaN counts as missing (``:371``);
  2. everything else keeps its own characters; ``case_sensitive=False`` lower-cases (``:383-384``);
  3. ``outliers`` (a list -> all mapped to ``oov``, or a dict) are replaced (``:386-394``);
  4. missing values become ``fillna`` or ``'__null__'`` (``:396-399``).

``encode_categories`` does 1-4 on the host (strings are host data in the reference too) and splits the result into the two
forms the kernels take: int64 ids for the numeric ones (hashed on the device from their decimal ASCII, ``ctr_hash_bucket_i64``
/ ``index_kind='hash'`` inside the lookup) and packed utf-8 bytes for the rest (``ctr_hash_bucket_bytes``).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import ops

_INT32_MIN, _INT32_MAX = -2 ** 31, 2 ** 31 - 1


def _as_int32(v):
    """The Int32 the reference's cast chain yields for ``v``, or None (transformer.py:371)."""
    if v is None or isinstance(v, bool):
        return None
    if isinstance(v, (int, np.integer)):
        f = float(v)
    elif isinstance(v, (float, np.floating)):
        f = float(v)
    elif isinstance(v, str):
        try:
            f = float(v.strip()) if v.strip() == v else float("nan")      # polars' strict=False cast rejects padded strings
        except ValueError:
            return None
    else:
        return None
    if math.isnan(f) or math.isinf(f):
        return None
    i = int(f)                                                             # truncation toward zero
    return i if _INT32_MIN <= i <= _INT32_MAX else None


def canonical_categories(values, case_sensitive: bool = True, outliers=None, fillna=None, oov: str = "other"):
    """-> list[str]: the canonical string of every value (steps 1-4 above)."""
    if outliers is not None and len(outliers) > 0:
        if isinstance(outliers, list):
            outliers = {v: oov for v in outliers}
        if not isinstance(outliers, dict):
            raise ValueError("Outliers must be a list or a dictionary")
        if not case_sensitive:
            outliers = {k.lower(): v for k, v in outliers.items()}
    else:
        outliers = None
    out = []
    for v in values:
        if isinstance(v, (float, np.floating)) and math.isnan(float(v)):
            v = None
        i = _as_int32(v)
        s = str(i) if i is not None else (None if v is None else str(v))
        if s is not None and not case_sensitive:
            s = s.lower()
        if s is not None and outliers is not None and s in outliers:
            s = outliers[s]
        if s is None:
            s = fillna if fillna else "__null__"
        out.append(s)
    return out


def hash_bucket_strings(strings, buckets: int, seed: int = 0, device="cuda") -> torch.Tensor:
    """``[hash_bucket(s, buckets, seed) for s in strings]`` (``torchctr/utils.py:103-119``) on the device -> int32 [n]."""
    enc = [s.encode("utf-8") for s in strings]
    offsets = np.zeros(len(enc) + 1, dtype=np.int64)
    np.cumsum([len(b) for b in enc], out=offsets[1:])
    data = np.frombuffer(b"".join(enc), dtype=np.uint8) if offsets[-1] else np.zeros(1, dtype=np.uint8)
    return ops.hash_bucket_bytes(torch.from_numpy(data.copy()).to(device), torch.from_numpy(offsets).to(device), buckets, seed)


def encode_categories(values, buckets: int, seed: int = 0, device="cuda", **canon) -> torch.Tensor:
    """Raw column -> int32 hash buckets, as ``FeatureTransformer.process_category`` does for a feature with ``hash_buckets``
    (``transformer.py:487-490``): canonicalise on the host, hash on the device."""
    return hash_bucket_strings(canonical_categories(values, **canon), buckets, seed, device)
