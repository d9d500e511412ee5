"""Oracle (test infrastructure): the canonical string of a raw category value.

Restates ``/root/reference/torchctr/transformer.py:367-401`` (``FeatureTransformer._process_category``) one value at a
time.  PARITY UNPINNED for this function: it is written in polars expressions and polars is not installed here, so the
restatement follows the source text and polars' documented cast rules (Utf8 -> Float64 ``strict=False`` yields null for
anything that is not a number literal; Float64 -> Int32 ``strict=False`` truncates toward zero and yields null outside the
Int32 range); what IS pinned is the hash applied afterwards (``hash_bucket`` of the reference, tests/golden/hash_golden.json).
"""
from __future__ import annotations

import math


def canonical(v, case_sensitive=True, outliers=None, fillna=None, oov="other"):
    # :371  int_s = s.cast(Float64, strict=False).fill_nan(None).cast(Int32, strict=False)
    f = None
    if isinstance(v, bool) or v is None:
        f = None
    elif isinstance(v, (int, float)):
        f = float(v)
    elif isinstance(v, str):
        try:
            f = float(v) if v == v.strip() else None
        except ValueError:
            f = None
    if f is not None and (math.isnan(f) or math.isinf(f)):
        f = None
        if not isinstance(v, str):
            v = None                      # a NaN number is a missing value
    int_s = None
    if f is not None:
        t = math.trunc(f)
        if -2 ** 31 <= t <= 2 ** 31 - 1:
            int_s = t
    # :377-381  when(is_int).then(int_s.cast(Utf8)).otherwise(s)
    s = str(int_s) if int_s is not None else (None if v is None else str(v))
    if s is not None and not case_sensitive:          # :383-384
        s = s.lower()
    if outliers:                                       # :386-394
        table = {o: oov for o in outliers} if isinstance(outliers, list) else dict(outliers)
        if not case_sensitive:
            table = {k.lower(): val for k, val in table.items()}
        if s in table:
            s = table[s]
    if s is None:                                      # :396-399
        s = fillna if fillna else "__null__"
    return s
