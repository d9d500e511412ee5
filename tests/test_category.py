"""Row a8: raw category value -> canonical string -> hash bucket, incl. string categories.

The hash is pinned to the REAL reference: tests/golden/hash_golden.json holds ``torchctr.utils.hash_bucket`` /
sklearn ``murmurhash3_32`` outputs for strings ('', 'a', '__null__', 'B001NPEBGU', non-ASCII ...) and integers.  The
canonicalisation (``transformer.py:367-401``) is polars code that cannot run here: product and oracle restatements are
compared with each other and the documented cases of SURVEY.md 3.3 ("12.0" -> "12", null -> '__null__', lowercase,
outliers, numerics outside Int32 keep their original form) are asserted explicitly -- parity unpinned for that step.
"""
import json
import os

import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hash_golden.json")
RAW = [12.0, "12.0", "12", "abc", None, "ABC", 2 ** 40, -5, "007", " 7", 1.5, float("nan"), "1e3", "-2147483649", "2147483647",
       "B001NPEBGU", "", "ünïcode", 0, -0.0, "nan", "__null__", "Other"]


def test_canonical_categories_follow_the_reference_rules():
    from oracle.category import canonical
    from torchctr_b200.category import canonical_categories
    for kw in ({}, {"case_sensitive": False}, {"outliers": ["abc", "12"], "oov": "other"}, {"fillna": "missing"},
               {"case_sensitive": False, "outliers": {"ABC": "x", "b001npebgu": "y"}}):
        got = canonical_categories(RAW, **kw)
        ref = [canonical(v, kw.get("case_sensitive", True), kw.get("outliers"), kw.get("fillna"), kw.get("oov", "other")) for v in RAW]
        assert got == ref, kw
    c = canonical_categories(RAW)
    by = dict(zip(map(repr, RAW), c))
    assert by["12.0"] == by["'12.0'"] == by["'12'"] == "12"            # transformer.py:371-381: numbers print as their Int32
    assert by["1.5"] == "1" and by["'1e3'"] == "1000" and by["'007'"] == "7"
    assert by["None"] == by["nan"] == "__null__"                        # :371 fill_nan(None), :399
    assert by[repr(2 ** 40)] == str(2 ** 40) and by["'-2147483649'"] == "-2147483649"     # outside Int32: original form
    assert by["'ABC'"] == "ABC" and canonical_categories(["ABC"], case_sensitive=False) == ["abc"]
    assert canonical_categories(["a", "b"], outliers=["a"], oov="other") == ["other", "b"]
    with pytest.raises(ValueError):
        canonical_categories(["a"], outliers="a")


def test_oracle_string_hash_matches_reference_golden():
    from oracle.hashing import hash_bucket
    rows = [r for r in json.load(open(GOLDEN)) if not r["is_int"]]
    assert len(rows) >= 100
    for r in rows:
        assert hash_bucket(r["v"], r["buckets"], r["seed"]) == r["bucket"]


@pytest.mark.gpu
def test_device_string_hash_matches_reference_golden():
    from torchctr_b200.category import encode_categories, hash_bucket_strings
    rows = json.load(open(GOLDEN))
    for seed in sorted({r["seed"] for r in rows}):
        for buckets in sorted({r["buckets"] for r in rows}):
            sel = [r for r in rows if r["seed"] == seed and r["buckets"] == buckets]
            strings = [str(r["v"]) for r in sel]                      # hash_bucket hashes str(v) (utils.py:113)
            got = hash_bucket_strings(strings, buckets, seed).cpu().tolist()
            assert got == [r["bucket"] for r in sel], (seed, buckets)
    # the whole transform: canonicalise on the host, hash on the device == oracle canonical + reference-pinned hash
    from oracle.category import canonical
    from oracle.hashing import hash_bucket
    got = encode_categories(RAW, 1000003, seed=7, case_sensitive=False).cpu().tolist()
    assert got == [hash_bucket(canonical(v, False), 1000003, 7) for v in RAW]
    # long strings / many of them: block loop and tail of the byte-wise murmur
    gen = torch.Generator().manual_seed(0)
    many = ["k" * int(n) + str(i) for i, n in enumerate(torch.randint(0, 70, (5000,), generator=gen))]
    assert hash_bucket_strings(many, 2 ** 31 - 1, 3).cpu().tolist() == [hash_bucket(s, 2 ** 31 - 1, 3) for s in many]
    # numeric categories agree between the two device entry points (decimal ASCII on the fly vs packed bytes)
    from torchctr_b200 import ops
    ids = torch.randint(-2 ** 31, 2 ** 31, (4096,), generator=gen)
    a = ops.hash_bucket(ids.cuda(), 1000003, 5).cpu().tolist()
    b = hash_bucket_strings([str(int(i)) for i in ids], 1000003, 5).cpu().tolist()
    assert a == b
