"""Oracle (test infrastructure): vocabulary build / grow / lookup.

Restates ``/root/reference/torchctr/transformer.py:451-498`` for integer keys
(the canonical string of a numeric category is its decimal form, so int <-> str
is a bijection and the dict below can be keyed by int):

* fit (``:451-482``): count keys of the batch, keep those with
  ``count >= min_freq`` (``:456-460``); ``idx`` starts at the largest index in
  the existing vocab (0 for an empty one, ``:462-466``); every admitted key that
  is not yet known takes ``idx += 1`` (``:469-472``), known keys add the batch
  count to ``cnt`` (``:473-474``); the OOV entry owns index 0 (``:476-477``);
  ``num_embeddings = idx + 1`` (``:482``).
* transform (``:492-498``): key -> idx, anything unknown (or null) -> OOV idx.

The reference iterates ``polars.value_counts`` rows, whose order is not defined
(hard part 6).  This restatement -- and the CUDA kernel --
pin the order of *new* keys to first occurrence in the batch; the key set, the
indices of old keys, index 0 for OOV and ``num_embeddings`` are order free.
polars is not installed in the build container, so this function is "parity
unpinned" against the real ``FeatureTransformer``; it is pinned against the
notebook's growth semantics only.
"""
from __future__ import annotations

import numpy as np


class Vocab:
    def __init__(self, min_freq: int = 0):
        self.min_freq = int(min_freq or 0)
        self.idx_of: dict[int, int] = {}
        self.cnt_of: dict[int, int] = {}
        self.max_idx = 0

    @property
    def num_embeddings(self) -> int:
        return self.max_idx + 1

    def fit(self, keys) -> int:
        """Grow with one batch of keys; returns the new ``num_embeddings``."""
        keys = np.asarray(keys, dtype=np.int64).ravel()
        counts: dict[int, int] = {}
        first: dict[int, int] = {}
        for p, k in enumerate(keys.tolist()):
            if k not in counts:
                counts[k] = 0
                first[k] = p
            counts[k] += 1
        admitted = [k for k in counts if counts[k] >= self.min_freq] if self.min_freq else list(counts)
        admitted.sort(key=lambda k: first[k])
        for k in admitted:
            if k in self.idx_of:
                self.cnt_of[k] += counts[k]
            else:
                self.max_idx += 1
                self.idx_of[k] = self.max_idx
                self.cnt_of[k] = counts[k]
        return self.num_embeddings

    def transform(self, keys) -> np.ndarray:
        keys = np.asarray(keys, dtype=np.int64)
        flat = [self.idx_of.get(k, 0) for k in keys.ravel().tolist()]
        return np.asarray(flat, dtype=np.int32).reshape(keys.shape)
