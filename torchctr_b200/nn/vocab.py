"""Device vocabulary: raw key -> table row, growing while training (host side of kernel K3).

Restates the vocabulary of ``torchctr/transformer.py:451-498`` as a module whose state (an
open-addressing hash map in HBM + the next free row) is checkpointed with the model: row 0 is
the OOV row (``:476-477``), new keys take rows ``max_idx + 1, ...`` (``:462-472``),
``num_embeddings = next_row`` (``:482``).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import _lib, ops


class VocabIndex(nn.Module):
    def __init__(self, capacity: int = 1 << 16, min_freq: int = 0):
        super().__init__()
        cap = 64
        while cap < capacity:
            cap <<= 1
        self.min_freq = int(min_freq or 0)
        self.register_buffer("map_keys", torch.full((cap,), _lib.VOCAB_EMPTY, dtype=torch.int64))
        self.register_buffer("map_rows", torch.zeros(cap, dtype=torch.int32))
        self.register_buffer("next_row", torch.ones(1, dtype=torch.int64))      # row 0 = OOV
        self._bound = 0          # host-side upper bound on the number of stored keys

    @property
    def capacity(self) -> int:
        return self.map_keys.numel()

    def handle(self) -> ops.VocabMapHandle:
        return ops.VocabMapHandle(self.map_keys, self.map_rows)

    def num_embeddings(self) -> int:
        """Rows in use incl. the OOV row (one device -> host read)."""
        return int(self.next_row.item())

    def _reserve(self, incoming: int):
        if (self._bound + incoming) * 2 <= self.capacity:
            return
        self._bound = self.num_embeddings() - 1
        if (self._bound + incoming) * 2 <= self.capacity:
            return
        cap = self.capacity
        while (self._bound + incoming) * 2 > cap:
            cap <<= 1
        live = self.map_keys != _lib.VOCAB_EMPTY
        keys, rows = self.map_keys[live].contiguous(), self.map_rows[live].contiguous()
        self.map_keys = torch.full((cap,), _lib.VOCAB_EMPTY, dtype=torch.int64, device=keys.device)
        self.map_rows = torch.zeros(cap, dtype=torch.int32, device=keys.device)
        ops.vocab_insert(self.handle(), keys, rows)

    def fit(self, keys: torch.Tensor) -> None:
        """Admit the new keys of one batch (negative keys are padding)."""
        keys = keys.to(self.map_keys.device, dtype=torch.int64, non_blocking=True).contiguous()
        self._reserve(keys.numel())
        status = torch.zeros(1, dtype=torch.int32, device=keys.device)
        ops.vocab_fit(self.handle(), keys.reshape(-1), self.next_row, self.min_freq, status=status)
        self._bound += keys.numel()
        self._last_status = status

    def fit_and_grow(self, keys: torch.Tensor, table) -> int:
        self.fit(keys)
        n = self.num_embeddings()
        if int(self._last_status.item()) & _lib.STATUS_MAP_FULL:
            raise RuntimeError("vocabulary map overflow")
        table.grow_to(n)
        return n

    def transform(self, keys: torch.Tensor, oov_row: int = 0) -> torch.Tensor:
        keys = keys.to(self.map_keys.device, dtype=torch.int64, non_blocking=True).contiguous()
        return ops.vocab_transform(self.handle(), keys, oov_row)

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        k = prefix + "map_keys"
        if k in state_dict and state_dict[k].numel() != self.capacity:
            dev = self.map_keys.device
            self.map_keys = torch.empty_like(state_dict[k], device=dev)
            self.map_rows = torch.empty_like(state_dict[prefix + "map_rows"], device=dev)
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)
        self._bound = self.capacity // 2     # unknown until the next exact read
