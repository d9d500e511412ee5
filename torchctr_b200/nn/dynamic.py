"""``DynamicEmbedding`` -- drop-in for ``torchctr.nn.DynamicEmbedding``
(``torchctr/nn/embedding.py:63-95``): an ``nn.Embedding`` that grows to ``max(id) + 1`` rows on
the fly (new rows ~ N(0, 0.01)) and merges checkpoints of a different size.

Differences in mechanism, not in behaviour: the table lives in a capacity-managed HBM store
(amortised growth instead of ``torch.cat`` of the whole table per new id, ``:77``), min / max
of the ids come from one device reduction and one host read (two ``.item()`` syncs upstream,
``:83-85``), and the gather / its sparse backward are libctr_b200 kernels.
"""
from __future__ import annotations

import torch

from .. import ops
from .embedding import EmbeddingTable


class DynamicEmbedding(EmbeddingTable):
    def __init__(self, num_embeddings, embedding_dim, padding_idx=None, max_norm=None, norm_type=2.0,
                 scale_grad_by_freq=False, sparse=False, _weight=None, **kw):
        super().__init__(num_embeddings, embedding_dim, padding_idx, max_norm, norm_type, scale_grad_by_freq, sparse,
                         _weight, **kw)

    def _expand_embeddings(self, new_num_embeddings):          # embedding.py:69-78
        self.grow_to(new_num_embeddings)

    def forward(self, input):
        if input.numel() == 0:
            raise ValueError("Indices tensor is empty")        # embedding.py:81
        ids = input.to(self.weight.device, dtype=torch.int64, non_blocking=True).contiguous()
        lo, hi = ops.ids_minmax(ids).tolist()
        if lo < 0:
            raise ValueError("Indices contain negative values")  # embedding.py:83-84
        self._expand_embeddings(hi + 1)
        return super().forward(ids)

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        key = prefix + "weight"
        if key in state_dict:                                  # embedding.py:89-95
            rows = state_dict[key].size(0)
            self._expand_embeddings(rows)
            if rows < self.num_embeddings:
                tail = self.weight.detach()[rows:].to(state_dict[key].device)
                state_dict[key] = torch.cat([state_dict[key], tail], dim=0)
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)
