"""``DeepFM`` on the fused lookup.  Not in the reference; definition of SURVEY.md 8c /
``oracle.models.OracleDeepFM``:
    logit = sum_f w1_f[id] + Linear(dense) + FM2(v) + tower(concat(v, dense))."""
from __future__ import annotations

import torch
import torch.nn as nn

from ..nn.embedding import EmbeddingTable, PlanLink, PooledLookupGroup
from ..nn.head import second_term_eligible
from ..nn.interaction import fm_interaction_passthrough
from .base import CTRModelBase, make_tower


class DeepFM(CTRModelBase):
    _fm_term = True          # the lookup may produce the FM second-order term (parallel.shard_model reads this)

    def __init__(self, feat_configs, hidden_units=[256, 128, 64], table_device=None):
        super().__init__(feat_configs, table_device)
        dims = {c["emb_dim"] for c in self._sparse}
        if len(dims) != 1:
            raise ValueError("DeepFM needs one common emb_dim for the FM term")
        self._dim = dims.pop()
        self.linear_embeddings = nn.ModuleDict()
        for c in self._sparse:
            src = self.embeddings[c["name"]]
            self.linear_embeddings[c["name"]] = EmbeddingTable(
                c["num_embeddings"], 1, pooling=src.pooling, index_kind=src.index_kind, hash_seed=src.hash_seed,
                vocab=src.vocab, use_id_weight=src.use_id_weight, **({} if table_device is None else {"device": table_device}))
        self._linear_lookup = PooledLookupGroup(self._names, self.linear_embeddings)
        self._groups.append(self._linear_lookup)
        self.linear_dense = nn.Linear(self._dense_width, 1) if self._dense_width else None
        self.tower = make_tower(self._sparse_width + self._dense_width, list(hidden_units))

    def _parts(self, input_feats, defer_dense_linear: bool = False):
        """(tower input x, the logit terms outside the tower: FM second order + first order (+ Linear on dense)).
        ``defer_dense_linear``: leave Linear(dense) to the caller when the fused logit head can take it
        (returns (x, extra, dense block | None))."""
        self._grow_vocabularies(input_feats)
        self._prepare_tower()
        dense = self.dense_block(input_feats)
        twins = None
        self._x_tower_only = False
        if self._sharded is None:                             # (sharded: the tables live in the shards, not in the ModuleDicts)
            twins = [self.linear_embeddings[n] for n in self._names]
            if self.training and torch.is_grad_enabled():
                for g in self._groups:
                    g.autobind()
        if self._sharded is not None and getattr(self._sharded, "fused_extra", False):
            # hybrid placement: the same fused terms, small tables replicated / large ones read over NVLink inside ONE lookup
            x, extra = self._sharded(input_feats, dense)
            extra = extra.unsqueeze(1)
        elif twins is not None and self._lookup.fused_extra_eligible(input_feats, twins, self.training):
            # Criteo-shaped (single-id, one width): the first-order weights and the FM term ride inside the lookup
            # kernel, their gradients inside the fused update -- no D = 1 launch group, no pass over [B, F * D] for FM
            link = PlanLink() if self.training else None
            x, extra = self._lookup(input_feats, dense, self.training, link, twins=twins, fm=True)
            extra = extra.unsqueeze(1)
            self._x_tower_only = True                         # the FM term came out of the lookup: only the tower reads x
        else:
            x, first = self._lookup_all(input_feats, dense)   # [B, pad4(F*D + Nd)], [B, pad4(F)]; one backward sort
            nf = len(self._names)
            x, extra = fm_interaction_passthrough(x, nf, self._dim, first, nf)
        if self.linear_dense is not None and defer_dense_linear:
            xe = dense.to(x.device, non_blocking=True)
            if second_term_eligible(xe, self.linear_dense):
                return x, extra, xe
        if self.linear_dense is not None:
            # the dense block itself, not the slice of x: same numbers, but no third gradient stream into x (autograd
            # would sum it with the tower's out of place: a zero-fill and an add over [B, F*D + Nd] per step)
            extra = extra + self.linear_dense(dense.to(x.device, dtype=torch.float32, non_blocking=True))
        return (x, extra, None) if defer_dense_linear else (x, extra)

    def forward(self, input_feats):
        x, extra = self._parts(input_feats)
        return extra + self._run_tower(x, blocked_ok=self._x_tower_only)

    def hidden_and_extra(self, input_feats):
        """(h, extra, (dense block, its Linear) | None): the third item is the Linear(dense) term left to the head kernels"""
        x, extra, xe = self._parts(input_feats, defer_dense_linear=True)
        return (self._run_tower(x, stop_before_last=True, blocked_ok=self._x_tower_only), extra,
                (None if xe is None else (xe, self.linear_dense)))
