"""Oracle (test infrastructure): eager fp32 PyTorch models on CPU.

* ``OracleDNN`` restates ``/root/reference/torchctr/models/dnn.py:10-82``: one
  ``nn.Embedding(num_embeddings, emb_dim)`` per ``type == 'sparse'`` feature
  (``:17-23``), dense width = 1 per dense feature, 3 per dense list feature
  (``:24-27``), masked sum pooling per feature in ``feat_configs`` order
  (``:53-59``), ``concat(sparse..., dense_features, *seq_dense_features)``
  (``:61-67``), tower ``[Linear, BatchNorm1d, ReLU, Dropout(0.5)] * k + Linear(., 1)``
  (``:35-46``) and ``binary_cross_entropy_with_logits`` steps (``:72-82``).
  Pinned against the imported reference by ``tests/golden/dnn_golden.pt``.
* ``OracleDeepFM`` / ``OracleDCNv2`` do not exist in the reference
  (SURVEY.md section 0): these are OUR definitions built from the reference's
  parts -- parity unpinned by the reference.
    DeepFM : logit = sum_f w1_f[id] + Linear(dense) + FM2(v) + tower(concat(v, dense)),
             FM2(v) = 0.5 * sum_d[(sum_f v_fd)^2 - sum_f v_fd^2]
    DCN-v2 : x0 = concat(v, dense); x_{l+1} = x0 * (x_l W_l^T + b_l) + x_l  (3 layers),
             logit = tower(x_3)   ("stacked")
Parameter names match ``torchctr_b200.models`` so state dicts can be swapped.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .embedding import pooled_lookup


def split_feature_configs(feat_configs):
    """(sparse configs, dense input width) following dnn.py:17-29."""
    sparse, dense_width = [], 0
    for cfg in feat_configs:
        kind = cfg["type"]
        if kind == "sparse":
            if "emb_dim" not in cfg:
                raise ValueError("emb_dim must be specified for sparse features.")
            sparse.append(cfg)
        elif kind == "dense":
            dense_width += 3 if cfg.get("islist") else 1
        else:
            raise ValueError(f"Unsupported feature type: {kind}")
    return sparse, dense_width


def make_tower(in_features: int, hidden_units, p_drop: float = 0.5) -> nn.Sequential:
    """dnn.py:35-46 -- same module order, hence the same state-dict keys."""
    mods, width = [], in_features
    for h in hidden_units:
        mods += [nn.Linear(width, h), nn.BatchNorm1d(h), nn.ReLU(), nn.Dropout(p=p_drop)]
        width = h
    mods.append(nn.Linear(width, 1))
    return nn.Sequential(*mods)


def fm_second_order(stacked: torch.Tensor) -> torch.Tensor:
    """stacked f32 [B, F, D] -> f32 [B, 1]."""
    s = stacked.sum(dim=1)
    sq = (stacked * stacked).sum(dim=1)
    return 0.5 * (s * s - sq).sum(dim=-1, keepdim=True)


class _OracleBase(nn.Module):
    def __init__(self, feat_configs):
        super().__init__()
        self.feat_configs = feat_configs
        self._sparse, self._dense_width = split_feature_configs(feat_configs)
        self.embeddings = nn.ModuleDict(
            {c["name"]: nn.Embedding(c["num_embeddings"], c["emb_dim"]) for c in self._sparse})
        self._sparse_width = sum(c["emb_dim"] for c in self._sparse)

    def pooled(self, feats):
        return [pooled_lookup(feats[c["name"]], self.embeddings[c["name"]].weight, c.get("pooling", "sum"))
                for c in self._sparse]

    @staticmethod
    def dense_parts(feats):
        parts = []
        if "dense_features" in feats:
            parts.append(feats["dense_features"])
        if "seq_dense_features" in feats:
            parts += list(feats["seq_dense_features"])
        return parts

    def training_step(self, batch, batch_idx):
        feats, labels = batch
        return F.binary_cross_entropy_with_logits(self(feats), labels)

    validation_step = training_step


class OracleDNN(_OracleBase):
    def __init__(self, feat_configs, hidden_units=(256, 128, 64)):
        super().__init__(feat_configs)
        self.tower = make_tower(self._sparse_width + self._dense_width, list(hidden_units))

    def forward(self, feats):
        return self.tower(torch.cat(self.pooled(feats) + self.dense_parts(feats), dim=-1))


class OracleDeepFM(_OracleBase):
    def __init__(self, feat_configs, hidden_units=(256, 128, 64)):
        super().__init__(feat_configs)
        dims = {c["emb_dim"] for c in self._sparse}
        if len(dims) != 1:
            raise ValueError("DeepFM needs one common emb_dim for the FM term")
        self.linear_embeddings = nn.ModuleDict(
            {c["name"]: nn.Embedding(c["num_embeddings"], 1) for c in self._sparse})
        self.linear_dense = nn.Linear(self._dense_width, 1) if self._dense_width else None
        self.tower = make_tower(self._sparse_width + self._dense_width, list(hidden_units))

    def forward(self, feats):
        v = self.pooled(feats)
        dense = self.dense_parts(feats)
        first = sum(pooled_lookup(feats[c["name"]], self.linear_embeddings[c["name"]].weight,
                                  c.get("pooling", "sum")) for c in self._sparse)
        if self.linear_dense is not None:
            first = first + self.linear_dense(torch.cat(dense, dim=-1))
        second = fm_second_order(torch.stack(v, dim=1))
        deep = self.tower(torch.cat(v + dense, dim=-1))
        return first + second + deep


class OracleDCNv2(_OracleBase):
    def __init__(self, feat_configs, hidden_units=(256, 128, 64), num_cross_layers: int = 3):
        super().__init__(feat_configs)
        width = self._sparse_width + self._dense_width
        self.cross = nn.ModuleList([nn.Linear(width, width) for _ in range(num_cross_layers)])
        self.tower = make_tower(width, list(hidden_units))

    def forward(self, feats):
        x0 = torch.cat(self.pooled(feats) + self.dense_parts(feats), dim=-1)
        x = x0
        for layer in self.cross:
            x = x0 * layer(x) + x
        return self.tower(x)
