"""``DCNv2`` (stacked) on the fused lookup.  Not in the reference; definition of SURVEY.md 8c /
``oracle.models.OracleDCNv2``:  x_{l+1} = x0 * (x_l W_l^T + b_l) + x_l,  logit = tower(x_L)."""
from __future__ import annotations

import torch.nn as nn

from ..nn.interaction import CrossLayer
from .base import CTRModelBase, make_tower


class DCNv2(CTRModelBase):
    def __init__(self, feat_configs, hidden_units=[256, 128, 64], num_cross_layers: int = 3, table_device=None):
        super().__init__(feat_configs, table_device)
        width = self._sparse_width + self._dense_width
        self.cross = nn.ModuleList([CrossLayer(width, width) for _ in range(num_cross_layers)])
        self.tower = make_tower(width, list(hidden_units))

    def _crossed(self, input_feats):
        self._grow_vocabularies(input_feats)
        self._prepare_tower()
        (x0,) = self._lookup_all(input_feats, self.dense_block(input_feats))
        x = x0
        for layer in self.cross:
            x = layer(x0, x)
        return x

    def forward(self, input_feats):
        return self._run_tower(self._crossed(input_feats))

    def hidden_and_extra(self, input_feats):
        return self._run_tower(self._crossed(input_feats), stop_before_last=True), None
