"""``DNN`` -- drop-in for ``torchctr.models.DNN`` (``torchctr/models/dnn.py:10-82``)."""
from __future__ import annotations

from .base import CTRModelBase, make_tower


class DNN(CTRModelBase):
    def __init__(self, feat_configs, hidden_units=[256, 128, 64], table_device=None):
        super().__init__(feat_configs, table_device)
        self.tower = make_tower(self._sparse_width + self._dense_width, list(hidden_units))

    def forward(self, input_feats):
        self._grow_vocabularies(input_feats)
        self._prepare_tower()
        (x,) = self._lookup_all(input_feats, self.dense_block(input_feats))          # dnn.py:53-67 in one launch
        return self._run_tower(x, blocked_ok=True)                                    # dnn.py:68 (x feeds the tower only)

    def hidden_and_extra(self, input_feats):
        self._grow_vocabularies(input_feats)
        self._prepare_tower()
        (x,) = self._lookup_all(input_feats, self.dense_block(input_feats))
        return self._run_tower(x, stop_before_last=True, blocked_ok=True), None
