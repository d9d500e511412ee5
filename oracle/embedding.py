"""Oracle (test infrastructure): embedding gather / pool / gradient / growth on CPU.

fp32 eager PyTorch on the host -- the same ATen operators the reference
dispatches to, so this *is* the reference arithmetic, restated as functions:

* ``pooled_lookup``       -- ``/root/reference/torchctr/models/dnn.py:53-59``
  (pad mask ``id >= 0``, pads read row 0, masked product, ``sum(dim=1)``).
  ``mode='mean'`` and ``per_id_weight`` are extensions:
  parity unpinned by the reference.
* ``dense_table_grad``    -- what autograd produces for the table behind that
  expression (``aten::embedding_dense_backward``): a dense ``[V, D]`` tensor.
* ``unique_row_grads``    -- the same gradient restricted to touched rows
  (sorted unique row ids + their summed gradient).
* ``grow_table`` / ``merge_smaller_checkpoint`` --
  ``/root/reference/torchctr/nn/embedding.py:69-95`` (``DynamicEmbedding``).
"""
from __future__ import annotations

import torch


def _valid(ids: torch.Tensor) -> torch.Tensor:
    return ids >= 0


def bag_scale(ids: torch.Tensor, mode: str) -> torch.Tensor:
    """Per-bag multiplier: 1 for sum; 1 / max(valid_count, 1) for mean."""
    if mode == "sum":
        return torch.ones(ids.shape[0], dtype=torch.float32)
    if mode == "mean":
        return 1.0 / _valid(ids).sum(dim=1).clamp(min=1).to(torch.float32)
    raise ValueError(f"unknown pooling mode {mode!r}")


def pooled_lookup(ids: torch.Tensor, weight: torch.Tensor, mode: str = "sum",
                  per_id_weight: torch.Tensor | None = None) -> torch.Tensor:
    """ids i64 [B, L] (negative = pad), weight f32 [V, D] -> f32 [B, D]."""
    keep = _valid(ids)
    safe = torch.where(keep, ids, torch.zeros_like(ids))      # pads read row 0 (dnn.py:56)
    rows = weight[safe]                                        # [B, L, D]   (dnn.py:57)
    coef = keep.to(weight.dtype)
    if per_id_weight is not None:
        coef = coef * per_id_weight
    pooled = (rows * coef.unsqueeze(-1)).sum(dim=1)            # (dnn.py:57-58)
    if mode != "sum":
        pooled = pooled * bag_scale(ids, mode).unsqueeze(-1)
    return pooled


def dense_table_grad(ids: torch.Tensor, grad_out: torch.Tensor, num_rows: int, mode: str = "sum",
                     per_id_weight: torch.Tensor | None = None) -> torch.Tensor:
    """d loss / d table as the dense [V, D] tensor autograd would build."""
    B, L = ids.shape
    keep = _valid(ids)
    coef = keep.to(grad_out.dtype)
    if per_id_weight is not None:
        coef = coef * per_id_weight
    coef = coef * bag_scale(ids, mode).unsqueeze(-1)
    contrib = grad_out.unsqueeze(1) * coef.unsqueeze(-1)      # [B, L, D]
    safe = torch.where(keep, ids, torch.zeros_like(ids))
    dense = torch.zeros(num_rows, grad_out.shape[1], dtype=grad_out.dtype)
    dense.index_add_(0, safe.reshape(-1), contrib.reshape(B * L, -1))
    return dense


def unique_row_grads(ids: torch.Tensor, grad_out: torch.Tensor, num_rows: int, mode: str = "sum",
                     per_id_weight: torch.Tensor | None = None):
    """(sorted unique valid row ids i64 [U], summed gradient f32 [U, D])."""
    dense = dense_table_grad(ids, grad_out, num_rows, mode, per_id_weight)
    rows = torch.unique(ids[_valid(ids)])                      # sorted
    return rows, dense[rows]


def grow_table(weight: torch.Tensor, max_index: int, generator: torch.Generator | None = None):
    """``DynamicEmbedding._expand_embeddings`` (nn/embedding.py:69-78).

    Rows ``[old_V, max_index]`` are appended, drawn N(0, 0.01); existing rows keep
    their values.  Returns ``(new_weight, new_rows_only)``.
    """
    old_rows, dim = weight.shape
    if max_index + 1 <= old_rows:
        return weight, weight.new_empty(0, dim)
    fresh = torch.empty(max_index + 1 - old_rows, dim, dtype=weight.dtype)
    fresh.normal_(mean=0.0, std=0.01, generator=generator)
    return torch.cat([weight, fresh], dim=0), fresh


def check_dynamic_ids(ids: torch.Tensor) -> None:
    """Input validation of ``DynamicEmbedding.forward`` (nn/embedding.py:80-84)."""
    if ids.numel() == 0:
        raise ValueError("Indices tensor is empty")
    if int(ids.min()) < 0:
        raise ValueError("Indices contain negative values")


def merge_smaller_checkpoint(current: torch.Tensor, ckpt: torch.Tensor) -> torch.Tensor:
    """``DynamicEmbedding._load_from_state_dict`` (nn/embedding.py:89-95).

    A checkpoint with more rows than the module grows the module first; one with
    fewer rows is padded with the module's current tail rows.  Returns the weight
    the module holds after ``load_state_dict``.
    """
    if ckpt.shape[0] >= current.shape[0]:
        return ckpt.clone()
    return torch.cat([ckpt, current[ckpt.shape[0]:]], dim=0)
