"""``torchctr.nn.functional`` pieces on the device kernels: ``target_attention`` (``torchctr/nn/functional.py:46-74``)."""
from __future__ import annotations

import torch

from .. import ops


class _TargetAttentionFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, target, cand, mask, honor_mask):
        target, cand = target.contiguous(), cand.contiguous()
        mask = None if mask is None else mask.to(torch.float32).contiguous()
        out, scores, row_max, row_sum = ops.target_attention_fwd(target, cand, mask, honor_mask)
        ctx.save_for_backward(target, cand, mask, out, scores, row_max, row_sum)
        ctx.honor_mask = honor_mask
        return out

    @staticmethod
    def backward(ctx, gout):
        target, cand, mask, out, scores, row_max, row_sum = ctx.saved_tensors
        gt, gc = ops.target_attention_bwd(target, cand, mask, ctx.honor_mask, out, scores, row_max, row_sum, gout.contiguous())
        return gt, gc, None, None


def target_attention(target_emb: torch.Tensor, candidate_embs: torch.Tensor, mask: torch.Tensor | None = None,
                     honor_mask: bool = False) -> torch.Tensor:
    """``torchctr.nn.functional.target_attention``: target [B, E], candidates [B, N, E], mask [B, N] (0 = masked) -> [B, E],
    as ONE kernel forward and one backward (no [B, N] / [B, N, E] temporaries besides the saved scores).

    ``honor_mask=False`` (default) is bit-for-bit the reference's behaviour: its ``masked_fill`` is not in place
    (``functional.py:63``), so the mask is ignored.  ``honor_mask=True`` applies it (masked candidates get weight 0)."""
    if not target_emb.is_cuda:
        raise RuntimeError("torchctr_b200: target_attention needs CUDA tensors -- the kernels have no CPU path")
    if candidate_embs.dim() != 3 or target_emb.dim() != 2 or candidate_embs.shape[0] != target_emb.shape[0] \
            or candidate_embs.shape[2] != target_emb.shape[1]:
        raise ValueError(f"shapes: target {tuple(target_emb.shape)} vs candidates {tuple(candidate_embs.shape)}")
    if candidate_embs.shape[1] == 0 or candidate_embs.shape[0] == 0:      # no candidates: the sum over an empty sequence
        return target_emb.float() * 0.0 + candidate_embs.float().sum(dim=1)
    return _TargetAttentionFn.apply(target_emb.float(), candidate_embs.float(), mask, bool(honor_mask))
