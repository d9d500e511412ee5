"""Final ``Linear(H, 1)`` of the tower + extra logit terms + mean ``binary_cross_entropy_with_logits`` as one autograd node
(``torchctr/models/dnn.py:46,68,75``): two small kernels forward, two backward, instead of ~15 element-wise /
reduction launches over ``[B, 1]`` tensors.  Used by ``CTRModelBase.training_step`` on CUDA; ``forward`` of the models
still returns plain logits through the ordinary modules."""
from __future__ import annotations

import torch

from .. import ops


class _LogitBceFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, weight, bias, extra, labels):
        w = weight.reshape(-1).contiguous()
        loss, dz, _ = ops.logit_bce_fwd(h, w, bias, extra, labels)
        ctx.save_for_backward(h, w, dz)
        ctx.has_extra = extra is not None
        ctx.has_bias = bias is not None
        ctx.wshape = tuple(weight.shape)
        return loss

    @staticmethod
    def backward(ctx, gloss):
        h, w, dz = ctx.saved_tensors
        g = gloss.reshape(1).to(torch.float32).contiguous()
        gh, gw, gb, gextra = ops.logit_bce_bwd(h, w, dz, g, ctx.needs_input_grad[0], ctx.has_extra and ctx.needs_input_grad[3])
        return gh, gw.view(ctx.wshape), (gb if ctx.has_bias else None), gextra, None


def head_eligible(h: torch.Tensor, linear: torch.nn.Linear, labels: torch.Tensor) -> bool:
    H = h.shape[1] if h.dim() == 2 else 0
    return (h.is_cuda and h.dtype == torch.float32 and h.dim() == 2 and h.stride(1) == 1 and h.stride(0) % 4 == 0
            and h.data_ptr() % 16 == 0 and linear.out_features == 1 and 4 <= H <= 128 and H & (H - 1) == 0
            and labels.dtype == torch.float32 and labels.shape[0] == h.shape[0])


def logit_bce(h, linear: torch.nn.Linear, extra, labels) -> torch.Tensor:
    """mean BCE-with-logits of ``linear(h) + extra`` against ``labels`` ([B] or [B, 1])."""
    if extra is not None and (extra.dim() != 2 or extra.stride(1) != 1 and extra.shape[1] != 1):
        extra = extra.reshape(-1, 1).contiguous()
    lab = labels.reshape(labels.shape[0], -1)
    return _LogitBceFn.apply(h, linear.weight, linear.bias, extra, lab)
