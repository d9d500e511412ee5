"""Final ``Linear(H, 1)`` of the tower + extra logit terms + mean ``binary_cross_entropy_with_logits`` as one autograd node
(``torchctr/models/dnn.py:46,68,75``): two small kernels forward, two backward, instead of ~15 element-wise /
reduction launches over ``[B, 1]`` tensors.  Used by ``CTRModelBase.training_step`` on CUDA; ``forward`` of the models
still returns plain logits through the ordinary modules."""
from __future__ import annotations

import torch

from .. import ops


class _LogitBceFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, weight, bias, extra, labels, xe, we, be):
        from .embedding import run_late_starts
        run_late_starts()                   # backward plans left to start here (embedding.PLAN_START = "head")
        w = weight.reshape(-1).contiguous()
        we_flat = None if we is None else we.reshape(-1).contiguous()
        loss, dz, _ = ops.logit_bce_fwd(h, w, bias, extra, labels, xe=xe, we=we_flat, be=be)
        ctx.save_for_backward(h, w, dz, xe)
        ctx.has_extra = extra is not None
        ctx.has_bias = bias is not None
        ctx.has_be = be is not None
        ctx.wshape = tuple(weight.shape)
        ctx.params = (weight, bias, we, be)
        ctx.weshape = None if we is None else tuple(we.shape)
        return loss

    @staticmethod
    def backward(ctx, gloss):
        from . import tower as _tower
        h, w, dz, xe = ctx.saved_tensors
        g = gloss.reshape(1).to(torch.float32).contiguous()
        params = [p for p in ctx.params if p is not None]
        # the parameter gradients (last Linear, Linear(dense)) are read by optimizer.step() only: their finalisation runs on the
        # second stream (nn/tower.py) when autograd takes the tensors over as they are
        defer = _tower.defer_weight_grads and all(_tower._takes_over(p) for p in params)
        gh, gw, gb, gextra, gwe = ops.logit_bce_bwd(h, w, dz, g, ctx.needs_input_grad[0],
                                                    ctx.has_extra and ctx.needs_input_grad[3], xe=xe,
                                                    defer_params_tag="head" if defer else None)
        gbe = None
        if defer:
            dev = h.device
            main, side, engine_joins = _tower._defer_begin(dev)
            _tower._deferred[dev][1] += [g, h]
            side.wait_stream(main)
            weight, bias, we, be = ctx.params
            with torch.cuda.stream(side):
                gw, gb, gwe = ops.logit_bce_bwd_params(h, g, 0 if xe is None else xe.shape[1], "head")
                if ctx.has_be:
                    gbe = _tower._deliver(be, gb.clone())
                gw = _tower._deliver(weight, gw.view(ctx.wshape))
                gb = _tower._deliver(bias, gb) if ctx.has_bias else None
                gwe = _tower._deliver(we, gwe.view(ctx.weshape)) if gwe is not None else None
            if not engine_joins:
                _tower.join_deferred(dev)
            return gh, gw, gb, gextra, None, None, gwe, gbe
        if ctx.has_be:
            gbe = gb.clone()
        return (gh, gw.view(ctx.wshape), (gb if ctx.has_bias else None), gextra, None, None,
                (gwe.view(ctx.weshape) if gwe is not None else None), gbe)


def head_eligible(h: torch.Tensor, linear: torch.nn.Linear, labels: torch.Tensor) -> bool:
    H = h.shape[1] if h.dim() == 2 else 0
    return (h.is_cuda and h.dtype == torch.float32 and h.dim() == 2 and h.stride(1) == 1 and h.stride(0) % 4 == 0
            and h.data_ptr() % 16 == 0 and linear.out_features == 1 and 4 <= H <= 128 and H & (H - 1) == 0
            and labels.dtype == torch.float32 and labels.shape[0] == h.shape[0]
            and (labels.dim() == 1 or labels.shape[1] == 1))


def second_term_eligible(xe, linear_e) -> bool:
    """``linear_e(xe)`` can ride inside the head kernels: a raw f32 [B, ne <= 32] block that needs no gradient."""
    return (xe is not None and linear_e is not None and xe.is_cuda and xe.dtype == torch.float32 and xe.dim() == 2
            and xe.stride(1) == 1 and 1 <= xe.shape[1] <= 32 and not xe.requires_grad and linear_e.out_features == 1
            and linear_e.in_features == xe.shape[1])


def logit_bce(h, linear: torch.nn.Linear, extra, labels, xe=None, linear_e=None) -> torch.Tensor:
    """mean BCE-with-logits of ``linear(h) + extra (+ linear_e(xe))`` against ``labels`` ([B] or [B, 1])."""
    if extra is not None and (extra.dim() != 2 or extra.stride(1) != 1 and extra.shape[1] != 1):
        extra = extra.reshape(-1, 1).contiguous()
    lab = labels.reshape(labels.shape[0], -1)
    if xe is None:
        return _LogitBceFn.apply(h, linear.weight, linear.bias, extra, lab, None, None, None)
    return _LogitBceFn.apply(h, linear.weight, linear.bias, extra, lab, xe, linear_e.weight, linear_e.bias)
