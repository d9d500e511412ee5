"""Tower / cross-layer GEMMs on tcgen05 (host side of kernel K6).

``linear_tc(x, weight, bias)`` == ``F.linear`` with TF32 inputs and fp32 accumulation, the same numerics
as ``torch.backends.cuda.matmul.allow_tf32 = True``.  Forward and the input gradient run on the
hand-written tcgen05 kernel; the weight gradient (a reduction over the batch, operands MN-major) and the
bias gradient stay on cuBLAS / torch for now.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .. import ops


def tc_eligible(x: torch.Tensor, weight: torch.Tensor) -> bool:
    return (x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1 and x.stride(0) % 4 == 0
            and x.data_ptr() % 16 == 0 and weight.shape[1] % 4 == 0 and weight.shape[0] % 4 == 0 and weight.shape[0] >= 16)


def wgrad_eligible(gy: torch.Tensor, x: torch.Tensor) -> bool:
    ok = lambda t: (t.is_cuda and t.dtype == torch.float32 and t.dim() == 2 and t.stride(1) == 1 and t.stride(0) % 4 == 0
                    and t.data_ptr() % 16 == 0)      # noqa: E731
    return ok(gy) and ok(x) and gy.shape[0] == x.shape[0] and gy.shape[0] >= 1


class _LinearTCFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        w = weight.contiguous()
        ctx.save_for_backward(x, w)
        ctx.has_bias = bias is not None
        return ops.linear_fwd(x, w, bias)

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        gy = gy.contiguous()
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = ops.linear_fwd(gy, w.t().contiguous())          # [M, N] @ [N, K]: W^T is small, transpose it once
        if ctx.needs_input_grad[1]:
            gw = ops.linear_wgrad(gy, x) if wgrad_eligible(gy, x) else gy.t() @ x
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = gy.sum(dim=0)
        return gx, gw, gb


def linear_tc(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None = None) -> torch.Tensor:
    if not tc_eligible(x, weight):
        return F.linear(x, weight, bias)
    return _LinearTCFn.apply(x, weight, bias)
