"""Tower / cross-layer GEMMs on tcgen05 (host side of kernel K6).

``linear_tc(x, weight, bias)`` == ``F.linear``; forward, input gradient and weight gradient all run on the hand-written
tcgen05 kernels (``ctr_linear_fwd`` / ``ctr_linear_wgrad``) in one of two precisions:

``"tf32"``    operands read as TF32 (10-bit mantissa), fp32 accumulate in TMEM -- the numerics of
              ``torch.backends.cuda.matmul.allow_tf32 = True``; chosen when that flag is on.
``"tf32x3"``  error-compensated: every operand is split into two TF32 parts (``ctr_split_tf32``) and the same kernels
              contract over lo.hi + hi.lo + hi.hi -- fp32-grade results (~1e-6 relative), the exact mode the 1e-5
              parity bound of the tower (``torchctr/models/dnn.py:35-46``) is tested in; chosen when ``allow_tf32`` is off.

``set_matmul_precision`` overrides the choice.  There is no library GEMM on this path: shapes the kernels cannot take
(fewer than 16 output features, e.g. the final ``Linear(H, 1)`` in eval mode) fall back to ``F.linear``.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .. import ops

_PRECISION: str | None = None


def set_matmul_precision(mode: str | None) -> None:
    """``"tf32"`` | ``"tf32x3"`` | ``None`` (follow ``torch.backends.cuda.matmul.allow_tf32``)."""
    global _PRECISION
    if mode not in (None, "tf32", "tf32x3"):
        raise ValueError(f"unknown matmul precision {mode!r}")
    _PRECISION = mode


def matmul_precision() -> str:
    if _PRECISION is not None:
        return _PRECISION
    return "tf32" if torch.backends.cuda.matmul.allow_tf32 else "tf32x3"


def _rows_ok(t: torch.Tensor) -> bool:
    return t.is_cuda and t.dtype == torch.float32 and t.dim() == 2 and t.stride(1) == 1


def _tma_ok(t: torch.Tensor) -> bool:
    return _rows_ok(t) and t.stride(0) % 4 == 0 and t.data_ptr() % 16 == 0


def tc_eligible(x: torch.Tensor, weight: torch.Tensor, precision: str | None = None) -> bool:
    """Can ``x @ weight.T`` (and its gradients) run on the tcgen05 kernels?  ``tf32`` feeds the tensors to TMA as they
    are (16-byte aligned rows); ``tf32x3`` goes through the split kernel, which takes any f32 matrix."""
    precision = precision or matmul_precision()
    if not (_rows_ok(x) and weight.is_cuda and weight.dtype == torch.float32 and weight.dim() == 2 and weight.shape[0] >= 16):
        return False
    if precision == "tf32x3":
        return True
    return _tma_ok(x) and weight.shape[1] % 4 == 0 and weight.shape[0] % 4 == 0


def wgrad_eligible(gy: torch.Tensor, x: torch.Tensor, precision: str | None = None) -> bool:
    precision = precision or matmul_precision()
    ok = _rows_ok if precision == "tf32x3" else _tma_ok
    return ok(gy) and ok(x) and gy.shape[0] == x.shape[0] and gy.shape[0] >= 1


def gemm_nt(a: torch.Tensor, w: torch.Tensor, bias=None, act: int = 0, out=None, precision: str = "tf32",
            out_block: int = 0) -> torch.Tensor:
    """``act(a @ w.T + bias)`` on tcgen05: a [M, K], w [N, K] (``out_block``: column-blocked result, ``ops.linear_fwd``)."""
    if precision == "tf32x3":
        return ops.linear_fwd(ops.split_tf32(a, 1, 0), ops.split_tf32(w, 1, 1), bias, act, out=out, out_block=out_block)
    return ops.linear_fwd(a, w, bias, act, out=out, out_block=out_block)


def gemm_nt_bn_stats(a, w, bias, bn, precision: str = "tf32"):
    """``z = a @ w.T + bias`` plus the BatchNorm1d batch statistics of ``z`` from the GEMM epilogue -> (z, mean, rstd) or None."""
    track = bn.track_running_stats and bn.running_mean is not None
    rm, rv, nbt = (bn.running_mean, bn.running_var, bn.num_batches_tracked) if track else (None, None, None)
    if precision == "tf32x3":
        a, w = ops.split_tf32(a, 1, 0), ops.split_tf32(w, 1, 1)
    return ops.linear_fwd_bn_stats(a, w, bias, bn.eps, bn.momentum, rm, rv, nbt)


def gemm_wgrad(g: torch.Tensor, x: torch.Tensor, precision: str = "tf32") -> torch.Tensor:
    """``g.T @ x`` -> [N, K] on tcgen05: g [B, N], x [B, K]."""
    if precision == "tf32x3":      # the split pads the rows to a multiple of 4 floats: hand the kernel views of the real width
        return ops.linear_wgrad(ops.split_tf32(g, 0, 0)[:, :g.shape[1]], ops.split_tf32(x, 0, 1)[:, :x.shape[1]])
    return ops.linear_wgrad(g, x)


class _LinearTCFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, precision):
        w = weight.contiguous()
        ctx.save_for_backward(x, w)
        ctx.has_bias = bias is not None
        ctx.precision = precision
        return gemm_nt(x, w, bias, precision=precision)

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        prec = ctx.precision
        gy = gy.contiguous()
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gx = gemm_nt(gy, w.t().contiguous(), precision=prec)   # [M, N] @ [N, K]: W^T is small, transpose it once
        if ctx.needs_input_grad[1]:
            gw = gemm_wgrad(gy, x, prec) if wgrad_eligible(gy, x, prec) else gy.t() @ x
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = gy.sum(dim=0)
        return gx, gw, gb, None


def linear_tc(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor | None = None,
              precision: str | None = None) -> torch.Tensor:
    precision = precision or matmul_precision()
    if not tc_eligible(x, weight, precision):
        return F.linear(x, weight, bias)
    return _LinearTCFn.apply(x, weight, bias, precision)
