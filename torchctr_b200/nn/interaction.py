"""FM second-order interaction and the DCN-v2 cross layer (host side of kernels K4 / K5).

Neither exists in the reference (``torchctr/models/__init__.py:1-2`` exports only ``DNN``);
the formulas are the ones stated in SURVEY.md section 8c and restated in ``oracle/models.py``.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from .linear import linear_tc


class _FMFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, first, num_fields, dim, num_first):
        ctx.save_for_backward(x)
        ctx.meta = (num_fields, dim, num_first, first is not None, None if first is None else first.shape)
        f = None if first is None else first[:, :num_first]
        return ops.fm_fwd(x, num_fields, dim, first=f)

    @staticmethod
    def backward(ctx, gout):
        (x,) = ctx.saved_tensors
        num_fields, dim, num_first, has_first, first_shape = ctx.meta
        gout = gout.contiguous()
        gx = torch.zeros_like(x) if x.shape[1] > num_fields * dim else torch.empty_like(x)
        gfirst = None
        if has_first:
            gfirst = torch.zeros(first_shape, dtype=torch.float32, device=x.device) \
                if first_shape[1] > num_first else torch.empty(first_shape, dtype=torch.float32, device=x.device)
        ops.fm_bwd(x, num_fields, dim, gout, gx, False, None if gfirst is None else gfirst[:, :num_first])
        return gx, gfirst, None, None, None


def fm_interaction(x: torch.Tensor, num_fields: int, dim: int, first: torch.Tensor | None = None,
                   num_first: int = 0) -> torch.Tensor:
    """x f32 [B, >= F*D] (fields in the leading columns) -> [B, 1]:
    0.5 * sum_d[(sum_f v)^2 - sum_f v^2]  (+ the row sums of first[:, :num_first])."""
    return _FMFn.apply(x, first, num_fields, dim, num_first)


class _FMPassFn(torch.autograd.Function):
    """FM term + pass-through of x: the tower reads the alias ``x_t``, so in backward the gradient of
    the tower arrives here and the FM gradient is ACCUMULATED into it in place -- no zero-filled
    gradient buffer and no extra ``add`` pass over [B, F*D]."""

    @staticmethod
    def forward(ctx, x, first, num_fields, dim, num_first):
        ctx.save_for_backward(x)
        ctx.meta = (num_fields, dim, num_first, None if first is None else first.shape)
        f = None if first is None else first[:, :num_first]
        return x, ops.fm_fwd(x, num_fields, dim, first=f)

    @staticmethod
    def backward(ctx, g_x, g_fm):
        (x,) = ctx.saved_tensors
        num_fields, dim, num_first, first_shape = ctx.meta
        gfirst = None
        if first_shape is not None:
            gfirst = torch.zeros(first_shape, dtype=torch.float32, device=x.device) \
                if first_shape[1] > num_first else torch.empty(first_shape, dtype=torch.float32, device=x.device)
        if g_fm is None:
            return g_x, (None if gfirst is None else gfirst.zero_()), None, None, None
        g_fm = g_fm.contiguous()
        if g_x is None:
            g_x = torch.zeros_like(x)
        elif g_x.stride(1) != 1 or g_x.stride(0) != x.shape[1]:
            g_x = g_x.contiguous()
        ops.fm_bwd(x, num_fields, dim, g_fm, g_x, True, None if gfirst is None else gfirst[:, :num_first])
        return g_x, gfirst, None, None, None


def fm_interaction_passthrough(x, num_fields: int, dim: int, first=None, num_first: int = 0):
    """Returns ``(x_alias, fm)``: feed ``x_alias`` to every other consumer of ``x`` (the tower)."""
    return _FMPassFn.apply(x, first, num_fields, dim, num_first)


class _CrossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x0, x, u, bias):
        ctx.save_for_backward(x0, u, bias)
        return ops.cross_combine_fwd(x0, x, u, bias)

    @staticmethod
    def backward(ctx, gy):
        x0, u, bias = ctx.saved_tensors
        gy = gy.contiguous()
        gx0 = torch.empty_like(x0)
        gu = ops.cross_combine_bwd(x0, u, bias, gy, gx0, False)
        return gx0, gy, gu, gu.sum(dim=0)


class CrossLayer(nn.Linear):
    """DCN-v2 cross layer  y = x0 * (x W^T + b) + x  over rows padded to a multiple of 4 floats.
    An ``nn.Linear(d, d)`` (same parameters / state-dict keys as ``oracle.models.OracleDCNv2``)."""

    def forward(self, x0: torch.Tensor, x: torch.Tensor) -> torch.Tensor:  # type: ignore[override]
        pad = x.shape[1] - self.in_features
        w = F.pad(self.weight, (0, pad, 0, pad)) if pad else self.weight
        b = F.pad(self.bias, (0, pad)) if pad else self.bias
        # dense contraction on the tcgen05 kernel: TF32 when TF32 matmuls are allowed, 3xTF32 (fp32-grade) otherwise
        u = linear_tc(x, w, None)
        return _CrossFn.apply(x0, x, u, b)
