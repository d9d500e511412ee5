"""One block of the reference tower -- Linear, BatchNorm1d, ReLU, Dropout (``torchctr/models/dnn.py:39-45``) -- as a
single autograd node in training mode (host side of kernels K6 / K6b).

forward   z = x W^T + b (tcgen05 kernel: TF32, or error-compensated 3xTF32 when TF32 matmuls are not allowed)
          -> batch statistics + running-statistics update (``ctr_bn_stats``)
          -> y = dropout(relu(batchnorm(z))) in one pass (``ctr_bn_relu_dropout_fwd``)
backward  one reduction pass + one apply pass give dL/dz, dgamma, dbeta and the bias gradient
          (``ctr_bn_relu_dropout_bwd``); dL/dx on the tcgen05 kernel; dL/dW = gz^T x.
The dropout mask is recomputed from (step seed on the device, layer id, element), never stored.  The modules keep
their parameters, buffers and ``state_dict`` keys; in eval mode they run as plain torch modules.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from .linear import gemm_nt, gemm_nt_bn_stats, gemm_wgrad, matmul_precision, tc_eligible, wgrad_eligible


class _TowerBlockFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, gamma, beta, bn, p_drop, seed_dev, layer_id, precision, gx_provider=None):
        w = weight.contiguous()
        tc = tc_eligible(x, w, precision)
        fused = gemm_nt_bn_stats(x, w, bias, bn, precision) if tc else None    # statistics out of the GEMM epilogue
        if fused is not None:
            z, mean, rstd = fused
        else:
            z = gemm_nt(x, w, bias, precision=precision) if tc else torch.addmm(bias, x, w.t())
            track = bn.track_running_stats and bn.running_mean is not None
            mean, rstd = ops.bn_stats(z, bn.eps, bn.momentum, bn.running_mean if track else None,
                                      bn.running_var if track else None, bn.num_batches_tracked if track else None)
        y = ops.bn_relu_dropout_fwd(z, mean, rstd, gamma, beta, p_drop, seed_dev, layer_id)
        ctx.save_for_backward(x, w, z, mean, rstd, gamma, beta)
        ctx.meta = (p_drop, seed_dev, layer_id, tc, precision)
        ctx.gx_provider = gx_provider
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w, z, mean, rstd, gamma, beta = ctx.saved_tensors
        p_drop, seed_dev, layer_id, tc, precision = ctx.meta
        if gy.stride(1) != 1 or gy.stride(0) % 4 != 0 or gy.data_ptr() % 16 != 0:
            gy = gy.contiguous()
        gz, dgamma, dbeta, dbias = ops.bn_relu_dropout_bwd(gy, z, mean, rstd, gamma, beta, p_drop, seed_dev, layer_id)
        gx = gw = None
        if ctx.needs_input_grad[0]:
            # gx_provider: a buffer the caller wants dL/dx in (the peer-visible gradient matrix of the sharded tables)
            out = ctx.gx_provider() if (tc and ctx.gx_provider is not None) else None
            if out is not None and tuple(out.shape) != tuple(x.shape):
                out = None
            gx = gemm_nt(gz, w.t().contiguous(), out=out, precision=precision) if tc else gz @ w
        if ctx.needs_input_grad[1]:
            gw = gemm_wgrad(gz, x, precision) if tc and wgrad_eligible(gz, x, precision) else gz.t() @ x
        return gx, gw, dbias, dgamma, dbeta, None, None, None, None, None, None


def block_is_fusable(linear, bn, act, drop) -> bool:
    return (isinstance(linear, nn.Linear) and linear.bias is not None and isinstance(bn, nn.BatchNorm1d) and bn.affine
            and bn.momentum is not None and isinstance(act, nn.ReLU) and isinstance(drop, nn.Dropout)
            and linear.out_features % 4 == 0 and 4 <= linear.out_features <= 1024)


def tower_block(x, linear: nn.Linear, bn: nn.BatchNorm1d, drop: nn.Dropout, seed_dev, layer_id: int, weight=None,
                gx_provider=None):
    """Training-mode forward of [linear, bn, ReLU, drop] on a CUDA tensor.  ``weight`` overrides ``linear.weight``
    (the first layer reads a zero-padded copy, see ``CTRModelBase._first_linear``)."""
    w = linear.weight if weight is None else weight
    return _TowerBlockFn.apply(x, w, linear.bias, bn.weight, bn.bias, bn, float(drop.p), seed_dev, layer_id,
                               matmul_precision(), gx_provider)
