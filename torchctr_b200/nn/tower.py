"""One block of the reference tower -- Linear, BatchNorm1d, ReLU, Dropout (``torchctr/models/dnn.py:39-45``) -- as a
single autograd node in training mode (host side of kernels K6 / K6b).

forward   z = x W^T + b (tcgen05 kernel: TF32, or error-compensated 3xTF32 when TF32 matmuls are not allowed)
          -> batch statistics + running-statistics update (``ctr_bn_stats``)
          -> y = dropout(relu(batchnorm(z))) in one pass (``ctr_bn_relu_dropout_fwd``)
backward  one reduction pass + one apply pass give dL/dz, dgamma, dbeta and the bias gradient
          (``ctr_bn_relu_dropout_bwd``); dL/dx on the tcgen05 kernel; dL/dW = gz^T x.
The dropout mask is recomputed from (step seed on the device, layer id, element), never stored.  The modules keep
their parameters, buffers and ``state_dict`` keys; in eval mode they run as plain torch modules.

Off the critical path: nothing downstream of a block's backward needs its WEIGHT gradient before the optimizer step, while
dL/dx feeds the next block and finally the embedding update.  So dL/dW (the split-K tcgen05 kernel + its reduction) is issued on
a second stream right after dL/dx and runs next to the following blocks' BatchNorm backward and the embedding sweep (a parallel
branch when the step is captured into a CUDA graph); the streams are joined when the backward pass ends (an autograd engine
callback), so ``loss.backward(); optimizer.step()`` needs nothing new.  Only when the parameter's ``.grad`` is unset, i.e. when
autograd takes the gradient tensor over without launching a kernel (``zero_grad(set_to_none=True)``, torch's default); with
accumulating gradients everything stays on the one stream.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from .. import ops
from .embedding import BlockedGrad
from .linear import gemm_nt, gemm_nt_bn_stats, gemm_wgrad, matmul_precision, tc_eligible, wgrad_eligible


_wgrad_streams: dict = {}
_deferred: dict = {}       # device -> [graph task id, tensors the second stream still reads (kept alive until the join)]
defer_weight_grads = os.environ.get("CTR_DEFER_WGRAD", "1") != "0"   # module switch (tests compare both schedules)
prepare_early = os.environ.get("CTR_PREPARE_EARLY", "1") != "0"      # first block's weight / seed preparation next to the lookup


def _wgrad_stream(device) -> torch.cuda.Stream:
    key = torch.device(device)
    st = _wgrad_streams.get(key)
    if st is None:
        st = _wgrad_streams[key] = torch.cuda.Stream(device=key, priority=-1 if os.environ.get("CTR_WGRAD_PRIO", "0") == "1" else 0)
    return st


def join_deferred(device=None) -> None:
    """Make the current stream wait for the weight gradients issued on the second stream and let go of the tensors that
    stream was reading.  Runs by itself at the end of every backward pass (on the stream backward() was called on); idempotent."""
    for dev in ([torch.device(device)] if device is not None else list(_deferred)):
        entry = _deferred.pop(dev, None)
        if entry is not None:
            torch.cuda.current_stream(dev).wait_stream(_wgrad_stream(dev))
            entry[1].clear()


def wait_deferred(device) -> None:
    """Make the CURRENT stream wait for the second stream without ending the deferral: for a consumer that reads the dense
    gradients inside the backward pass (the all-reduce of the hybrid placement, issued from the lookup's backward node)."""
    dev = torch.device(device)
    if dev in _deferred:
        torch.cuda.current_stream(dev).wait_stream(_wgrad_stream(dev))


def _defer_begin(dev):
    """-> (main stream, second stream, engine_joins).  Registers the end-of-backward join once per backward pass."""
    main, side = torch.cuda.current_stream(dev), _wgrad_stream(dev)
    engine_joins = True
    task = torch._C._current_graph_task_id()
    if dev in _deferred and _deferred[dev][0] != task:
        join_deferred(dev)                 # left over from a backward pass that did not finish
    if dev not in _deferred:
        _deferred[dev] = [task, []]
        try:        # joined when this backward pass ends, on the stream backward() was called on
            torch.autograd.Variable._execution_engine.queue_callback(lambda: join_deferred(dev))
        except RuntimeError:               # no engine to call back (backward() called by hand)
            engine_joins = False
    return main, side, engine_joins


def _grad_mode(param) -> str:
    """How the gradient of ``param`` can leave the second stream:
    "steal"   ``.grad`` is unset: autograd takes the tensor over as it is (no kernel on its own stream);
    "direct"  ``.grad`` is a view of the model's flat gradient buffer (``CTRModelBase.enable_flat_dense_grads`` marks those
              parameters): the second stream adds into it and autograd gets None;
    "main"    anything else (plain accumulation): autograd adds on its own stream, so the gradient is made there too."""
    if param is None or not param.is_leaf:
        return "main"
    if param.grad is None:
        return "steal"
    return "direct" if getattr(param, "_ctr_direct_grad", False) else "main"


def _takes_over(param) -> bool:
    return _grad_mode(param) != "main"


def _deliver(param, g):
    """(on the second stream) hand ``g`` to autograd, or add it into the flat buffer and hand None"""
    if g is not None and _grad_mode(param) == "direct":
        param.grad.add_(g.view_as(param.grad))
        return None
    return g


class _TowerBlockFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, gamma, beta, bn, p_drop, seed_dev, layer_id, precision, gx_provider=None, pad=0,
                prepared=None):
        # pad: zero columns appended to the weight here (the first layer reads the 4-float-padded lookup output); the gradient
        # handed back is that of the unpadded parameter
        # prepared: (padded weight, its transpose or None, event) made by ``prepare_first_block`` on the second stream BEFORE the
        # lookup was issued -- the small copies then run next to the lookup instead of between it and the first GEMM
        ctx.param = weight
        ctx.bias_param = bias
        ctx.bn_params = (gamma, beta)
        ctx.pad = pad
        ctx.wt = ctx.wt_ready = None
        if prepared is not None:
            w, wt, ready = prepared
            torch.cuda.current_stream(x.device).wait_event(ready)
            tc = tc_eligible(x, w, precision)
            if tc and ctx.needs_input_grad[0] and wt is not None:
                ctx.wt, ctx.wt_ready = wt, ready
        else:
            w = torch.nn.functional.pad(weight.detach(), (0, pad)) if pad else weight.contiguous()
            tc = tc_eligible(x, w, precision)
        if prepared is None and tc and defer_weight_grads and ctx.needs_input_grad[0] and precision == "tf32":
            # W^T for the input-gradient GEMM of backward, made NOW on the second stream (next to this block's forward)
            # instead of on the critical path of backward
            dev = x.device
            main, side = torch.cuda.current_stream(dev), _wgrad_stream(dev)
            wt = torch.empty(w.shape[1], w.shape[0], dtype=w.dtype, device=dev)
            side.wait_stream(main)
            with torch.cuda.stream(side):
                wt.copy_(w.t())
                ctx.wt_ready = torch.cuda.Event()
                ctx.wt_ready.record(side)
            ctx.wt = wt
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
        dev = gy.device
        want_dbias = ctx.needs_input_grad[2]
        defer_dbias = defer_weight_grads and want_dbias and _takes_over(ctx.bias_param)
        gz, dgamma, dbeta, dbias = ops.bn_relu_dropout_bwd(gy, z, mean, rstd, gamma, beta, p_drop, seed_dev, layer_id,
                                                           want_dbias=want_dbias,
                                                           defer_dbias_tag=("block", layer_id) if defer_dbias else None)
        gx = gw = None
        if ctx.needs_input_grad[0]:
            # gx_provider: a buffer the caller wants dL/dx in (the peer-visible gradient matrix of the sharded tables)
            out = ctx.gx_provider() if (tc and ctx.gx_provider is not None) else None
            blocked = out if isinstance(out, BlockedGrad) else None
            if blocked is not None:
                out = None
                if (blocked.B, blocked.stride) != tuple(x.shape) or blocked.cols > x.shape[1] or blocked.cols % blocked.block:
                    blocked = None
            if out is not None and tuple(out.shape) != tuple(x.shape):
                out = None
            if tc:
                wt = ctx.wt
                if wt is not None:
                    torch.cuda.current_stream(dev).wait_event(ctx.wt_ready)
                else:
                    wt = w.t().contiguous()
                if blocked is not None:
                    # the embedding update reads dL/dx feature by feature: write it column-blocked (and only the table columns:
                    # the dense block and the padding behind them need no gradient)
                    gemm_nt(gz, wt[:blocked.cols], out=blocked.buffer, precision=precision, out_block=blocked.block)
                    blocked.marked = True
                    gx = blocked.buffer.view(blocked.B, blocked.stride)
                else:
                    gx = gemm_nt(gz, wt, out=out, precision=precision)
            else:
                gx = gz @ w
        want_gw = ctx.needs_input_grad[1]
        defer_gw = defer_weight_grads and want_gw and tc and _takes_over(ctx.param)

        def weight_grad():
            g = gemm_wgrad(gz, x, precision) if tc and wgrad_eligible(gz, x, precision) else gz.t() @ x
            return g[:, :g.shape[1] - ctx.pad].contiguous() if ctx.pad else g

        if want_gw and not defer_gw:
            gw = weight_grad()
        if defer_gw or defer_dbias:
            main, side, engine_joins = _defer_begin(dev)
            _deferred[dev][1] += [gz, x]           # main-stream blocks the second stream reads: alive until the join
            direct_bn = [defer_weight_grads and _grad_mode(p) == "direct" for p in ctx.bn_params]
            if any(direct_bn):
                _deferred[dev][1] += [dgamma, dbeta]
            side.wait_stream(main)                 # after dL/dx was issued: the critical path goes first
            with torch.cuda.stream(side):
                if defer_gw:
                    gw = _deliver(ctx.param, weight_grad())
                if defer_dbias:
                    dbias = _deliver(ctx.bias_param, ops.bn_bias_grad_deferred(gz, ("block", layer_id)))
                if direct_bn[0]:
                    dgamma = _deliver(ctx.bn_params[0], dgamma)
                if direct_bn[1]:
                    dbeta = _deliver(ctx.bn_params[1], dbeta)
            if not engine_joins:
                join_deferred(dev)
        return gx, gw, dbias, dgamma, dbeta, None, None, None, None, None, None, None, None


def block_is_fusable(linear, bn, act, drop) -> bool:
    return (isinstance(linear, nn.Linear) and linear.bias is not None and isinstance(bn, nn.BatchNorm1d) and bn.affine
            and bn.momentum is not None and isinstance(act, nn.ReLU) and isinstance(drop, nn.Dropout)
            and linear.out_features % 4 == 0 and 4 <= linear.out_features <= 1024)


def prepare_first_block(linear: nn.Linear, pad: int, seed_counter: torch.Tensor, want_wt: bool = True):
    """What the first block's forward needs besides its input, made on the SECOND stream before the lookup is issued: the step's
    dropout seed (counter advanced, snapshot taken), the zero-padded weight and (TF32 mode) its transpose for the input-gradient
    GEMM.  In a step captured from one stream these five tiny kernels otherwise sit, one after the other, between the lookup
    and the first GEMM (~20 us of a 0.8 ms step); forked off before the lookup they run next to it.
    -> (seed snapshot, padded weight [N, K + pad], transpose [K + pad, N] | None, event recorded on the second stream)."""
    w0 = linear.weight.detach()
    dev = w0.device
    N, K = w0.shape
    main, side = torch.cuda.current_stream(dev), _wgrad_stream(dev)
    snap = torch.empty_like(seed_counter)            # allocated on the stream that will read them, written on the second one
    w = torch.empty(N, K + pad, dtype=w0.dtype, device=dev) if pad else None
    wt = torch.empty(K + pad, N, dtype=w0.dtype, device=dev) if want_wt else None
    side.wait_stream(main)
    with torch.cuda.stream(side):
        seed_counter += 1
        snap.copy_(seed_counter)
        if pad:
            w[:, :K].copy_(w0)
            w[:, K:].zero_()
        else:
            w = w0.contiguous()
        if wt is not None:
            wt[:K].copy_(w0.t())
            if pad:
                wt[K:].zero_()
        ready = torch.cuda.Event()
        ready.record(side)
    return snap, w, wt, ready


def tower_block(x, linear: nn.Linear, bn: nn.BatchNorm1d, drop: nn.Dropout, seed_dev, layer_id: int, weight=None,
                gx_provider=None, prepared=None):
    """Training-mode forward of [linear, bn, ReLU, drop] on a CUDA tensor.  When ``x`` is wider than ``linear.in_features`` (the
    first layer reads the 4-float-padded lookup output) the weight is zero-padded inside the node.  ``weight`` overrides
    ``linear.weight`` as is.  ``prepared``: (padded weight, transpose | None, event) from ``prepare_first_block``."""
    if weight is not None:
        return _TowerBlockFn.apply(x, weight, linear.bias, bn.weight, bn.bias, bn, float(drop.p), seed_dev, layer_id,
                                   matmul_precision(), gx_provider, 0, None)
    pad = x.shape[1] - linear.in_features
    return _TowerBlockFn.apply(x, linear.weight, linear.bias, bn.weight, bn.bias, bn, float(drop.p), seed_dev, layer_id,
                               matmul_precision(), gx_provider, pad, prepared)
