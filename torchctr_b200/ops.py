"""Tensor-level wrappers of the C ABI (``include/ctr_b200.h``).

Each function takes CUDA tensors, checks dtypes / contiguity, and calls the matching
``ctr_*`` entry point on the current torch CUDA stream.  Nothing here computes on the host.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import torch

from . import _lib
from ._lib import (INDEX_DIRECT, INDEX_HASH, INDEX_REMAP, INDEX_WINDOW, OPT_ADAGRAD, OPT_ADAM, OPT_GRAD_OUT, OPT_NONE,  # noqa: F401
                   OPT_ROWWISE_ADAGRAD, OPT_SGD, POOL_MEAN, POOL_SUM)

_INDEX_KINDS = {"direct": INDEX_DIRECT, "hash": INDEX_HASH, "vocab": INDEX_REMAP, "remap": INDEX_REMAP,
                "window": INDEX_WINDOW}      # window: rows [hash_seed, hash_seed + num_rows) of a table, other ids are padding
_POOLINGS = {"sum": POOL_SUM, "mean": POOL_MEAN}
_OPT_KINDS = {"none": OPT_NONE, "sgd": OPT_SGD, "adagrad": OPT_ADAGRAD, "rowwise_adagrad": OPT_ROWWISE_ADAGRAD,
              "adam": OPT_ADAM, "grad_out": OPT_GRAD_OUT}


def _chk(t, name, dtype, cuda=True):
    if t is None:
        return
    if cuda:
        _lib.require_cuda(t, name)
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")


class VocabMapHandle:
    """Device hash map raw key (int64) -> row (int32); owns nothing but references."""

    def __init__(self, keys: torch.Tensor, rows: torch.Tensor):
        _chk(keys, "map keys", torch.int64)
        _chk(rows, "map rows", torch.int32)
        cap = keys.numel()
        if cap == 0 or cap & (cap - 1) or rows.numel() != cap:
            raise ValueError("vocabulary map capacity must be a power of two and match rows")
        self.keys, self.rows = keys, rows
        self.struct = _lib.VocabMap(keys.data_ptr(), rows.data_ptr(), cap)


@dataclass
class FeatureSpec:
    """One table of a launch group (mirrors ``ctr_feature_t``)."""
    ids: torch.Tensor                       # i64 [B, L]
    table: torch.Tensor | None              # f32 [V, D]
    num_rows: int
    D: int
    out_col: int
    pooling: str = "sum"
    index_kind: str = "direct"
    hash_seed: int = 0
    id_weight: torch.Tensor | None = None   # f32 [B, L]
    vocab: VocabMapHandle | None = None
    state0: torch.Tensor | None = None
    state1: torch.Tensor | None = None
    bag_scale: torch.Tensor | None = None   # f32 [B]
    twin_table: torch.Tensor | None = None  # f32 [V] / [V, 1]: one-column twin indexed by the same ids (DeepFM first order)
    twin_state0: torch.Tensor | None = None
    twin_state1: torch.Tensor | None = None


@dataclass
class GroupCall:
    """A lowered ``ctr_group_t`` plus the Python objects that keep its pointers alive."""
    struct: _lib.Group
    keep: list = field(default_factory=list)
    device: torch.device | None = None      # where the group's tensors live: launches run under this device and on ITS stream


def make_group(features, B: int, out: torch.Tensor | None, out_stride: int, dense: torch.Tensor | None = None,
               dense_col: int = 0, zero_from: int = -1, status: torch.Tensor | None = None,
               extra: torch.Tensor | None = None, fm_sum: torch.Tensor | None = None, fm: bool = False,
               grad_blocked: bool = False) -> GroupCall:
    """``extra`` f32 [B] (+ ``fm_sum`` f32 [B, D] when ``fm``): the fused per-bag scalar of ``ctr_group_t`` -- sum of the
    features' twin tables plus, with ``fm``, the FM second-order term; forward writes them, backward reads ``extra`` as
    dL/d extra.  ``grad_blocked`` (backward): ``out`` holds the gradient column-blocked, feature by feature
    (``ctr_group_t.grad_blocked``, written by ``linear_fwd(..., out_block=D)``)."""
    n = len(features)
    if n > _lib.MAX_FEATURES:
        raise ValueError(f"a launch group holds at most {_lib.MAX_FEATURES} features, got {n}")
    arr = (_lib.Feature * max(n, 1))()
    keep = [arr, out, dense, status, extra, fm_sum]
    for i, f in enumerate(features):
        _chk(f.ids, f"feature {i} ids", torch.int64)
        if f.ids.dim() != 2 or f.ids.shape[0] != B:
            raise ValueError(f"feature {i}: ids must be [B={B}, L], got {tuple(f.ids.shape)}")
        _chk(f.table, f"feature {i} table", torch.float32)
        _chk(f.id_weight, f"feature {i} id_weight", torch.float32)
        _chk(f.state0, f"feature {i} state0", torch.float32)
        _chk(f.state1, f"feature {i} state1", torch.float32)
        _chk(f.bag_scale, f"feature {i} bag_scale", torch.float32)
        if f.id_weight is not None and f.id_weight.shape != f.ids.shape:
            raise ValueError(f"feature {i}: id_weight shape {tuple(f.id_weight.shape)} != ids {tuple(f.ids.shape)}")
        if f.table is not None and tuple(f.table.shape) != (f.num_rows, f.D):
            raise ValueError(f"feature {i}: table shape {tuple(f.table.shape)} != ({f.num_rows}, {f.D})")
        s = arr[i]
        s.ids = f.ids.data_ptr()
        s.id_weight = _lib.ptr(f.id_weight)
        s.table = _lib.ptr(f.table)
        s.state0 = _lib.ptr(f.state0)
        s.state1 = _lib.ptr(f.state1)
        s.bag_scale = _lib.ptr(f.bag_scale)
        s.map = C.pointer(f.vocab.struct) if f.vocab is not None else None
        s.num_rows = f.num_rows
        s.L = f.ids.shape[1]
        s.D = f.D
        s.out_col = f.out_col
        s.pooling = _POOLINGS[f.pooling]
        s.index_kind = _INDEX_KINDS[f.index_kind]
        s.hash_seed = f.hash_seed & 0xFFFFFFFF
        for name in ("twin_table", "twin_state0", "twin_state1"):
            t = getattr(f, name)
            _chk(t, f"feature {i} {name}", torch.float32)
            if t is not None and t.numel() != f.num_rows:
                raise ValueError(f"feature {i}: {name} must hold num_rows={f.num_rows} floats, got {tuple(t.shape)}")
            setattr(s, name, _lib.ptr(t))
        keep.append((f.twin_table, f.twin_state0, f.twin_state1))
        keep.append((f.ids, f.id_weight, f.table, f.state0, f.state1, f.bag_scale, f.vocab))
    _chk(out, "out", torch.float32)
    _chk(dense, "dense", torch.float32)
    if status is not None:
        _chk(status, "status", torch.int32)
    g = _lib.Group()
    g.features = arr
    g.num_features = n
    g.B = B
    g.out = _lib.ptr(out)
    g.out_stride = out_stride
    g.dense = _lib.ptr(dense)
    g.dense_width = 0 if dense is None else dense.shape[1]
    g.dense_col = dense_col
    g.zero_from = zero_from
    g.status = _lib.ptr(status)
    g.grad_blocked = 1 if grad_blocked else 0
    _chk(extra, "extra", torch.float32)
    _chk(fm_sum, "fm_sum", torch.float32)
    if extra is not None and extra.numel() != B:
        raise ValueError(f"extra must hold B={B} floats")
    if fm and (extra is None or fm_sum is None or fm_sum.shape[0] != B):
        raise ValueError("fm needs extra [B] and fm_sum [B, D]")
    g.extra = _lib.ptr(extra)
    g.fm_sum = _lib.ptr(fm_sum)
    g.fm = 1 if fm else 0
    dev = None
    for t in [out, dense] + [f.ids for f in features] + [f.table for f in features]:
        if t is not None and t.is_cuda:
            dev = t.device
            break
    return GroupCall(g, keep, dev)


def _stream(t=None):
    """Raw handle of the current torch stream of the device ``t`` (a tensor, a GroupCall or None = current device) lives on."""
    dev = getattr(t, "device", None)
    return _lib.stream_ptr(dev)


def _device_of(args, kwargs):
    for a in list(args) + list(kwargs.values()):
        if isinstance(a, (list, tuple)) and a and isinstance(a[0], torch.Tensor):
            a = a[0]
        if isinstance(a, GroupCall) and a.device is not None:
            return a.device
        if isinstance(a, torch.Tensor) and a.is_cuda:
            return a.device
        if isinstance(a, VocabMapHandle):
            return a.keys.device
    return None


def _guarded(fn):
    """Runs ``fn`` with the CUDA device of its first tensor / group argument current: a kernel launched from another
    device's context onto this device's stream (or pointers) is an error or, worse, an unordered access."""
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        dev = _device_of(args, kwargs)
        if dev is None or dev.index is None or dev.index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)
    return wrapper


class KernelTimer:
    """Optional CUDA-event timing of the C-ABI calls on the launching stream (bench.py's roofline
    numbers).  Off by default; ``with KernelTimer() as kt: ...; kt.summary()``."""
    active = None

    def __init__(self):
        self.spans = {}

    def __enter__(self):
        KernelTimer.active = self
        return self

    def __exit__(self, *exc):
        KernelTimer.active = None

    def summary(self):
        """name -> (calls, total ms); synchronises."""
        torch.cuda.synchronize()
        return {k: (len(v), sum(a.elapsed_time(b) for a, b in v)) for k, v in self.spans.items()}


class _timed:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        kt = KernelTimer.active
        if kt is not None:
            self.start = torch.cuda.Event(enable_timing=True)
            self.start.record()

    def __exit__(self, *exc):
        kt = KernelTimer.active
        if kt is not None:
            end = torch.cuda.Event(enable_timing=True)
            end.record()
            kt.spans.setdefault(self.name, []).append((self.start, end))


def kernel_launches() -> int:
    return int(_lib.lib().ctr_kernel_launches())


def _width_tag(call: GroupCall) -> str:
    """Timer label suffix: the embedding dim of the group's first table (DeepFM runs a D=16 and a D=1 group)."""
    return f"_d{call.struct.features[0].D}" if call.struct.num_features > 0 else ""


@_guarded
def emb_pool_fwd(call: GroupCall) -> None:
    with _timed("emb_pool_fwd" + _width_tag(call)):
        _lib.check(_lib.lib().ctr_emb_pool_fwd(C.byref(call.struct), _stream(call)), "ctr_emb_pool_fwd")


@_guarded
def emb_bwd_workspace_bytes(call: GroupCall) -> int:
    return _lib.check(_lib.lib().ctr_emb_bwd_workspace_bytes(C.byref(call.struct)), "ctr_emb_bwd_workspace_bytes")


@_guarded
def emb_bwd_plan(call: GroupCall, workspace: torch.Tensor, runs: bool = True) -> None:
    """Sort the (row, slot) pairs of the group; ``runs=False`` skips the run list, which only the unique-row
    outputs of ``emb_bwd_apply`` need (a fused training step does not ask for them)."""
    _chk(workspace, "workspace", torch.uint8)
    with _timed("emb_bwd_plan"):
        _lib.check(_lib.lib().ctr_emb_bwd_plan_ex(C.byref(call.struct), workspace.data_ptr(), workspace.numel(),
                                                  0 if runs else _lib.PLAN_NO_RUNS, _stream(call)), "ctr_emb_bwd_plan")


def make_opt(kind: str = "none", lr: float = 0.0, eps: float = 1e-10, betas=(0.9, 0.999), step: int = 1,
             device_hyper: torch.Tensor | None = None) -> _lib.Opt:
    """``device_hyper``: optional f32 [5] CUDA tensor the kernels read the scalars from (see ``opt_hyper``)."""
    _chk(device_hyper, "device_hyper", torch.float32)
    return _lib.Opt(_OPT_KINDS[kind], step, lr, eps, betas[0], betas[1], _lib.ptr(device_hyper))


def opt_hyper(opt: _lib.Opt):
    """The five fp32 scalars (lr, eps, 1-beta1, 1-beta2, adam step size) the kernels use for ``opt``."""
    h = _lib.Hyper()
    _lib.lib().ctr_opt_hyper(C.byref(opt), C.byref(h))
    return [h.lr, h.eps, h.one_minus_beta1, h.one_minus_beta2, h.adam_step_size]


@_guarded
def emb_bwd_apply(call: GroupCall, workspace: torch.Tensor, opt: _lib.Opt, uniq_feature=None, uniq_row=None,
                  row_grad=None, num_unique=None) -> None:
    _chk(uniq_feature, "uniq_feature", torch.int32)
    _chk(uniq_row, "uniq_row", torch.int32)
    _chk(row_grad, "row_grad", torch.float32)
    _chk(num_unique, "num_unique", torch.int64)
    stride = 0 if row_grad is None else row_grad.shape[1]
    with _timed("emb_bwd_apply" + _width_tag(call)):
        _lib.check(_lib.lib().ctr_emb_bwd_apply(C.byref(call.struct), workspace.data_ptr(), C.byref(opt),
                                                _lib.ptr(uniq_feature), _lib.ptr(uniq_row), _lib.ptr(row_grad), stride,
                                                _lib.ptr(num_unique), _stream(call)), "ctr_emb_bwd_apply")


@_guarded
def fm_pack_grads(gx: torch.Tensor, extra_grad, fm_sum, cols, D: int, packed: torch.Tensor) -> None:
    """packed[b, j*D:(j+1)*D] = gx[b, cols[j]:cols[j]+D] + extra_grad[b] * fm_sum[b, :] (hybrid placement, requester side)."""
    _rows2d(gx, "gx")
    _chk(extra_grad, "extra_grad", torch.float32)
    _chk(fm_sum, "fm_sum", torch.float32)
    _chk(packed, "packed", torch.float32)
    B, J = gx.shape[0], len(cols)
    if packed.numel() < B * J * D:
        raise ValueError("packed is too small")
    arr = (C.c_int32 * J)(*[int(c) for c in cols])
    with _timed("fm_pack_grads"):
        _lib.check(_lib.lib().ctr_fm_pack_grads(gx.data_ptr(), gx.stride(0), _lib.ptr(extra_grad), _lib.ptr(fm_sum), B, D, arr, J,
                                                packed.data_ptr(), _stream(gx)), "ctr_fm_pack_grads")


@_guarded
def rows_dense_apply(params: torch.Tensor, grads: torch.Tensor, state0: torch.Tensor | None, opt: _lib.Opt, clear: bool = True) -> None:
    """Dense sgd / adagrad update of a replicated table block from its (all-reduced) gradient buffer; zero gradients are
    skipped and, with ``clear``, the buffer is zero again afterwards."""
    _chk(params, "params", torch.float32)
    _chk(grads, "grads", torch.float32)
    _chk(state0, "state0", torch.float32)
    if grads.numel() != params.numel() or (state0 is not None and state0.numel() != params.numel()):
        raise ValueError("params / grads / state0 must hold the same number of elements")
    with _timed("rows_dense_apply"):
        _lib.check(_lib.lib().ctr_rows_dense_apply(C.byref(opt), params.data_ptr(), grads.data_ptr(), _lib.ptr(state0),
                                                   params.numel(), 1 if clear else 0, _stream(params)), "ctr_rows_dense_apply")


@_guarded
def hash_bucket(ids: torch.Tensor, buckets: int, seed: int = 0) -> torch.Tensor:
    """``torchctr.utils.hash_bucket`` (utils.py:103-119) on a tensor of integer ids -> int32 buckets."""
    _chk(ids, "ids", torch.int64)
    if not 0 <= seed <= 0xFFFFFFFF:
        raise OverflowError("seed must fit uint32")      # sklearn raises OverflowError too
    if not 0 < buckets < 2 ** 31:
        raise ValueError("buckets must be in [1, 2^31)")
    out = torch.empty(ids.shape, dtype=torch.int32, device=ids.device)
    _lib.check(_lib.lib().ctr_hash_bucket_i64(ids.data_ptr(), ids.numel(), buckets, seed, out.data_ptr(), _stream(ids)),
               "ctr_hash_bucket_i64")
    return out


@_guarded
def hash_bucket_bytes(data: torch.Tensor, offsets: torch.Tensor, buckets: int, seed: int = 0) -> torch.Tensor:
    """``hash_bucket`` of n byte strings packed back to back: data u8 [total], offsets i64 [n + 1] -> int32 [n]."""
    _chk(data, "data", torch.uint8)
    _chk(offsets, "offsets", torch.int64)
    if not 0 <= seed <= 0xFFFFFFFF:
        raise OverflowError("seed must fit uint32")
    if not 0 < buckets < 2 ** 31:
        raise ValueError("buckets must be in [1, 2^31)")
    n = offsets.numel() - 1
    out = torch.empty(max(n, 0), dtype=torch.int32, device=offsets.device)
    _lib.check(_lib.lib().ctr_hash_bucket_bytes(data.data_ptr(), offsets.data_ptr(), n, buckets, seed, out.data_ptr(), _stream(offsets)),
               "ctr_hash_bucket_bytes")
    return out


@_guarded
def rows_gather(ids: torch.Tensor, table: torch.Tensor, status: torch.Tensor | None = None) -> torch.Tensor:
    _chk(ids, "ids", torch.int64)
    _chk(table, "table", torch.float32)
    out = torch.empty(*ids.shape, table.shape[1], dtype=torch.float32, device=table.device)
    _lib.check(_lib.lib().ctr_rows_gather(ids.data_ptr(), ids.numel(), table.data_ptr(), table.shape[0], table.shape[1],
                                          out.data_ptr(), _lib.ptr(status), _stream(table)), "ctr_rows_gather")
    return out


@_guarded
def normal_fill_rows(table: torch.Tensor, row0: int, n: int, mean: float, std: float, seed: int) -> None:
    _chk(table, "table", torch.float32)
    if row0 < 0 or row0 + n > table.shape[0]:
        raise ValueError("row range outside the table")
    _lib.check(_lib.lib().ctr_normal_fill_rows(table.data_ptr(), row0, n, table.shape[1], mean, std,
                                               seed & 0xFFFFFFFFFFFFFFFF, _stream(table)), "ctr_normal_fill_rows")


@_guarded
def normal_fill_rows_strided(table: torch.Tensor, row0: int, n: int, mean: float, std: float, seed: int, global_first: int,
                             global_stride: int) -> None:
    """Rows [row0, row0 + n) of a shard := the values ``normal_fill_rows`` gives global rows first, first + stride, ..."""
    _chk(table, "table", torch.float32)
    if row0 < 0 or row0 + n > table.shape[0]:
        raise ValueError("row range outside the table")
    _lib.check(_lib.lib().ctr_normal_fill_rows_strided(table.data_ptr(), row0, n, table.shape[1], mean, std,
                                                       seed & 0xFFFFFFFFFFFFFFFF, global_first, global_stride, _stream(table)),
               "ctr_normal_fill_rows_strided")


@_guarded
def ids_minmax(ids: torch.Tensor) -> torch.Tensor:
    """Device tensor i64 [2] = (min, max) of ids."""
    _chk(ids, "ids", torch.int64)
    out = torch.empty(2, dtype=torch.int64, device=ids.device)
    _lib.check(_lib.lib().ctr_ids_minmax(ids.data_ptr(), ids.numel(), out.data_ptr(), _stream(ids)), "ctr_ids_minmax")
    return out


@_guarded
def vocab_fit(vmap: VocabMapHandle, keys: torch.Tensor, next_row: torch.Tensor, min_freq: int = 0,
              counts: torch.Tensor | None = None, status: torch.Tensor | None = None) -> None:
    _chk(keys, "keys", torch.int64)
    _chk(next_row, "next_row", torch.int64)
    _chk(counts, "counts", torch.int64)
    n = keys.numel()
    nbytes = _lib.check(_lib.lib().ctr_vocab_fit_workspace_bytes(n))
    ws = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=keys.device)
    _lib.check(_lib.lib().ctr_vocab_fit(C.byref(vmap.struct), keys.data_ptr(), n, int(min_freq or 0), next_row.data_ptr(),
                                        _lib.ptr(counts), _lib.ptr(status), ws.data_ptr(), ws.numel(), _stream(keys)),
               "ctr_vocab_fit")


@_guarded
def vocab_transform(vmap: VocabMapHandle, keys: torch.Tensor, oov_row: int = 0) -> torch.Tensor:
    _chk(keys, "keys", torch.int64)
    rows = torch.empty(keys.shape, dtype=torch.int32, device=keys.device)
    _lib.check(_lib.lib().ctr_vocab_transform(C.byref(vmap.struct), keys.data_ptr(), keys.numel(), oov_row,
                                              rows.data_ptr(), _stream(keys)), "ctr_vocab_transform")
    return rows


@_guarded
def vocab_insert(vmap: VocabMapHandle, keys: torch.Tensor, rows: torch.Tensor, status: torch.Tensor | None = None) -> None:
    _chk(keys, "keys", torch.int64)
    _chk(rows, "rows", torch.int32)
    _lib.check(_lib.lib().ctr_vocab_insert(C.byref(vmap.struct), keys.data_ptr(), rows.data_ptr(), keys.numel(),
                                           _lib.ptr(status), _stream(keys)), "ctr_vocab_insert")


@_guarded
def vocab_clear(vmap: VocabMapHandle) -> None:
    _lib.check(_lib.lib().ctr_vocab_clear(C.byref(vmap.struct), _stream(vmap.keys)), "ctr_vocab_clear")


@_guarded
def fm_fwd(x: torch.Tensor, F: int, D: int, first: torch.Tensor | None = None, out: torch.Tensor | None = None,
           accumulate: bool = False) -> torch.Tensor:
    """x f32 [B, stride>=F*D] -> out f32 [B, 1] = FM second-order term (+ row sums of ``first`` [B, n])."""
    _lib.require_cuda(x, "x")
    if x.dtype != torch.float32 or x.dim() != 2 or x.stride(1) != 1:
        raise ValueError("x must be f32 [B, *] with unit inner stride")
    B = x.shape[0]
    if out is None:
        out = torch.empty(B, 1, dtype=torch.float32, device=x.device)
    nfirst = 0 if first is None else first.shape[1]
    if first is not None and (first.dtype != torch.float32 or first.stride(1) != 1):
        raise ValueError("first must be f32 with unit inner stride")
    with _timed("fm_fwd"):
        _lib.check(_lib.lib().ctr_fm_fwd(x.data_ptr(), x.stride(0), B, F, D, _lib.ptr(first), nfirst,
                                         0 if first is None else first.stride(0), out.data_ptr(), out.stride(0),
                                         int(accumulate), _stream(x)), "ctr_fm_fwd")
    return out


@_guarded
def fm_bwd(x: torch.Tensor, F: int, D: int, gout: torch.Tensor, gx: torch.Tensor, accumulate: bool,
           gfirst: torch.Tensor | None = None) -> None:
    B = x.shape[0]
    nfirst = 0 if gfirst is None else gfirst.shape[1]
    with _timed("fm_bwd"):
        _lib.check(_lib.lib().ctr_fm_bwd(x.data_ptr(), x.stride(0), B, F, D, gout.data_ptr(), gout.stride(0), gx.data_ptr(),
                                         gx.stride(0), int(accumulate), _lib.ptr(gfirst), nfirst,
                                         0 if gfirst is None else gfirst.stride(0), _stream(x)), "ctr_fm_bwd")


@_guarded
def target_attention_fwd(target, cand, mask, honor_mask: bool):
    """-> (out [B, E], scores [B, N], row_max [B], row_sum [B]); see ``ctr_target_attention_fwd``."""
    _chk(target, "target", torch.float32)
    _chk(cand, "cand", torch.float32)
    _chk(mask, "mask", torch.float32)
    B, N, E = cand.shape
    dev = cand.device
    out = torch.empty(B, E, dtype=torch.float32, device=dev)
    scores = torch.empty(B, N, dtype=torch.float32, device=dev)
    row_max = torch.empty(B, dtype=torch.float32, device=dev)
    row_sum = torch.empty(B, dtype=torch.float32, device=dev)
    with _timed("target_attention_fwd"):
        _lib.check(_lib.lib().ctr_target_attention_fwd(target.data_ptr(), cand.data_ptr(), _lib.ptr(mask), B, N, E, int(honor_mask),
                                                       out.data_ptr(), scores.data_ptr(), row_max.data_ptr(), row_sum.data_ptr(),
                                                       _stream(cand)), "ctr_target_attention_fwd")
    return out, scores, row_max, row_sum


@_guarded
def target_attention_bwd(target, cand, mask, honor_mask: bool, out, scores, row_max, row_sum, gout):
    B, N, E = cand.shape
    gtarget = torch.empty_like(target)
    gcand = torch.empty_like(cand)
    with _timed("target_attention_bwd"):
        _lib.check(_lib.lib().ctr_target_attention_bwd(target.data_ptr(), cand.data_ptr(), _lib.ptr(mask), B, N, E, int(honor_mask),
                                                       out.data_ptr(), scores.data_ptr(), row_max.data_ptr(), row_sum.data_ptr(),
                                                       gout.data_ptr(), gtarget.data_ptr(), gcand.data_ptr(), _stream(cand)),
                   "ctr_target_attention_bwd")
    return gtarget, gcand


@_guarded
def cross_combine_fwd(x0, x, u, bias) -> torch.Tensor:
    for name, t in (("x0", x0), ("x", x), ("u", u)):
        _chk(t, name, torch.float32)
    _chk(bias, "bias", torch.float32)
    B, d = x0.shape
    y = torch.empty_like(x0)
    _lib.check(_lib.lib().ctr_cross_combine_fwd(x0.data_ptr(), x.data_ptr(), u.data_ptr(), bias.data_ptr(), B, d, d,
                                                y.data_ptr(), _stream(x0)), "ctr_cross_combine_fwd")
    return y


@_guarded
def cross_combine_bwd(x0, u, bias, gy, gx0: torch.Tensor, accumulate: bool) -> torch.Tensor:
    for name, t in (("x0", x0), ("u", u), ("gy", gy), ("gx0", gx0)):
        _chk(t, name, torch.float32)
    B, d = x0.shape
    gu = torch.empty_like(x0)
    _lib.check(_lib.lib().ctr_cross_combine_bwd(x0.data_ptr(), u.data_ptr(), bias.data_ptr(), gy.data_ptr(), B, d, d,
                                                gu.data_ptr(), gx0.data_ptr(), int(accumulate), _stream(x0)),
               "ctr_cross_combine_bwd")
    return gu


@_guarded
def route_build(call: GroupCall, world: int, base: torch.Tensor, S: int):
    """Buckets the id slots of ``call`` by owner rank.  Returns (counts i64 [world+1], send_rows i64 [S],
    inv i64 [S], workspace) -- see ``ctr_route_build``."""
    _chk(base, "base", torch.int64)
    dev = base.device
    counts = torch.empty(world + 1, dtype=torch.int64, device=dev)
    send_rows = torch.empty(S, dtype=torch.int64, device=dev)
    inv = torch.empty(S, dtype=torch.int64, device=dev)
    nbytes = _lib.check(_lib.lib().ctr_route_workspace_bytes(C.byref(call.struct), world))
    ws = torch.empty(nbytes + 256, dtype=torch.uint8, device=dev)
    with _timed("route_build"):
        _lib.check(_lib.lib().ctr_route_build(C.byref(call.struct), world, base.data_ptr(), counts.data_ptr(),
                                              send_rows.data_ptr(), inv.data_ptr(), ws.data_ptr(), ws.numel(),
                                              _stream(base)), "ctr_route_build")
    return counts, send_rows, inv, ws


@_guarded
def route_grad_gather(call: GroupCall, world: int, workspace: torch.Tensor, n: int, D: int) -> torch.Tensor:
    g_send = torch.empty(n, D, dtype=torch.float32, device=workspace.device)
    with _timed("route_grad_gather"):
        _lib.check(_lib.lib().ctr_route_grad_gather(C.byref(call.struct), world, workspace.data_ptr(), n, D,
                                                    g_send.data_ptr(), _stream(workspace)), "ctr_route_grad_gather")
    return g_send


@_guarded
def linear_fwd(A: torch.Tensor, W: torch.Tensor, bias: torch.Tensor | None = None, act: int = 0,
               out: torch.Tensor | None = None, out_block: int = 0) -> torch.Tensor:
    """``act(A @ W.T + bias)`` on tcgen05 tensor cores (TF32 inputs, fp32 accumulate).  A f32 [M, K] with unit
    inner stride and a row pitch that is a multiple of 4 floats, W f32 [N, K] likewise.  ``out_block = D > 0``: the result is
    stored column-blocked into the flat buffer ``out`` (>= M * N floats): columns [j D, (j + 1) D) are the contiguous matrix
    ``out[j * M * D:]`` viewed [M, D] (``ctr_linear_fwd_blocked``; needs M > 128)."""
    _lib.require_cuda(A, "A")
    for name, t in (("A", A), ("W", W)):
        if t.dtype != torch.float32 or t.dim() != 2 or t.stride(1) != 1:
            raise ValueError(f"{name} must be f32 [rows, K] with unit inner stride")
    M, K = A.shape
    N = W.shape[0]
    if W.shape[1] != K:
        raise ValueError(f"shape mismatch: A {tuple(A.shape)} vs W {tuple(W.shape)}")
    _chk(bias, "bias", torch.float32)
    if out_block:
        if out is None or not out.is_contiguous() or out.numel() < M * N or out.dtype != torch.float32:
            raise ValueError("out_block needs a contiguous f32 `out` of at least M * N elements")
        with _timed("linear_fwd"):
            _lib.check(_lib.lib().ctr_linear_fwd_blocked(A.data_ptr(), A.stride(0), W.data_ptr(), W.stride(0), _lib.ptr(bias),
                                                         out.data_ptr(), out_block, M * out_block, M, N, K, act, _stream(A)),
                       "ctr_linear_fwd_blocked")
        return out
    if out is None:
        out = torch.empty(M, N, dtype=torch.float32, device=A.device)
    with _timed("linear_fwd"):
        _lib.check(_lib.lib().ctr_linear_fwd(A.data_ptr(), A.stride(0), W.data_ptr(), W.stride(0), _lib.ptr(bias),
                                             out.data_ptr(), out.stride(0), M, N, K, act, _stream(A)), "ctr_linear_fwd")
    return out


@_guarded
def split_tf32(x: torch.Tensor, axis: int, role: int) -> torch.Tensor:
    """3xTF32 operand of ``x`` f32 [rows, cols] (``ctr_split_tf32``): hi / lo TF32 parts laid out as three segments along
    the reduction axis of the GEMM that consumes them -- axis 1: [rows, 3 * seg] (seg = cols rounded up to 4) for
    ``linear_fwd``; axis 0: [3 * rows, seg] for ``linear_wgrad``.  role 0 = left operand (lo, hi, hi), 1 = right (hi, lo, hi)."""
    _lib.require_cuda(x, "x")
    if x.dtype != torch.float32 or x.dim() != 2 or x.stride(1) != 1:
        raise ValueError("x must be f32 [rows, cols] with unit inner stride")
    rows, cols = x.shape
    seg = (cols + 3) // 4 * 4
    out = torch.empty((rows, 3 * seg) if axis == 1 else (3 * rows, seg), dtype=torch.float32, device=x.device)
    with _timed("split_tf32"):
        _lib.check(_lib.lib().ctr_split_tf32(x.data_ptr(), x.stride(0), rows, cols, out.data_ptr(), out.stride(0), axis, role,
                                             _stream(x)), "ctr_split_tf32")
    return out


@_guarded
def linear_fwd_bn_stats(A, W, bias, eps: float, momentum: float, running_mean=None, running_var=None, num_batches_tracked=None):
    """z = A @ W.T + bias together with the BatchNorm batch statistics of z, which come out of the GEMM epilogue
    (``ctr_linear_fwd_stats`` + ``ctr_bn_stats_from_partials``) -> (z, mean [N], rstd [N]); None when this M has no such path."""
    M, K = A.shape
    N = W.shape[0]
    blocks = int(_lib.lib().ctr_linear_stats_blocks(M))
    if blocks == 0 or N % 4:
        return None
    dev = A.device
    z = torch.empty(M, N, dtype=torch.float32, device=dev)
    stats = torch.empty(blocks * 2 * N, dtype=torch.float32, device=dev)
    mean = torch.empty(N, dtype=torch.float32, device=dev)
    rstd = torch.empty(N, dtype=torch.float32, device=dev)
    with _timed("linear_fwd"):
        _lib.check(_lib.lib().ctr_linear_fwd_stats(A.data_ptr(), A.stride(0), W.data_ptr(), W.stride(0), _lib.ptr(bias), z.data_ptr(),
                                                   z.stride(0), M, N, K, stats.data_ptr(), stats.numel(), _stream(A)), "ctr_linear_fwd_stats")
    with _timed("bn_stats"):
        _lib.check(_lib.lib().ctr_bn_stats_from_partials(stats.data_ptr(), blocks, M, N, eps, momentum, mean.data_ptr(), rstd.data_ptr(),
                                                         _lib.ptr(running_mean), _lib.ptr(running_var), _lib.ptr(num_batches_tracked),
                                                         _stream(A)), "ctr_bn_stats_from_partials")
    return z, mean, rstd


# ---- row-sharded tables over peer memory -----------------------------------------------------------------------
def ptr_array(ptrs):
    """ctypes array of device pointers (one per rank)."""
    return (C.c_void_p * len(ptrs))(*[int(p) for p in ptrs])


def make_shard(world: int, rank: int, adj: torch.Tensor) -> _lib.Shard:
    _chk(adj, "adj", torch.int64)
    return _lib.Shard(world, rank, adj.data_ptr())


@_guarded
def emb_pool_fwd_sharded(call: GroupCall, shard: _lib.Shard, tables, twin_tables=None) -> None:
    """``twin_tables``: per-rank pointers of the one-column twin shards (the fused DeepFM terms, ``make_group(extra=...)``)."""
    with _timed("emb_pool_fwd" + _width_tag(call)):
        if twin_tables is None and not call.struct.extra:
            _lib.check(_lib.lib().ctr_emb_pool_fwd_sharded(C.byref(call.struct), C.byref(shard), tables, _stream(call)),
                       "ctr_emb_pool_fwd_sharded")
        else:
            _lib.check(_lib.lib().ctr_emb_pool_fwd_sharded_ex(C.byref(call.struct), C.byref(shard), tables, twin_tables,
                                                              _stream(call)), "ctr_emb_pool_fwd_sharded_ex")


@_guarded
def route_p2p_workspace_bytes(call: GroupCall) -> int:
    return _lib.check(_lib.lib().ctr_route_p2p_workspace_bytes(C.byref(call.struct)), "ctr_route_p2p_workspace_bytes")


@_guarded
def route_p2p_build(call: GroupCall, shard: _lib.Shard, counts_ptr: int, keys_ptr: int, slots_ptr: int,
                    workspace: torch.Tensor) -> None:
    with _timed("route_p2p_build"):
        _lib.check(_lib.lib().ctr_route_p2p_build(C.byref(call.struct), C.byref(shard), counts_ptr, keys_ptr, slots_ptr,
                                                  workspace.data_ptr(), workspace.numel(), _stream(call)), "ctr_route_p2p_build")


@_guarded
def emb_bwd_p2p_workspace_bytes(call: GroupCall, world: int) -> int:
    return _lib.check(_lib.lib().ctr_emb_bwd_p2p_workspace_bytes(C.byref(call.struct), world), "ctr_emb_bwd_p2p_workspace_bytes")


@_guarded
def emb_bwd_plan_p2p(call: GroupCall, shard: _lib.Shard, counts, keys, slots, workspace: torch.Tensor) -> None:
    with _timed("emb_bwd_plan"):
        _lib.check(_lib.lib().ctr_emb_bwd_plan_p2p(C.byref(call.struct), C.byref(shard), counts, keys, slots,
                                                   workspace.data_ptr(), workspace.numel(), _stream(call)), "ctr_emb_bwd_plan_p2p")


@_guarded
def emb_bwd_apply_p2p(call: GroupCall, shard: _lib.Shard, workspace: torch.Tensor, opt: _lib.Opt, grads,
                      num_unique: torch.Tensor | None = None, peer_extra=None, peer_fm_sum=None) -> None:
    """``peer_extra`` (+ ``peer_fm_sum``): per-rank pointers of dL/d extra [B] (and fm_sum [B, D]) -- the fused DeepFM terms on
    the owner side; the group then carries ``extra`` and the features their twin shard slices."""
    _chk(num_unique, "num_unique", torch.int64)
    with _timed("emb_bwd_apply_owner" + _width_tag(call)):
        if peer_extra is None:
            _lib.check(_lib.lib().ctr_emb_bwd_apply_p2p(C.byref(call.struct), C.byref(shard), workspace.data_ptr(), C.byref(opt),
                                                        grads, _lib.ptr(num_unique), _stream(call)), "ctr_emb_bwd_apply_p2p")
        else:
            _lib.check(_lib.lib().ctr_emb_bwd_apply_p2p_ex(C.byref(call.struct), C.byref(shard), workspace.data_ptr(), C.byref(opt),
                                                           grads, peer_extra, peer_fm_sum, _lib.ptr(num_unique), _stream(call)),
                       "ctr_emb_bwd_apply_p2p_ex")


# ---- tower block (BatchNorm1d + ReLU + Dropout around a Linear) ------------------------------------------------------
_tower_ws: dict = {}


def tower_workspace(device, N: int, tag=None) -> torch.Tensor:
    """Scratch of the tower / head kernels for width N.  ``tag``: a workspace of its own (the calls whose per-block partials
    are finalised later on another stream must not share theirs with the kernels that run in between)."""
    key = (device, N, tag)
    ws = _tower_ws.get(key)
    if ws is None:
        ws = torch.empty(_lib.check(_lib.lib().ctr_tower_workspace_bytes(N)), dtype=torch.uint8, device=device)
        _tower_ws[key] = ws
    return ws


def _rows2d(t, name):
    _lib.require_cuda(t, name)
    if t.dtype != torch.float32 or t.dim() != 2 or t.stride(1) != 1:
        raise ValueError(f"{name} must be f32 [B, N] with unit inner stride")


@_guarded
def bn_stats(z, eps: float, momentum: float, running_mean=None, running_var=None, num_batches_tracked=None):
    """Batch statistics of z [B, N] -> (mean [N], rstd [N]); updates the running statistics in place."""
    _rows2d(z, "z")
    B, N = z.shape
    mean = torch.empty(N, dtype=torch.float32, device=z.device)
    rstd = torch.empty(N, dtype=torch.float32, device=z.device)
    with _timed("bn_stats"):
        _lib.check(_lib.lib().ctr_bn_stats(z.data_ptr(), z.stride(0), B, N, eps, momentum, mean.data_ptr(), rstd.data_ptr(),
                                           _lib.ptr(running_mean), _lib.ptr(running_var), _lib.ptr(num_batches_tracked),
                                           tower_workspace(z.device, N).data_ptr(), _stream(z)), "ctr_bn_stats")
    return mean, rstd


@_guarded
def bn_relu_dropout_fwd(z, mean, rstd, gamma, beta, p_drop: float, seed_dev, seed_offset: int):
    _rows2d(z, "z")
    B, N = z.shape
    y = torch.empty(B, N, dtype=torch.float32, device=z.device)
    with _timed("bn_act_fwd"):
        _lib.check(_lib.lib().ctr_bn_relu_dropout_fwd(z.data_ptr(), z.stride(0), B, N, mean.data_ptr(), rstd.data_ptr(),
                                                      gamma.data_ptr(), beta.data_ptr(), p_drop, _lib.ptr(seed_dev), seed_offset,
                                                      y.data_ptr(), y.stride(0), _stream(z)), "ctr_bn_relu_dropout_fwd")
    return y


@_guarded
def bn_relu_dropout_bwd(gy, z, mean, rstd, gamma, beta, p_drop: float, seed_dev, seed_offset: int, want_dbias: bool = True,
                        defer_dbias_tag=None):
    """-> (gz [B, N], dgamma [N], dbeta [N], dbias [N] or None).  ``defer_dbias_tag``: leave the bias gradient's per-block
    partials in a workspace of that tag and return ``dbias = None``; ``bn_bias_grad_deferred`` finishes it (on any stream)."""
    _rows2d(gy, "gy")
    _rows2d(z, "z")
    B, N = z.shape
    dev = z.device
    gz = torch.empty(B, N, dtype=torch.float32, device=dev)
    dgamma = torch.empty(N, dtype=torch.float32, device=dev)
    dbeta = torch.empty(N, dtype=torch.float32, device=dev)
    deferred = want_dbias and defer_dbias_tag is not None
    dbias = torch.empty(N, dtype=torch.float32, device=dev) if (want_dbias and not deferred) else None
    ws = tower_workspace(dev, N, defer_dbias_tag if deferred else None)
    with _timed("bn_act_bwd"):
        _lib.check(_lib.lib().ctr_bn_relu_dropout_bwd(gy.data_ptr(), gy.stride(0), z.data_ptr(), z.stride(0), B, N, mean.data_ptr(),
                                                      rstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(), p_drop,
                                                      _lib.ptr(seed_dev), seed_offset, gz.data_ptr(), gz.stride(0),
                                                      dgamma.data_ptr(), dbeta.data_ptr(), _lib.ptr(dbias),
                                                      ws.data_ptr(), _stream(z)), "ctr_bn_relu_dropout_bwd")
    return gz, dgamma, dbeta, dbias


@_guarded
def bn_bias_grad_deferred(gz: torch.Tensor, tag) -> torch.Tensor:
    """dbias [N] of the ``bn_relu_dropout_bwd(..., defer_dbias_tag=tag)`` call that produced ``gz`` [B, N] (the last one on
    (device, N, tag)); launched on the CURRENT stream, which the caller has ordered after that call."""
    B, N = gz.shape
    device = gz.device
    dbias = torch.empty(N, dtype=torch.float32, device=device)
    with _timed("bn_bias_grad"):
        _lib.check(_lib.lib().ctr_bn_bias_grad_from_partials(tower_workspace(device, N, tag).data_ptr(), B, N, dbias.data_ptr(),
                                                             _stream(dbias)), "ctr_bn_bias_grad_from_partials")
    return dbias


_wgrad_ws: dict = {}


@_guarded
def linear_wgrad(gz: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """``gz.T @ x`` -> [N, K]: the weight gradient of ``y = x W^T`` on tcgen05 (TF32 in, fp32 accumulate, split over
    the batch).  gz f32 [B, N], x f32 [B, K], unit inner stride, row pitches multiples of 4 floats."""
    _rows2d(gz, "gz")
    _rows2d(x, "x")
    B, N = gz.shape
    K = x.shape[1]
    if x.shape[0] != B:
        raise ValueError(f"batch mismatch: gz {tuple(gz.shape)} vs x {tuple(x.shape)}")
    need = _lib.check(_lib.lib().ctr_linear_wgrad_workspace_bytes(B, N, K))
    key = (gz.device, need)
    ws = _wgrad_ws.get(key)
    if ws is None:
        ws = _wgrad_ws[key] = torch.empty(need, dtype=torch.uint8, device=gz.device)
    out = torch.empty(N, K, dtype=torch.float32, device=gz.device)
    with _timed("linear_wgrad"):
        _lib.check(_lib.lib().ctr_linear_wgrad(gz.data_ptr(), gz.stride(0), x.data_ptr(), x.stride(0), B, N, K, out.data_ptr(),
                                               out.stride(0), ws.data_ptr(), ws.numel(), _stream(gz)), "ctr_linear_wgrad")
    return out


# ---- logit head + dense optimizer ----------------------------------------------------------------------------------
@_guarded
def logit_bce_fwd(h, w, bias, extra, labels, want_logits: bool = False, xe=None, we=None, be=None):
    """h [B, H], w [H], bias [1] | None, extra [B, *] | None (column 0), labels [B, *] (column 0)
    (xe [B, ne <= 32], we [ne], be [1] | None): a second linear term ``xe @ we + be`` added to the logit
    -> (loss [] mean BCE-with-logits, dz [B], logits [B] | None)"""
    _rows2d(h, "h")
    B, H = h.shape
    dev = h.device
    dz = torch.empty(B, dtype=torch.float32, device=dev)
    loss = torch.empty((), dtype=torch.float32, device=dev)
    logits = torch.empty(B, dtype=torch.float32, device=dev) if want_logits else None
    ne = 0
    if xe is not None:
        _rows2d(xe, "xe")
        ne = xe.shape[1]
    with _timed("head_fwd"):
        _lib.check(_lib.lib().ctr_logit_bce_fwd_ex(h.data_ptr(), h.stride(0), B, H, w.data_ptr(), _lib.ptr(bias), _lib.ptr(extra),
                                                   0 if extra is None else extra.stride(0), _lib.ptr(xe),
                                                   0 if xe is None else xe.stride(0), ne, _lib.ptr(we), _lib.ptr(be),
                                                   labels.data_ptr(), labels.stride(0), _lib.ptr(logits), dz.data_ptr(),
                                                   loss.data_ptr(), tower_workspace(dev, H).data_ptr(), _stream(h)),
                   "ctr_logit_bce_fwd_ex")
    return loss, dz, logits


@_guarded
def logit_bce_bwd(h, w, dz, gscale, want_gh: bool, want_gextra: bool, xe=None, defer_params_tag=None):
    """-> (gh [B, H] | None, gw [H], gb [1], gextra [B, 1] | None, gwe [ne] | None) for the upstream scalar gradient
    ``gscale`` (device); the second linear term's bias gradient equals gb.  ``defer_params_tag``: gw / gb / gwe come back as
    None, their per-block partials stay in a workspace of that tag for ``logit_bce_bwd_params``."""
    B, H = h.shape
    dev = h.device
    deferred = defer_params_tag is not None
    gh = torch.empty(B, H, dtype=torch.float32, device=dev) if want_gh else None
    gw = None if deferred else torch.empty(H, dtype=torch.float32, device=dev)
    gb = None if deferred else torch.empty(1, dtype=torch.float32, device=dev)
    gextra = torch.empty(B, 1, dtype=torch.float32, device=dev) if want_gextra else None
    ne = 0 if xe is None else xe.shape[1]
    gwe = torch.empty(ne, dtype=torch.float32, device=dev) if (xe is not None and not deferred) else None
    with _timed("head_bwd"):
        _lib.check(_lib.lib().ctr_logit_bce_bwd_ex(h.data_ptr(), h.stride(0), B, H, w.data_ptr(), dz.data_ptr(), gscale.data_ptr(),
                                                   _lib.ptr(gh), 0 if gh is None else gh.stride(0), _lib.ptr(gw), _lib.ptr(gb),
                                                   _lib.ptr(gextra), 1, _lib.ptr(xe), 0 if xe is None else xe.stride(0), ne,
                                                   _lib.ptr(gwe), tower_workspace(dev, H, defer_params_tag).data_ptr(), _stream(h)),
                   "ctr_logit_bce_bwd_ex")
    return gh, gw, gb, gextra, gwe


@_guarded
def logit_bce_bwd_params(h: torch.Tensor, gscale, ne: int, tag):
    """(gw [H], gb [1], gwe [ne] | None) of the ``logit_bce_bwd(h, ..., defer_params_tag=tag)`` call that ran last on
    (device, H, tag); launched on the CURRENT stream, which the caller has ordered after that call."""
    B, H = h.shape
    device = h.device
    gw = torch.empty(H, dtype=torch.float32, device=device)
    gb = torch.empty(1, dtype=torch.float32, device=device)
    gwe = torch.empty(ne, dtype=torch.float32, device=device) if ne else None
    with _timed("head_bwd_params"):
        _lib.check(_lib.lib().ctr_logit_bce_bwd_finalize(B, H, gscale.data_ptr(), gw.data_ptr(), gb.data_ptr(), ne, _lib.ptr(gwe),
                                                         tower_workspace(device, H, tag).data_ptr(), _stream(gw)),
                   "ctr_logit_bce_bwd_finalize")
    return gw, gb, gwe


@_guarded
def dense_adagrad(params, grads, sums, lr: float, eps: float) -> None:
    """One launch of torch.optim.Adagrad's update (lr_decay = weight_decay = 0) over up to 48 dense fp32 tensors."""
    n = len(params)
    if n == 0:
        return
    pa = (C.c_void_p * n)(*[p.data_ptr() for p in params])
    ga = (C.c_void_p * n)(*[g.data_ptr() for g in grads])
    sa = (C.c_void_p * n)(*[s.data_ptr() for s in sums])
    na = (C.c_int64 * n)(*[p.numel() for p in params])
    with _timed("dense_adagrad"):
        _lib.check(_lib.lib().ctr_dense_adagrad(n, pa, ga, sa, na, lr, eps, _stream(params[0])), "ctr_dense_adagrad")


# ---- de-duplicated exchange -------------------------------------------------------------------------------------------
@_guarded
def unique_fetch(call: GroupCall, shard: _lib.Shard, tables, plan_ws: torch.Tensor, staging: torch.Tensor,
                 uidx: torch.Tensor | None) -> None:
    _chk(staging, "staging", torch.float32)
    _chk(uidx, "uidx", torch.int64)
    with _timed("unique_fetch" + _width_tag(call)):
        _lib.check(_lib.lib().ctr_unique_fetch(C.byref(call.struct), C.byref(shard), tables, plan_ws.data_ptr(), staging.data_ptr(),
                                               _lib.ptr(uidx), _stream(call)), "ctr_unique_fetch")


@_guarded
def emb_bwd_p2p_unique_workspace_bytes(call: GroupCall, capacity: int) -> int:
    return _lib.check(_lib.lib().ctr_emb_bwd_p2p_unique_workspace_bytes(C.byref(call.struct), capacity),
                      "ctr_emb_bwd_p2p_unique_workspace_bytes")


@_guarded
def emb_bwd_plan_p2p_unique(call: GroupCall, shard: _lib.Shard, peer_nu, peer_uf, peer_ur, capacity: int,
                            workspace: torch.Tensor) -> None:
    with _timed("emb_bwd_plan_owner"):
        _lib.check(_lib.lib().ctr_emb_bwd_plan_p2p_unique(C.byref(call.struct), C.byref(shard), peer_nu, peer_uf, peer_ur, capacity,
                                                          workspace.data_ptr(), workspace.numel(), _stream(call)),
                   "ctr_emb_bwd_plan_p2p_unique")


@_guarded
def emb_bwd_apply_p2p_unique(call: GroupCall, shard: _lib.Shard, workspace: torch.Tensor, opt: _lib.Opt, peer_row_grads,
                             capacity: int) -> None:
    with _timed("emb_bwd_apply_owner" + _width_tag(call)):
        _lib.check(_lib.lib().ctr_emb_bwd_apply_p2p_unique(C.byref(call.struct), C.byref(shard), workspace.data_ptr(), C.byref(opt),
                                                           peer_row_grads, capacity, None, _stream(call)), "ctr_emb_bwd_apply_p2p_unique")
