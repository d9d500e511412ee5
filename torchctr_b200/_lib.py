"""ctypes binding of ``libctr_b200.so`` (the C ABI of ``include/ctr_b200.h``).

There is no CPU fallback: if the library cannot be loaded (or built with nvcc when it is
missing), importing an operator raises.  Tensors cross the boundary as raw device pointers.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from . import build as _build

MAX_FEATURES = 40
ABI_VERSION = 4          # CTR_B200_ABI_VERSION of include/ctr_b200.h

OK = 0
STATUS_INDEX_OOB = 1
STATUS_MAP_FULL = 2
STATUS_NO_RUNS = 4
PLAN_NO_RUNS = 1

INDEX_DIRECT, INDEX_HASH, INDEX_REMAP, INDEX_WINDOW = 0, 1, 2, 3
POOL_SUM, POOL_MEAN = 0, 1
OPT_NONE, OPT_SGD, OPT_ADAGRAD, OPT_ROWWISE_ADAGRAD, OPT_ADAM, OPT_GRAD_OUT = 0, 1, 2, 3, 4, 5
VOCAB_EMPTY = -(2 ** 63)


class VocabMap(C.Structure):
    _fields_ = [("keys", C.c_void_p), ("rows", C.c_void_p), ("capacity", C.c_int64)]


class Feature(C.Structure):
    _fields_ = [
        ("ids", C.c_void_p), ("id_weight", C.c_void_p), ("table", C.c_void_p),
        ("state0", C.c_void_p), ("state1", C.c_void_p), ("bag_scale", C.c_void_p),
        ("map", C.POINTER(VocabMap)),
        ("num_rows", C.c_int64),
        ("L", C.c_int32), ("D", C.c_int32), ("out_col", C.c_int32), ("pooling", C.c_int32),
        ("index_kind", C.c_int32), ("hash_seed", C.c_uint32),
        ("twin_table", C.c_void_p), ("twin_state0", C.c_void_p), ("twin_state1", C.c_void_p),
    ]


class Group(C.Structure):
    _fields_ = [
        ("features", C.POINTER(Feature)), ("num_features", C.c_int32), ("B", C.c_int32),
        ("out", C.c_void_p), ("out_stride", C.c_int64),
        ("dense", C.c_void_p), ("dense_width", C.c_int32), ("dense_col", C.c_int32),
        ("zero_from", C.c_int32),
        ("status", C.c_void_p),
        ("extra", C.c_void_p), ("fm_sum", C.c_void_p), ("fm", C.c_int32), ("grad_blocked", C.c_int32),
    ]


class Shard(C.Structure):
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("adj", C.c_void_p)]


MAX_WORLD = 16
PEER_HANDLE_BYTES = 64


class Opt(C.Structure):
    _fields_ = [("kind", C.c_int32), ("step", C.c_int32), ("lr", C.c_double), ("eps", C.c_double),
                ("beta1", C.c_double), ("beta2", C.c_double), ("device_hyper", C.c_void_p)]


class Hyper(C.Structure):
    _fields_ = [("lr", C.c_float), ("eps", C.c_float), ("one_minus_beta1", C.c_float),
                ("one_minus_beta2", C.c_float), ("adam_step_size", C.c_float)]


_P = C.c_void_p
_SIGNATURES = {
    # name: (restype, argtypes)
    "ctr_last_error_string": (C.c_char_p, []),
    "ctr_abi_version": (C.c_int, []),
    "ctr_kernel_launches": (C.c_int64, []),
    "ctr_opt_hyper": (None, [C.POINTER(Opt), C.POINTER(Hyper)]),
    "ctr_emb_pool_fwd": (C.c_int, [C.POINTER(Group), _P]),
    "ctr_hash_bucket_i64": (C.c_int, [_P, C.c_int64, C.c_uint32, C.c_uint32, _P, _P]),
    "ctr_hash_bucket_bytes": (C.c_int, [_P, _P, C.c_int64, C.c_uint32, C.c_uint32, _P, _P]),
    "ctr_rows_gather": (C.c_int, [_P, C.c_int64, _P, C.c_int64, C.c_int32, _P, _P, _P]),
    "ctr_normal_fill_rows": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int32, C.c_float, C.c_float, C.c_uint64, _P]),
    "ctr_normal_fill_rows_strided": (C.c_int, [_P, C.c_int64, C.c_int64, C.c_int32, C.c_float, C.c_float, C.c_uint64, C.c_int64,
                                               C.c_int64, _P]),
    "ctr_ids_minmax": (C.c_int, [_P, C.c_int64, _P, _P]),
    "ctr_emb_bwd_workspace_bytes": (C.c_int64, [C.POINTER(Group)]),
    "ctr_emb_bwd_plan": (C.c_int, [C.POINTER(Group), _P, C.c_int64, _P]),
    "ctr_emb_bwd_plan_ex": (C.c_int, [C.POINTER(Group), _P, C.c_int64, C.c_uint32, _P]),
    "ctr_emb_bwd_apply": (C.c_int, [C.POINTER(Group), _P, C.POINTER(Opt), _P, _P, _P, C.c_int64, _P, _P]),
    "ctr_vocab_fit_workspace_bytes": (C.c_int64, [C.c_int64]),
    "ctr_vocab_fit": (C.c_int, [C.POINTER(VocabMap), _P, C.c_int64, C.c_int32, _P, _P, _P, _P, C.c_int64, _P]),
    "ctr_vocab_transform": (C.c_int, [C.POINTER(VocabMap), _P, C.c_int64, C.c_int32, _P, _P]),
    "ctr_vocab_insert": (C.c_int, [C.POINTER(VocabMap), _P, _P, C.c_int64, _P, _P]),
    "ctr_vocab_clear": (C.c_int, [C.POINTER(VocabMap), _P]),
    "ctr_fm_fwd": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_int32, C.c_int32, _P, C.c_int32, C.c_int64, _P,
                             C.c_int64, C.c_int32, _P]),
    "ctr_fm_bwd": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_int32, C.c_int32, _P, C.c_int64, _P, C.c_int64,
                             C.c_int32, _P, C.c_int32, C.c_int64, _P]),
    "ctr_target_attention_fwd": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _P, _P]),
    "ctr_target_attention_bwd": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _P, _P, _P, _P, _P]),
    "ctr_cross_combine_fwd": (C.c_int, [_P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int64, _P, _P]),
    "ctr_cross_combine_bwd": (C.c_int, [_P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int64, _P, _P, C.c_int32, _P]),
    "ctr_linear_fwd": (C.c_int, [_P, C.c_int64, _P, C.c_int64, _P, _P, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                 C.c_int32, _P]),
    "ctr_linear_stats_blocks": (C.c_int32, [C.c_int32]),
    "ctr_linear_fwd_stats": (C.c_int, [_P, C.c_int64, _P, C.c_int64, _P, _P, C.c_int64, C.c_int32, C.c_int32, C.c_int32, _P, C.c_int64, _P]),
    "ctr_bn_stats_from_partials": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float, _P, _P, _P, _P, _P, _P]),
    "ctr_linear_fwd_blocked": (C.c_int, [_P, C.c_int64, _P, C.c_int64, _P, _P, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                         C.c_int32, _P]),
    "ctr_split_tf32": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_int32, _P, C.c_int64, C.c_int32, C.c_int32, _P]),
    "ctr_route_workspace_bytes": (C.c_int64, [C.POINTER(Group), C.c_int32]),
    "ctr_route_build": (C.c_int, [C.POINTER(Group), C.c_int32, _P, _P, _P, _P, _P, C.c_int64, _P]),
    "ctr_route_grad_gather": (C.c_int, [C.POINTER(Group), C.c_int32, _P, C.c_int64, C.c_int32, _P, _P]),
    "ctr_linear_wgrad_workspace_bytes": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32]),
    "ctr_linear_wgrad": (C.c_int, [_P, C.c_int64, _P, C.c_int64, C.c_int32, C.c_int32, C.c_int32, _P, C.c_int64, _P, C.c_int64, _P]),
    "ctr_tower_workspace_bytes": (C.c_int64, [C.c_int32]),
    "ctr_bn_stats": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_int32, C.c_float, C.c_float, _P, _P, _P, _P, _P, _P, _P]),
    "ctr_bn_relu_dropout_fwd": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_int32, _P, _P, _P, _P, C.c_float, _P, C.c_uint64,
                                          _P, C.c_int64, _P]),
    "ctr_bn_relu_dropout_bwd": (C.c_int, [_P, C.c_int64, _P, C.c_int64, C.c_int32, C.c_int32, _P, _P, _P, _P, C.c_float, _P,
                                          C.c_uint64, _P, C.c_int64, _P, _P, _P, _P, _P]),
    "ctr_bn_bias_grad_from_partials": (C.c_int, [_P, C.c_int32, C.c_int32, _P, _P]),
    "ctr_logit_bce_bwd_finalize": (C.c_int, [C.c_int32, C.c_int32, _P, _P, _P, C.c_int32, _P, _P, _P]),
    "ctr_logit_bce_fwd": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_int32, _P, _P, _P, C.c_int64, _P, C.c_int64, _P, _P, _P, _P, _P]),
    "ctr_logit_bce_bwd": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_int32, _P, _P, _P, _P, C.c_int64, _P, _P, _P, C.c_int64, _P, _P]),
    "ctr_logit_bce_fwd_ex": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_int32, _P, _P, _P, C.c_int64, _P, C.c_int64, C.c_int32, _P, _P, _P,
                                       C.c_int64, _P, _P, _P, _P, _P]),
    "ctr_logit_bce_bwd_ex": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_int32, _P, _P, _P, _P, C.c_int64, _P, _P, _P, C.c_int64, _P,
                                       C.c_int64, C.c_int32, _P, _P, _P]),
    "ctr_dense_adagrad": (C.c_int, [C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                    C.POINTER(C.c_int64), C.c_float, C.c_float, _P]),
    "ctr_peer_alloc": (C.c_int, [C.c_int64, C.POINTER(C.c_void_p)]),
    "ctr_peer_free": (C.c_int, [_P]),
    "ctr_peer_export": (C.c_int, [_P, _P]),
    "ctr_peer_open": (C.c_int, [_P, C.POINTER(C.c_void_p)]),
    "ctr_peer_close": (C.c_int, [_P]),
    "ctr_fm_pack_grads": (C.c_int, [_P, C.c_int64, _P, _P, C.c_int32, C.c_int32, C.POINTER(C.c_int32), C.c_int32, _P, _P]),
    "ctr_rows_dense_apply": (C.c_int, [C.POINTER(Opt), _P, _P, _P, C.c_int64, C.c_int32, _P]),
    "ctr_emb_pool_fwd_sharded": (C.c_int, [C.POINTER(Group), C.POINTER(Shard), C.POINTER(C.c_void_p), _P]),
    "ctr_emb_pool_fwd_sharded_ex": (C.c_int, [C.POINTER(Group), C.POINTER(Shard), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), _P]),
    "ctr_route_p2p_workspace_bytes": (C.c_int64, [C.POINTER(Group)]),
    "ctr_route_p2p_build": (C.c_int, [C.POINTER(Group), C.POINTER(Shard), _P, _P, _P, _P, C.c_int64, _P]),
    "ctr_emb_bwd_p2p_workspace_bytes": (C.c_int64, [C.POINTER(Group), C.c_int32]),
    "ctr_emb_bwd_plan_p2p": (C.c_int, [C.POINTER(Group), C.POINTER(Shard), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                       C.POINTER(C.c_void_p), _P, C.c_int64, _P]),
    "ctr_unique_fetch": (C.c_int, [C.POINTER(Group), C.POINTER(Shard), C.POINTER(C.c_void_p), _P, _P, _P, _P]),
    "ctr_emb_bwd_p2p_unique_workspace_bytes": (C.c_int64, [C.POINTER(Group), C.c_int64]),
    "ctr_emb_bwd_plan_p2p_unique": (C.c_int, [C.POINTER(Group), C.POINTER(Shard), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                              C.POINTER(C.c_void_p), C.c_int64, _P, C.c_int64, _P]),
    "ctr_emb_bwd_apply_p2p_unique": (C.c_int, [C.POINTER(Group), C.POINTER(Shard), _P, C.POINTER(Opt), C.POINTER(C.c_void_p),
                                               C.c_int64, _P, _P]),
    "ctr_emb_bwd_apply_p2p": (C.c_int, [C.POINTER(Group), C.POINTER(Shard), _P, C.POINTER(Opt), C.POINTER(C.c_void_p), _P, _P]),
    "ctr_emb_bwd_apply_p2p_ex": (C.c_int, [C.POINTER(Group), C.POINTER(Shard), _P, C.POINTER(Opt), C.POINTER(C.c_void_p),
                                           C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), _P, _P]),
}

_lib = None


def library_path() -> str:
    return _build.LIB_PATH


def lib():
    """The loaded library; builds it with nvcc first if the .so is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        _build.build()
    handle = C.CDLL(path)
    for name, (restype, argtypes) in _SIGNATURES.items():
        fn = getattr(handle, name)       # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = handle
    return _lib


def declared_symbols():
    return sorted(_SIGNATURES)


def check(rc: int, what: str = "libctr_b200") -> int:
    if rc < 0:
        msg = lib().ctr_last_error_string().decode("utf-8", "replace")
        if rc == -1:
            raise ValueError(f"{what}: {msg}")
        raise RuntimeError(f"{what}: {msg} (code {rc})")
    return rc


def ptr(t: torch.Tensor | None):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def stream_ptr(device=None):
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise RuntimeError(f"torchctr_b200: {name} must be a CUDA tensor -- the kernels have no CPU path "
                           f"(got device {t.device})")
