"""Batch assembly for the fused lookup: drop-in for ``torchctr.dataset.get_dataloader`` (``torchctr/dataset.py:5-82``).

The reference ``collate_fn`` builds every list feature with ``pad_sequence`` + ``F.pad`` / slicing
(``torchctr/nn/functional.py:6-44``: three copies of a ``[B, maxlen]`` tensor) and leaves the batch in pageable memory.
``make_collate_fn`` returns a function with the same input (a list of per-sample dicts) and the same output --
``({'dense_features': f32 [B, Nd], '<sparse>': i64 [B, 1], '<list>': i64 [B, maxlen] right-padded,
'<list>_weight': f32 [B, maxlen]}, labels f32 [B, T])``, same keys in the same order, same dtypes, same truncation
-- but writes every feature once into its final tensor, optionally in pinned memory, so that
``GraphedTrainStep.prefetch`` / ``.to(device, non_blocking=True)`` overlap the upload with the running step.
Pure host code; the padded layout is what the lookup kernels consume (negative = padding, ``dataset.py:9``).
"""
from __future__ import annotations

import torch


def _as_1d(x, dtype):
    t = x if torch.is_tensor(x) else torch.as_tensor(x)
    return t.reshape(-1).to(dtype)


def make_collate_fn(feat_configs, target_cols, list_padding_value=-100, list_padding_maxlen=256, pin_memory: bool = False):
    """collate_fn(batch: list[dict]) -> (features dict, labels), as ``dataset.py:38-78`` builds them."""
    feat_configs = list(feat_configs)
    target_cols = list(target_cols)
    pin = bool(pin_memory) and torch.cuda.is_available()

    def new(shape, dtype, fill=None):
        t = torch.empty(shape, dtype=dtype, pin_memory=pin)
        if fill is not None:
            t.fill_(fill)
        return t

    def collate_fn(batch):
        B = len(batch)
        dense_names = [k["name"] for k in feat_configs if k["type"] == "dense"]
        dense = new((B, len(dense_names)), torch.float32)
        for j, name in enumerate(dense_names):
            dense[:, j] = torch.as_tensor([float(sample[name]) for sample in batch], dtype=torch.float32)
        sparse = {}
        for k in feat_configs:
            if k["type"] != "sparse":
                continue
            name = k["name"]
            if k.get("islist"):
                maxlen = k.get("maxlen", list_padding_maxlen)
                out = new((B, maxlen), torch.int64, k.get("padding_value", list_padding_value))
                weight_col = k.get("weight")
                wout = new((B, maxlen), torch.float32, 0.0) if weight_col else None
                for i, sample in enumerate(batch):
                    seq = _as_1d(sample[name], torch.int64)
                    n = min(seq.numel(), maxlen)             # truncate on the right (functional.py:38-42)
                    out[i, :n] = seq[:n]
                    if wout is not None:
                        w = _as_1d(sample[weight_col], torch.float32)
                        m = min(w.numel(), maxlen)
                        wout[i, :m] = w[:m]
                if wout is not None:
                    sparse[name + "_weight"] = wout          # the reference inserts the weight before the ids (dataset.py:67-70)
                sparse[name] = out
            else:
                col = new((B, 1), torch.int64)
                col[:, 0] = torch.as_tensor([int(sample[name]) for sample in batch], dtype=torch.int64)
                sparse[name] = col
        labels = new((B, len(target_cols)), torch.float32)
        for j, name in enumerate(target_cols):
            labels[:, j] = torch.as_tensor([float(sample[name]) for sample in batch], dtype=torch.float32)
        return {"dense_features": dense, **sparse}, labels

    return collate_fn


def get_dataloader(ds, feat_configs, target_cols, list_padding_value=-100, list_padding_maxlen=256, pin_memory: bool = False,
                   **kwargs):
    """Same signature and batch format as ``torchctr.dataset.get_dataloader``; ``ds`` is a huggingface
    ``datasets.Dataset`` (it is switched to torch format) or any map-style dataset of per-sample dicts."""
    if hasattr(ds, "with_format"):
        ds = ds.with_format("torch")
    collate = make_collate_fn(feat_configs, target_cols, list_padding_value, list_padding_maxlen, pin_memory)
    return torch.utils.data.DataLoader(ds, collate_fn=collate, **kwargs)
