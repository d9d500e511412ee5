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


# ---- columnar batch assembly: no per-sample Python at all -----------------------------------------------------------------
class ColumnarBatches:
    """The batches of ``torchctr.dataset.get_dataloader`` (``torchctr/dataset.py:38-78``: same keys, order, dtypes, padding and
    truncation) cut straight out of COLUMNS instead of per-sample dicts.

    The reference collate runs a Python comprehension per feature per sample (``dataset.py:48,52,69,76``) and pads every list
    feature three times (``nn/functional.py:6-44``); at B200 step times that is the bottleneck of the whole loop (SURVEY.md 8f
    rank 1).  Here a list feature is held as CSR -- ``values`` (all ids back to back) + ``offsets`` [n + 1], which is exactly an
    Arrow ``ListArray`` -- and a batch is a handful of vectorised numpy slices written once into (optionally pinned) tensors.
    ``ColumnarBatches.from_dataset(ds, ...)`` reads the columns of a huggingface ``datasets.Dataset`` through its Arrow table
    without materialising Python objects.  ``csr=True`` additionally emits ``<name>_values`` (i32) and ``<name>_offsets`` (i32,
    [B + 1], after truncation to ``maxlen``) for consumers that do not want the padding.
    """

    def __init__(self, columns: dict, feat_configs, target_cols, batch_size: int, list_padding_value=-100, list_padding_maxlen=256,
                 pin_memory: bool = False, drop_last: bool = False, shuffle: bool = False, seed: int = 0, csr: bool = False):
        import numpy as np
        self.np = np
        self.feat_configs = list(feat_configs)
        self.target_cols = list(target_cols)
        self.bs = int(batch_size)
        self.pad, self.maxlen = list_padding_value, list_padding_maxlen
        self.pin = bool(pin_memory) and torch.cuda.is_available()
        self.drop_last, self.shuffle, self.seed, self.csr = drop_last, shuffle, seed, csr
        self.cols = {}
        n = None
        for name, col in columns.items():
            if isinstance(col, tuple):                      # (values, offsets)
                vals, offs = np.asarray(col[0]), np.asarray(col[1], dtype=np.int64)
                self.cols[name] = (vals, offs)
                m = offs.shape[0] - 1
            else:
                arr = np.asarray(col)
                self.cols[name] = arr
                m = arr.shape[0]
            n = m if n is None else n
            if m != n:
                raise ValueError(f"column {name!r} has {m} rows, expected {n}")
        self.n = n or 0

    @classmethod
    def from_lists(cls, columns: dict, *args, **kwargs):
        """Columns given as Python lists (list features as lists of lists): converted to CSR once, up front."""
        import numpy as np
        out = {}
        for name, col in columns.items():
            if len(col) and isinstance(col[0], (list, tuple)) or (len(col) and hasattr(col[0], "__len__") and not isinstance(col[0], str)):
                lens = np.fromiter((len(x) for x in col), dtype=np.int64, count=len(col))
                offs = np.zeros(len(col) + 1, dtype=np.int64)
                np.cumsum(lens, out=offs[1:])
                vals = np.concatenate([np.asarray(x) for x in col]) if offs[-1] else np.zeros(0)
                out[name] = (vals, offs)
            else:
                out[name] = np.asarray(col)
        return cls(out, *args, **kwargs)

    @classmethod
    def from_dataset(cls, ds, feat_configs, target_cols, *args, **kwargs):
        """From a huggingface ``datasets.Dataset``: list columns become (values, offsets) views of the Arrow buffers."""
        import pyarrow as pa
        table = ds.data.table if hasattr(ds.data, "table") else ds.data
        wanted = [k["name"] for k in feat_configs] + [k["weight"] for k in feat_configs if k.get("weight")] + list(target_cols)
        cols = {}
        for name in dict.fromkeys(wanted):
            arr = table.column(name).combine_chunks()
            if pa.types.is_list(arr.type) or pa.types.is_large_list(arr.type):
                offs = arr.offsets.to_numpy(zero_copy_only=False).astype("int64")
                cols[name] = (arr.values.to_numpy(zero_copy_only=False), offs - offs[0])
            else:
                cols[name] = arr.to_numpy(zero_copy_only=False)
        return cls(cols, feat_configs, target_cols, *args, **kwargs)

    def __len__(self):
        return self.n // self.bs if self.drop_last else (self.n + self.bs - 1) // self.bs

    def _new(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, pin_memory=self.pin)

    def _rows(self, idx, name, dtype):
        np = self.np
        col = self.cols[name]
        return torch.from_numpy(np.ascontiguousarray(col[idx]).astype(dtype, copy=False))

    def _padded(self, idx, name, maxlen, pad_value, np_dtype, torch_dtype):
        """-> (padded [B, maxlen], values, offsets after truncation): right-padded, truncated on the right."""
        np = self.np
        vals, offs = self.cols[name]
        start = offs[idx]
        lens = np.minimum(offs[idx + 1] - start, maxlen)
        B = idx.shape[0]
        out_offs = np.zeros(B + 1, dtype=np.int64)
        np.cumsum(lens, out=out_offs[1:])
        total = int(out_offs[-1])
        row = np.repeat(np.arange(B), lens)
        col = np.arange(total) - np.repeat(out_offs[:-1], lens)
        flat = vals[np.repeat(start, lens) + col].astype(np_dtype, copy=False)
        out = self._new((B, maxlen), torch_dtype)
        view = out.numpy()
        view[...] = pad_value         # (numpy, not Tensor.fill_: torch's CPU thread pool costs tens of ms on a tensor this small)
        view[row, col] = flat
        return out, flat, out_offs

    def __iter__(self):
        np = self.np
        order = np.arange(self.n)
        if self.shuffle:
            np.random.default_rng(self.seed).shuffle(order)
            self.seed += 1
        stop = self.n - self.n % self.bs if self.drop_last else self.n
        for i0 in range(0, stop, self.bs):
            idx = order[i0:min(i0 + self.bs, stop)]
            yield self.batch(idx)

    def batch(self, idx):
        np = self.np
        B = idx.shape[0]
        dense_names = [k["name"] for k in self.feat_configs if k["type"] == "dense"]
        dense = self._new((B, len(dense_names)), torch.float32)
        for j, name in enumerate(dense_names):
            dense.numpy()[:, j] = self.cols[name][idx]
        sparse = {}
        for k in self.feat_configs:
            if k["type"] != "sparse":
                continue
            name = k["name"]
            if k.get("islist"):
                maxlen = k.get("maxlen", self.maxlen)
                out, flat, offs = self._padded(idx, name, maxlen, k.get("padding_value", self.pad), np.int64, torch.int64)
                if k.get("weight"):
                    w, _, _ = self._padded(idx, k["weight"], maxlen, 0.0, np.float32, torch.float32)
                    sparse[name + "_weight"] = w                 # the reference inserts the weight before the ids (dataset.py:67-70)
                sparse[name] = out
                if self.csr:
                    sparse[name + "_values"] = torch.from_numpy(flat.astype(np.int32))
                    sparse[name + "_offsets"] = torch.from_numpy(offs.astype(np.int32))
            else:
                col = self._new((B, 1), torch.int64)
                col.numpy()[:, 0] = self.cols[name][idx]
                sparse[name] = col
        labels = self._new((B, len(self.target_cols)), torch.float32)
        for j, name in enumerate(self.target_cols):
            labels.numpy()[:, j] = self.cols[name][idx]
        return {"dense_features": dense, **sparse}, labels
