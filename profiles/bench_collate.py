"""Host-side batch assembly, cfg1 shape (B = 4096: 8 categorical + 1 sequence of <= 50 ids + 4 numerical + 1 label):
the reference's collate_fn (torchctr/dataset.py:38-78, imported from baseline/_ref) vs torchctr_b200.data.make_collate_fn
(same per-sample input) vs ColumnarBatches (columns / CSR in, no per-sample Python).  CPU only; prints samples/s."""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baseline import refshim
from torchctr_b200.data import ColumnarBatches, make_collate_fn

B, NB = 4096, 4
rng = np.random.default_rng(0)
fc = [{"name": f"c{i}", "type": "sparse", "num_embeddings": 10 ** 5, "emb_dim": 16} for i in range(8)]
fc.append({"name": "hist", "type": "sparse", "num_embeddings": 5 * 10 ** 5, "emb_dim": 16, "islist": True, "maxlen": 50})
fc += [{"name": f"d{i}", "type": "dense"} for i in range(4)]
n = B * NB
cols = {f"c{i}": rng.integers(0, 10 ** 5, n) for i in range(8)}
lens = rng.integers(0, 51, n)
cols["hist"] = [rng.integers(0, 5 * 10 ** 5, int(l)).tolist() for l in lens]
for i in range(4):
    cols[f"d{i}"] = rng.standard_normal(n).astype(np.float32)
cols["y"] = (rng.random(n) < 0.25).astype(np.float32)
import datasets
ds = datasets.Dataset.from_dict({k: (v if isinstance(v, list) else v.tolist()) for k, v in cols.items()})

def run(name, it):
    t0 = time.perf_counter(); k = 0
    for feats, labels in it:
        k += labels.shape[0]
    dt = time.perf_counter() - t0
    print(f"{name:58s} {k / dt:12.0f} samples/s  ({1e3 * dt / NB:8.1f} ms per batch of {B})")

ref = refshim.load_reference()
run("reference get_dataloader (dataset.py:6-82, num_workers=0)", ref.dataset.get_dataloader(ds, fc, ["y"], batch_size=B, list_padding_maxlen=50))
samples = list(ds.with_format("torch"))
coll = make_collate_fn(fc, ["y"], list_padding_maxlen=50)
run("torchctr_b200.data.make_collate_fn (per-sample dicts in)", (coll(samples[i:i + B]) for i in range(0, n, B)))
t0 = time.perf_counter(); cb = ColumnarBatches.from_dataset(ds, fc, ["y"], B, list_padding_maxlen=50); t_setup = time.perf_counter() - t0
run(f"ColumnarBatches.from_dataset (Arrow CSR, setup {1e3 * t_setup:.0f} ms)", cb)
