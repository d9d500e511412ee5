"""Generates the golden fixtures in this directory from the REAL reference.

Run in the build container only (``/root/reference`` must exist)::

    python tests/golden/make_golden.py

It imports ``torchctr`` from ``/root/reference`` (with the 6-line ``polars`` stub
of SURVEY.md section 8c, the only missing dependency) and scikit-learn's
``murmurhash3_32`` and records their outputs on seeded inputs:

* ``hash_golden.json``     -- ``torchctr.utils.hash_bucket`` / ``murmurhash3_32`` values
* ``dnn_golden.pt``        -- ``torchctr.models.DNN`` forward / loss / gradients / one
  ``torch.optim.Adagrad`` step and a short ``torchctr.trainer.Trainer.fit`` loss trace
* ``dynamic_golden.pt``    -- ``torchctr.nn.DynamicEmbedding`` growth and state-dict merge
* ``optim_golden.pt``      -- ``torch.optim.{Adagrad,SparseAdam,SGD}`` on an embedding table
* ``collate_golden.pt``    -- ``torchctr.dataset.get_dataloader`` batches (padding / truncation / weights)

The fixtures are small and committed; tests never import the reference.
"""
import importlib.machinery
import json
import os
import sys
import types

import datasets  # noqa: F401  (must be imported before the polars stub)
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def import_reference():
    stub = types.ModuleType("polars")
    stub.__spec__ = importlib.machinery.ModuleSpec("polars", None)
    stub.DataFrame = stub.Expr = type("X", (), {})
    sys.modules["polars"] = stub
    sys.path.insert(0, "/root/reference")
    import torchctr  # noqa: F401
    return torchctr


def golden_hash(torchctr):
    from sklearn.utils import murmurhash3_32
    from torchctr.utils import hash_bucket
    rng = np.random.default_rng(20240601)
    values = ["", "a", "abc", "123", "__null__", "other", "B001NPEBGU", "hello world", "ünïcode"]
    ints = [0, 1, 9, 10, 42, 99, 100, 12345, 2147483647, -1, -100, -2147483648, 2**40 + 7, 2**63 - 1, -2**63]
    ints += [int(x) for x in rng.integers(0, 2**31, 64)] + [int(x) for x in rng.integers(-2**62, 2**62, 32)]
    rows = []
    for seed in (0, 1, 7, 2**32 - 1):
        for v in values + ints:
            for buckets in (100, 1000003, 2**31 - 1):
                rows.append({"v": v, "is_int": isinstance(v, int), "seed": seed, "buckets": buckets,
                             "murmur": int(murmurhash3_32(str(v), seed=seed, positive=True)),
                             "bucket": int(hash_bucket(v, buckets, seed=seed))})
    with open(os.path.join(HERE, "hash_golden.json"), "w") as f:
        json.dump(rows, f)
    print("hash_golden.json", len(rows))


def make_feat_configs():
    return [
        {"name": "user", "type": "sparse", "num_embeddings": 50, "emb_dim": 16},
        {"name": "item", "type": "sparse", "num_embeddings": 97, "emb_dim": 16},
        {"name": "cate", "type": "sparse", "num_embeddings": 11, "emb_dim": 9},
        {"name": "hist", "type": "sparse", "num_embeddings": 97, "emb_dim": 16, "islist": True, "maxlen": 12},
        {"name": "price", "type": "dense"},
        {"name": "age", "type": "dense"},
    ]


def make_batch(gen, B, feat_configs, pad=-100):
    feats = {}
    for c in feat_configs:
        if c["type"] != "sparse":
            continue
        V = c["num_embeddings"]
        if c.get("islist"):
            L = c["maxlen"]
            ids = torch.randint(0, V, (B, L), generator=gen)
            lens = torch.randint(0, L + 1, (B,), generator=gen)
            ids[torch.arange(L).unsqueeze(0) >= lens.unsqueeze(1)] = pad
            feats[c["name"]] = ids
        else:
            feats[c["name"]] = torch.randint(0, V, (B, 1), generator=gen)
    nd = sum(1 for c in feat_configs if c["type"] == "dense")
    feats["dense_features"] = torch.randn(B, nd, generator=gen)
    labels = (torch.rand(B, 1, generator=gen) < 0.25).float()
    return feats, labels


def golden_dnn(torchctr):
    from torchctr.models import DNN
    from torchctr.trainer import Trainer
    fc = make_feat_configs()
    gen = torch.Generator().manual_seed(1234)
    feats, labels = make_batch(gen, 64, fc)

    torch.manual_seed(0)
    model = DNN(fc, hidden_units=[32, 16])
    init_state = {k: v.clone() for k, v in model.state_dict().items()}

    model.eval()
    with torch.no_grad():
        eval_logits = model(feats).clone()

    # training-mode forward/backward with dropout disabled (p=0) so it is RNG free
    for m in model.tower:
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    model.train()
    opt = torch.optim.Adagrad(model.parameters(), lr=0.05)
    opt.zero_grad()
    loss = model.training_step((feats, labels), 0)
    loss.backward()
    grads = {k: p.grad.clone() for k, p in model.named_parameters()}
    opt.step()
    after_step = {k: v.clone() for k, v in model.state_dict().items()}

    # a short run of the real Trainer (use_accelerate=False), dropout still 0
    torch.manual_seed(0)
    model2 = DNN(fc, hidden_units=[32, 16])
    for m in model2.tower:
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    batches = [make_batch(gen, 64, fc) for _ in range(6)]
    trace = []
    trainer = Trainer(model2, optimizer=torch.optim.Adagrad(model2.parameters(), lr=0.05), max_epochs=1,
                      use_accelerate=False, log_steps=1000,
                      callback_train_epoch_end=lambda rets: trace.extend(rets))
    trainer.fit(batches, batches[:1])
    torch.save({
        "feat_configs": fc, "hidden_units": [32, 16],
        "feats": feats, "labels": labels, "init_state": init_state,
        "eval_logits": eval_logits, "train_loss": loss.detach().clone(), "grads": grads,
        "after_adagrad_step": after_step, "adagrad_lr": 0.05,
        "trainer_batches": batches, "trainer_loss_trace": torch.tensor(trace),
        "trainer_final_state": {k: v.clone() for k, v in model2.state_dict().items()},
    }, os.path.join(HERE, "dnn_golden.pt"))
    print("dnn_golden.pt loss", float(loss), "trace", trace)


def golden_dynamic(torchctr):
    from torchctr.nn import DynamicEmbedding

    class Two(torch.nn.Module):
        def __init__(self, n1, n2):
            super().__init__()
            self.emb1 = DynamicEmbedding(n1, 5)
            self.emb2 = DynamicEmbedding(n2, 5)

    torch.manual_seed(0)
    m = Two(5, 6)
    w_before = m.emb1.weight.detach().clone()
    ids = torch.arange(1, 10).reshape(3, 3)
    torch.manual_seed(11)
    out = m.emb1(ids).detach().clone()
    w_after = m.emb1.weight.detach().clone()
    sd = {k: v.clone() for k, v in m.state_dict().items()}

    torch.manual_seed(5)
    bigger = Two(8, 15)                       # emb1 smaller than ckpt (grows), emb2 larger (pads)
    emb2_tail = bigger.emb2.weight.detach()[6:].clone()
    torch.manual_seed(6)
    bigger.load_state_dict({k: v.clone() for k, v in sd.items()})
    errs = {}
    for name, bad in (("empty", torch.zeros(0, dtype=torch.long)), ("negative", torch.tensor([[1, -1]]))):
        try:
            m.emb1(bad)
        except ValueError as e:
            errs[name] = str(e)
    torch.save({
        "w_before": w_before, "ids": ids, "out": out, "w_after": w_after, "state_dict": sd,
        "loaded_emb1": bigger.emb1.weight.detach().clone(), "loaded_emb2": bigger.emb2.weight.detach().clone(),
        "emb2_tail_before_load": emb2_tail, "errors": errs,
    }, os.path.join(HERE, "dynamic_golden.pt"))
    print("dynamic_golden.pt", tuple(w_after.shape), tuple(bigger.emb1.weight.shape), tuple(bigger.emb2.weight.shape), errs)


def golden_optim():
    """torch.optim on an embedding table: three steps, Zipf-ish duplicate ids."""
    gen = torch.Generator().manual_seed(77)
    V, D, B, L = 40, 8, 32, 5
    w0 = torch.randn(V, D, generator=gen)
    steps = []
    for _ in range(3):
        ids = torch.randint(0, V, (B, L), generator=gen)
        ids[torch.rand(B, L, generator=gen) < 0.3] = -100
        ids[:, 0] = torch.randint(0, 3, (B,), generator=gen)     # hot rows, many duplicates
        steps.append((ids, torch.randn(B, D, generator=gen)))
    out = {"w0": w0, "steps": steps}

    def run(make_opt, sparse):
        emb = torch.nn.Embedding(V, D, sparse=sparse, _weight=w0.clone())
        opt = make_opt(emb.parameters())
        for ids, gout in steps:
            opt.zero_grad()
            keep = (ids >= 0)
            pooled = (emb(ids * keep) * keep.unsqueeze(-1).float()).sum(1)
            pooled.backward(gout)
            opt.step()
        return emb.weight.detach().clone(), opt.state_dict()["state"]

    w, st = run(lambda p: torch.optim.Adagrad(p, lr=0.1, eps=1e-10), False)
    out["adagrad"] = {"w": w, "sum": st[0]["sum"].clone(), "lr": 0.1, "eps": 1e-10}
    w, st = run(lambda p: torch.optim.SparseAdam(list(p), lr=0.01, betas=(0.9, 0.999), eps=1e-8), True)
    out["sparse_adam"] = {"w": w, "exp_avg": st[0]["exp_avg"].clone(), "exp_avg_sq": st[0]["exp_avg_sq"].clone(),
                          "lr": 0.01, "betas": (0.9, 0.999), "eps": 1e-8}
    w, st = run(lambda p: torch.optim.SGD(p, lr=0.1), False)
    out["sgd"] = {"w": w, "lr": 0.1}
    torch.save(out, os.path.join(HERE, "optim_golden.pt"))
    print("optim_golden.pt")


def golden_collate(torchctr):
    """``torchctr.dataset.get_dataloader`` (dataset.py:5-82) on a small in-memory datasets.Dataset: the raw columns and
    the batches its collate_fn produces (padding, truncation, weights, dtypes, key order)."""
    import datasets
    from torchctr.dataset import get_dataloader
    rng = np.random.default_rng(7)
    n = 23
    lens = rng.integers(0, 9, n)
    lens[3] = 0
    lens[5] = 8
    cols = {
        "price": [float(x) for x in rng.normal(size=n)],
        "age": [float(x) for x in rng.integers(18, 80, n)],
        "user": [int(x) for x in rng.integers(0, 1000, n)],
        "cat": [int(x) for x in rng.integers(0, 50, n)],
        "hist": [[int(x) for x in rng.integers(0, 500, l)] for l in lens],
        "hist_w": [[float(x) for x in rng.random(l)] for l in lens],
        "tags": [[int(x) for x in rng.integers(0, 30, max(1, l // 2))] for l in lens],
        "click": [float(x) for x in rng.integers(0, 2, n)],
        "buy": [float(x) for x in rng.integers(0, 2, n)],
    }
    feat_configs = [
        {"name": "price", "type": "dense"}, {"name": "user", "type": "sparse", "num_embeddings": 1000, "emb_dim": 16},
        {"name": "hist", "type": "sparse", "islist": True, "maxlen": 6, "weight": "hist_w", "num_embeddings": 500, "emb_dim": 16},
        {"name": "age", "type": "dense"}, {"name": "cat", "type": "sparse", "num_embeddings": 50, "emb_dim": 16},
        {"name": "tags", "type": "sparse", "islist": True, "padding_value": -1, "num_embeddings": 30, "emb_dim": 16},
    ]
    target_cols = ["click", "buy"]
    ds = datasets.Dataset.from_dict(cols)
    out = {"columns": cols, "feat_configs": feat_configs, "target_cols": target_cols, "runs": []}
    for kw in (dict(list_padding_maxlen=5), dict(list_padding_value=-7, list_padding_maxlen=12)):
        dl = get_dataloader(ds, feat_configs, target_cols, batch_size=8, shuffle=False, **kw)
        out["runs"].append({"kwargs": kw, "batches": [(dict(f), l) for f, l in dl]})
    torch.save(out, os.path.join(HERE, "collate_golden.pt"))
    print("collate_golden.pt", [len(r["batches"]) for r in out["runs"]])


def golden_attention(torchctr):
    """torchctr.nn.functional.target_attention (nn/functional.py:46-74) incl. its gradients, with and without a mask: the
    reference's ``masked_fill`` is not in place (:63), so the mask has NO effect -- the fixture records exactly that."""
    from torchctr.nn.functional import target_attention
    gen = torch.Generator().manual_seed(11)
    cases = []
    for B, N, E in ((7, 5, 16), (16, 50, 16), (2, 200, 64), (9, 3, 4), (5, 17, 32)):
        tgt = torch.randn(B, E, generator=gen).requires_grad_(True)
        cand = torch.randn(B, N, E, generator=gen).requires_grad_(True)
        lens = torch.randint(0, N + 1, (B, 1), generator=gen)
        mask = (torch.arange(N)[None, :] < lens).float()
        gout = torch.randn(B, E, generator=gen)
        outs = {}
        for name, m in (("nomask", None), ("mask", mask)):
            tgt.grad = cand.grad = None
            out = target_attention(tgt, cand, m)
            out.backward(gout)
            outs[name] = {"out": out.detach().clone(), "gtarget": tgt.grad.clone(), "gcand": cand.grad.clone()}
        same = all(torch.equal(outs["mask"][k], outs["nomask"][k]) for k in outs["mask"])      # the reference ignores its mask
        cases.append({"target": tgt.detach().clone(), "cand": cand.detach().clone(), "mask": mask, "gout": gout,
                      "mask_ignored_by_reference": same, **outs["nomask"]})
    torch.save(cases, os.path.join(HERE, "attention_golden.pt"))
    print("attention_golden.pt", len(cases))


if __name__ == "__main__":
    ref = import_reference()
    if "--only-collate" in sys.argv:
        golden_collate(ref)
        sys.exit(0)
    if "--only-attention" in sys.argv:
        golden_attention(ref)
        sys.exit(0)
    golden_hash(ref)
    golden_dnn(ref)
    golden_dynamic(ref)
    golden_optim()
    golden_collate(ref)
    golden_attention(ref)
