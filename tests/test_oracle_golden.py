"""Pins the CPU oracle against fixtures produced by the real reference
(tests/golden/make_golden.py): hashing, DNN forward/backward/Adagrad step,
Trainer loss trace, DynamicEmbedding growth, torch optimizers on unique rows."""
import ctypes
import json
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import embedding as oe
from oracle import hashing as oh
from oracle import models as om
from oracle import optim as oo
from oracle.vocab import Vocab

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def hash_rows(golden_dir):
    with open(os.path.join(golden_dir, "hash_golden.json")) as f:
        return json.load(f)


def test_murmur_python_matches_sklearn_golden(hash_rows):
    for r in hash_rows:
        assert oh.murmur3_32(str(r["v"]).encode("utf-8"), r["seed"]) == r["murmur"]
        assert oh.hash_bucket(r["v"], r["buckets"], r["seed"]) == r["bucket"]


def test_murmur_vectorised_ints_match_golden(hash_rows):
    rows = [r for r in hash_rows if r["is_int"]]
    for seed in sorted({r["seed"] for r in rows}):
        for buckets in sorted({r["buckets"] for r in rows}):
            sel = [r for r in rows if r["seed"] == seed and r["buckets"] == buckets]
            ids = np.array([r["v"] for r in sel], dtype=np.int64)
            assert (oh.murmur3_decimal(ids, seed) == np.array([r["murmur"] for r in sel], dtype=np.uint32)).all()
            assert (oh.hash_bucket_ids(ids, buckets, seed) == np.array([r["bucket"] for r in sel])).all()


def test_survey_known_answers():
    # SURVEY.md section 8c (sklearn 1.9.0)
    assert oh.murmur3_32(b"", 0) == 0
    assert oh.murmur3_32(b"a", 0) == 1009084850
    assert oh.murmur3_32(b"abc", 0) == 3017643002
    assert oh.murmur3_32(b"B001NPEBGU", 7) == 1989969711
    assert [oh.hash_bucket(i, 1000003, 0) for i in (0, 1, 42, 2147483647)] == [659617, 506487, 916337, 309571]
    with pytest.raises(OverflowError):
        oh.murmur3_32(b"x", 2 ** 32)


def test_c_oracle_matches_golden(hash_rows):
    so = os.path.join(ROOT, "oracle", "_build", "liboracle_ctr.so")
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    lib = ctypes.CDLL(so)
    lib.ctr_oracle_murmur3_32.restype = ctypes.c_uint32
    for r in hash_rows[::7]:
        data = str(r["v"]).encode("utf-8")
        got = lib.ctr_oracle_murmur3_32(data, ctypes.c_int(len(data)), ctypes.c_uint32(r["seed"]))
        assert got == r["murmur"]
    rows = [r for r in hash_rows if r["is_int"] and r["seed"] == 7 and r["buckets"] == 1000003]
    ids = np.array([r["v"] for r in rows], dtype=np.int64)
    out = np.zeros(len(ids), dtype=np.int32)
    lib.ctr_oracle_hash_bucket_i64(ids.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(len(ids)),
                                   ctypes.c_uint32(1000003), ctypes.c_uint32(7), out.ctypes.data_as(ctypes.c_void_p))
    assert (out == np.array([r["bucket"] for r in rows])).all()
    # C pooling restatement == torch restatement
    g = torch.Generator().manual_seed(3)
    idt = torch.randint(-3, 20, (17, 6), generator=g)
    tab = torch.randn(20, 8, generator=g)
    for mode, name in ((0, "sum"), (1, "mean")):
        o = np.zeros((17, 8), dtype=np.float32)
        lib.ctr_oracle_pool(idt.numpy().ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(17), ctypes.c_int64(6),
                            tab.numpy().ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(8), ctypes.c_int(mode),
                            o.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(8))
        np.testing.assert_allclose(o, oe.pooled_lookup(idt, tab, name).numpy(), rtol=1e-6, atol=1e-6)


@pytest.fixture(scope="module")
def dnn_gold(golden_dir):
    return torch.load(os.path.join(golden_dir, "dnn_golden.pt"), weights_only=True)


def _oracle_dnn(gold, p_drop=0.0):
    m = om.OracleDNN(gold["feat_configs"], gold["hidden_units"])
    m.load_state_dict(gold["init_state"])
    for mod in m.tower:
        if isinstance(mod, torch.nn.Dropout):
            mod.p = p_drop
    return m


def test_oracle_dnn_forward_backward_step_match_reference(dnn_gold):
    m = _oracle_dnn(dnn_gold)
    m.eval()
    with torch.no_grad():
        torch.testing.assert_close(m(dnn_gold["feats"]), dnn_gold["eval_logits"], rtol=1e-6, atol=1e-6)
    m.train()
    opt = torch.optim.Adagrad(m.parameters(), lr=dnn_gold["adagrad_lr"])
    loss = m.training_step((dnn_gold["feats"], dnn_gold["labels"]), 0)
    loss.backward()
    torch.testing.assert_close(loss.detach(), dnn_gold["train_loss"], rtol=1e-6, atol=1e-7)
    for k, p in m.named_parameters():
        torch.testing.assert_close(p.grad, dnn_gold["grads"][k], rtol=1e-5, atol=1e-7)
    opt.step()
    for k, v in m.state_dict().items():
        torch.testing.assert_close(v, dnn_gold["after_adagrad_step"][k], rtol=1e-5, atol=1e-7)


def test_oracle_table_grad_functions_match_autograd(dnn_gold):
    feats, grads = dnn_gold["feats"], dnn_gold["grads"]
    m = _oracle_dnn(dnn_gold)
    m.train()
    pooled = [p.detach().requires_grad_(True) for p in m.pooled(feats)]
    x = torch.cat(pooled + [feats["dense_features"]], dim=-1)
    loss = torch.nn.functional.binary_cross_entropy_with_logits(m.tower(x), dnn_gold["labels"])
    loss.backward()
    for cfg, p in zip(dnn_gold["feat_configs"], pooled):
        name = cfg["name"]
        dense = oe.dense_table_grad(feats[name], p.grad, cfg["num_embeddings"])
        torch.testing.assert_close(dense, grads[f"embeddings.{name}.weight"], rtol=1e-5, atol=1e-8)
        rows, vals = oe.unique_row_grads(feats[name], p.grad, cfg["num_embeddings"])
        torch.testing.assert_close(vals, grads[f"embeddings.{name}.weight"][rows], rtol=1e-5, atol=1e-8)
        untouched = torch.ones(cfg["num_embeddings"], dtype=torch.bool)
        untouched[rows] = False
        # pads read row 0 but contribute an exactly-zero gradient (dnn.py:56-57)
        assert float(grads[f"embeddings.{name}.weight"][untouched].abs().sum()) == 0.0


def test_training_loop_protocol_matches_reference_trainer(dnn_gold):
    """zero_grad -> training_step -> backward -> step (trainer.py:291-303)."""
    m = _oracle_dnn(dnn_gold)
    m.train()
    opt = torch.optim.Adagrad(m.parameters(), lr=dnn_gold["adagrad_lr"])
    trace = []
    for k, batch in enumerate(dnn_gold["trainer_batches"]):
        opt.zero_grad()
        loss = m.training_step(batch, k)
        trace.append(loss.item())
        loss.backward()
        opt.step()
    torch.testing.assert_close(torch.tensor(trace), dnn_gold["trainer_loss_trace"], rtol=1e-5, atol=1e-6)
    for k, v in m.state_dict().items():
        torch.testing.assert_close(v, dnn_gold["trainer_final_state"][k], rtol=1e-4, atol=1e-6)


def test_dynamic_embedding_semantics(golden_dir):
    g = torch.load(os.path.join(golden_dir, "dynamic_golden.pt"), weights_only=True)
    # notebook-pinned shapes (examples/test_dynamic_embedding.ipynb:73,108,117)
    assert tuple(g["w_after"].shape) == (10, 5)
    assert tuple(g["loaded_emb1"].shape) == (10, 5) and tuple(g["loaded_emb2"].shape) == (15, 5)
    gen = torch.Generator().manual_seed(11)          # make_golden seeds the global RNG with 11
    torch.manual_seed(11)
    grown, fresh = oe.grow_table(g["w_before"], int(g["ids"].max()))
    torch.testing.assert_close(grown, g["w_after"], rtol=0, atol=0)
    assert float(fresh.abs().max()) < 0.1             # N(0, 0.01) rows
    torch.testing.assert_close(grown[g["ids"]], g["out"], rtol=0, atol=0)
    # state-dict merge: larger ckpt replaces, smaller ckpt padded with the module's tail rows
    torch.testing.assert_close(oe.merge_smaller_checkpoint(torch.zeros(8, 5), g["state_dict"]["emb1.weight"]),
                               g["loaded_emb1"], rtol=0, atol=0)
    cur = torch.cat([torch.zeros(6, 5), g["emb2_tail_before_load"]])
    torch.testing.assert_close(oe.merge_smaller_checkpoint(cur, g["state_dict"]["emb2.weight"]),
                               g["loaded_emb2"], rtol=0, atol=0)
    with pytest.raises(ValueError, match=g["errors"]["empty"]):
        oe.check_dynamic_ids(torch.zeros(0, dtype=torch.long))
    with pytest.raises(ValueError, match=g["errors"]["negative"]):
        oe.check_dynamic_ids(torch.tensor([[1, -1]]))


def test_row_optimizers_match_torch_optim(golden_dir):
    g = torch.load(os.path.join(golden_dir, "optim_golden.pt"), weights_only=True)
    V, D = g["w0"].shape

    def run(kind):
        w = g["w0"].clone()
        s1, s2 = torch.zeros(V, D), torch.zeros(V, D)
        for step, (ids, gout) in enumerate(g["steps"], start=1):
            rows, vals = oe.unique_row_grads(ids, gout, V)
            if kind == "adagrad":
                oo.adagrad_rows(w, s1, rows, vals, g[kind]["lr"], g[kind]["eps"])
            elif kind == "sparse_adam":
                oo.sparse_adam_rows(w, s1, s2, rows, vals, g[kind]["lr"], step, *g[kind]["betas"], g[kind]["eps"])
            else:
                oo.sgd_rows(w, rows, vals, g[kind]["lr"])
        return w, s1, s2

    w, s1, _ = run("adagrad")
    torch.testing.assert_close(w, g["adagrad"]["w"], rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(s1, g["adagrad"]["sum"], rtol=1e-5, atol=1e-7)
    w, s1, s2 = run("sparse_adam")
    torch.testing.assert_close(w, g["sparse_adam"]["w"], rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(s1, g["sparse_adam"]["exp_avg"], rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(s2, g["sparse_adam"]["exp_avg_sq"], rtol=1e-5, atol=1e-9)
    w, _, _ = run("sgd")
    torch.testing.assert_close(w, g["sgd"]["w"], rtol=1e-5, atol=1e-7)


def test_vocab_growth_rules():
    v = Vocab(min_freq=2)
    assert v.fit([7, 7, 3, 9, 9, 9, 3, 5]) == 4            # 5 appears once -> not admitted
    assert v.idx_of == {7: 1, 3: 2, 9: 3}                   # first-occurrence order, idx from 1
    assert v.transform([[7, 5], [9, 100]]).tolist() == [[1, 0], [3, 0]]   # unknown -> OOV idx 0
    assert v.fit([5, 5, 7, 11]) == 5                        # old keys keep idx, 11 below min_freq
    assert v.idx_of[5] == 4 and v.idx_of[7] == 1 and 11 not in v.idx_of
    assert v.cnt_of[7] == 2 and v.cnt_of[9] == 3            # 7 seen once in batch 2 -> filtered before cnt update
    v0 = Vocab()
    assert v0.fit([4, 4, 2]) == 3 and v0.idx_of == {4: 1, 2: 2}


def test_fm_identity():
    g = torch.Generator().manual_seed(0)
    v = torch.randn(5, 7, 4, generator=g)
    brute = torch.zeros(5, 1)
    for i in range(7):
        for j in range(i + 1, 7):
            brute += (v[:, i] * v[:, j]).sum(-1, keepdim=True)
    torch.testing.assert_close(om.fm_second_order(v), brute, rtol=1e-5, atol=1e-5)
