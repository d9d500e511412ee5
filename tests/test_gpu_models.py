"""GPU parity tests at module level: DNN against the golden fixtures recorded from the real
reference (tests/golden/dnn_golden.pt), DeepFM / DCN-v2 / FM against the oracle definitions,
DynamicEmbedding against dynamic_golden.pt, the device vocabulary against oracle.vocab.

Embedding-path tolerances are the north_star's 1e-5 (see test_gpu_lookup.py); quantities that
pass through the fp32 cuBLAS tower GEMMs are compared at 1e-4 (GPU vs CPU GEMM summation order).
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOWER_RTOL = 1e-4


def close(got, ref, rtol):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    assert got.shape == ref.shape, (got.shape, ref.shape)
    scale = max(float(ref.abs().max()), 1e-30)
    err = float((got - ref).abs().max())
    assert err <= rtol * scale, f"max abs err {err:.3e} > {rtol} * {scale:.3e}"


def no_dropout(model):
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    return model


@pytest.fixture(autouse=True)
def _fp32_matmul():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield


def test_dnn_matches_reference_golden(golden_dir):
    from torchctr_b200.models import DNN
    g = torch.load(os.path.join(golden_dir, "dnn_golden.pt"))
    model = no_dropout(DNN(g["feat_configs"], g["hidden_units"])).cuda()
    model.load_state_dict(g["init_state"])                   # reference checkpoint loads unchanged
    model.eval()
    with torch.no_grad():
        close(model(g["feats"]), g["eval_logits"], TOWER_RTOL)
    # one Adagrad step: dense params through torch.optim, tables through the fused update
    model.train()
    opt = torch.optim.Adagrad(model.parameters(), lr=g["adagrad_lr"])
    model.bind_optimizer(opt)
    opt.zero_grad()
    loss = model.training_step((g["feats"], g["labels"]), 0)
    close(loss, g["train_loss"], TOWER_RTOL)
    loss.backward()
    # a Linear bias in front of BatchNorm has a mathematically zero gradient: what the fixture holds there is
    # rounding noise, which the first Adagrad step turns into +-lr -- such parameters are not comparable
    noise = {n for n, gr in g["grads"].items() if float(gr.abs().max()) < 1e-6}
    for name, p in model.named_parameters():
        if name.startswith("tower"):
            if name not in noise:
                close(p.grad, g["grads"][name], 5e-4)
        else:
            assert p.grad is None                            # no dense [V, D] gradient exists
    opt.step()
    after = model.state_dict()
    for k, ref in g["after_adagrad_step"].items():
        if k not in noise:
            close(after[k], ref, 5e-4 if k.startswith("tower") else 2e-4)


def test_dnn_sparse_grads_without_binding(golden_dir):
    """No optimizer bound: backward hands autograd sparse gradients equal to the reference's dense ones."""
    from torchctr_b200.models import DNN
    g = torch.load(os.path.join(golden_dir, "dnn_golden.pt"))
    model = no_dropout(DNN(g["feat_configs"], g["hidden_units"])).cuda()
    model.load_state_dict(g["init_state"])
    model.train()
    model.training_step((g["feats"], g["labels"]), 0).backward()
    for name, p in model.named_parameters():
        if name.startswith("embeddings"):
            assert p.grad.is_sparse
            close(p.grad.to_dense(), g["grads"][name], 2e-4)


def test_trainer_loop_trace_matches_reference(golden_dir):
    """Six steps of the reference Trainer.fit (Adagrad) recorded in the fixture; here the same loop
    (zero_grad / training_step / backward / step, trainer.py:291-303) over our model."""
    from torchctr_b200.models import DNN
    g = torch.load(os.path.join(golden_dir, "dnn_golden.pt"))
    model = no_dropout(DNN(g["feat_configs"], g["hidden_units"])).cuda()
    model.load_state_dict(g["init_state"])
    opt = torch.optim.Adagrad(model.parameters(), lr=0.05)
    model.bind_optimizer(opt)
    model.train()
    trace = []
    for k, batch in enumerate(g["trainer_batches"]):
        opt.zero_grad()
        loss = model.training_step(batch, k)
        trace.append(loss.item())
        loss.backward()
        opt.step()
    close(torch.tensor(trace), g["trainer_loss_trace"], 2e-4)
    final = model.state_dict()
    for k, ref in g["trainer_final_state"].items():
        if k.startswith("embeddings"):
            close(final[k], ref, 1e-3)


def _criteo_like(gen, B, F, V, D, nd):
    fc = [{"name": f"c{i}", "type": "sparse", "num_embeddings": V + i, "emb_dim": D} for i in range(F)]
    fc += [{"name": f"d{i}", "type": "dense"} for i in range(nd)]
    feats = {f"c{i}": torch.randint(0, V + i, (B, 1), generator=gen) for i in range(F)}
    feats["dense_features"] = torch.randn(B, nd, generator=gen)
    labels = (torch.rand(B, 1, generator=gen) < 0.25).float()
    return fc, feats, labels


@pytest.mark.parametrize("which", ["deepfm", "dcnv2"])
def test_deepfm_dcn_match_oracle(which):
    from oracle import models as om
    from torchctr_b200.models import DCNv2, DeepFM
    gen = torch.Generator().manual_seed(8)
    D = 16 if which == "deepfm" else 32
    fc, feats, labels = _criteo_like(gen, 512, 6, 40, D, 3)
    torch.manual_seed(0)
    ref = no_dropout(om.OracleDeepFM(fc, [32, 16]) if which == "deepfm" else om.OracleDCNv2(fc, [32, 16], 3))
    ours = no_dropout(DeepFM(fc, [32, 16]) if which == "deepfm" else DCNv2(fc, [32, 16], 3)).cuda()
    ours.load_state_dict(ref.state_dict())
    ref.train(); ours.train()
    lr = 0.1
    opt_ref = torch.optim.SGD(ref.parameters(), lr=lr)
    opt = torch.optim.SGD(ours.parameters(), lr=lr)
    ours.bind_optimizer(opt)
    for step in range(2):
        opt_ref.zero_grad(); opt.zero_grad()
        l_ref = ref.training_step((feats, labels), step)
        l = ours.training_step((feats, labels), step)
        close(l, l_ref, TOWER_RTOL)
        l_ref.backward(); l.backward()
        opt_ref.step(); opt.step()
    sd, sd_ref = ours.state_dict(), ref.state_dict()
    for k in sd_ref:
        close(sd[k], sd_ref[k], 5e-4)


def test_fm_kernels_match_autograd():
    from oracle.models import fm_second_order
    from torchctr_b200.nn import fm_interaction
    gen = torch.Generator().manual_seed(1)
    for F, D, stride, nfirst in ((26, 16, 432, 26), (5, 9, 48, 0), (3, 32, 96, 3), (7, 4, 28, 7)):
        B = 777
        x = torch.zeros(B, stride)
        x[:, :F * D] = torch.randn(B, F * D, generator=gen)
        first = torch.randn(B, (nfirst + 3) // 4 * 4, generator=gen) if nfirst else None
        xr = x.clone().requires_grad_(True)
        fr = first.clone().requires_grad_(True) if nfirst else None
        ref = fm_second_order(xr[:, :F * D].reshape(B, F, D))
        if nfirst:
            ref = ref + fr[:, :nfirst].sum(1, keepdim=True)
        gout = torch.randn(B, 1, generator=gen)
        ref.backward(gout)
        xg = x.cuda().requires_grad_(True)
        fg = first.cuda().requires_grad_(True) if nfirst else None
        out = fm_interaction(xg, F, D, fg, nfirst)
        out.backward(gout.cuda())
        close(out, ref, 1e-5)
        close(xg.grad, xr.grad, 1e-5)
        if nfirst:
            close(fg.grad, fr.grad, 1e-5)


def test_dynamic_embedding_matches_reference_golden(golden_dir):
    from torchctr_b200.nn import DynamicEmbedding
    g = torch.load(os.path.join(golden_dir, "dynamic_golden.pt"))

    class Two(torch.nn.Module):
        def __init__(self, n1, n2):
            super().__init__()
            self.emb1 = DynamicEmbedding(n1, 5)
            self.emb2 = DynamicEmbedding(n2, 5)

    m = Two(5, 6).cuda()
    with torch.no_grad():
        m.emb1.weight.copy_(g["w_before"])
    out = m.emb1(g["ids"].cuda())
    assert tuple(m.emb1.weight.shape) == tuple(g["w_after"].shape) == (10, 5)      # grew to max id + 1
    assert tuple(m.emb2.weight.shape) == (6, 5)
    assert torch.equal(m.emb1.weight[:5].cpu(), g["w_before"])                    # old rows untouched
    new = m.emb1.weight[5:].detach().cpu()
    assert float(new.abs().max()) < 0.08 and float(new.std()) > 0.002               # ~ N(0, 0.01)
    assert torch.equal(out.detach().cpu(), m.emb1.weight.detach().cpu()[g["ids"]])
    assert out.shape == g["out"].shape
    # state-dict merge: a smaller table grows to the checkpoint, a larger one keeps its tail rows
    bigger = Two(8, 15).cuda()
    tail = bigger.emb2.weight.detach()[6:].clone().cpu()
    bigger.load_state_dict({k: v.clone() for k, v in g["state_dict"].items()})
    assert tuple(bigger.emb1.weight.shape) == tuple(g["loaded_emb1"].shape)
    assert tuple(bigger.emb2.weight.shape) == tuple(g["loaded_emb2"].shape)
    assert torch.equal(bigger.emb1.weight.detach().cpu(), g["state_dict"]["emb1.weight"])
    assert torch.equal(bigger.emb2.weight.detach().cpu()[:6], g["state_dict"]["emb2.weight"])
    assert torch.equal(bigger.emb2.weight.detach().cpu()[6:], tail)
    for name, bad in (("empty", torch.zeros(0, dtype=torch.long)), ("negative", torch.tensor([[1, -1]]))):
        with pytest.raises(ValueError) as e:
            m.emb1(bad.cuda())
        assert str(e.value) == g["errors"][name]
    # gradient flows through the gather as a sparse update
    m.emb1(torch.tensor([[1, 1, 3]]).cuda()).sum().backward()
    dense = m.emb1.weight.grad.to_dense().cpu()
    assert torch.equal(dense[1], torch.full((5,), 2.0)) and torch.equal(dense[3], torch.ones(5)) and float(dense[0].abs().sum()) == 0


def test_vocab_fit_transform_match_oracle():
    from oracle.vocab import Vocab
    from torchctr_b200.nn import VocabIndex
    rng = np.random.default_rng(5)
    for min_freq in (0, 2):
        ref = Vocab(min_freq=min_freq)
        dev = VocabIndex(capacity=256, min_freq=min_freq).cuda()      # small: forces a rebuild
        for step in range(5):
            keys = rng.zipf(1.3, size=(400, 5)).astype(np.int64) * 7919 + step
            keys[rng.random(keys.shape) < 0.2] = -100
            flat = keys[keys >= 0]
            n_ref = ref.fit(flat)
            dev.fit(torch.from_numpy(keys))
            assert dev.num_embeddings() == n_ref
            probe = np.concatenate([flat[:500], rng.integers(0, 10 ** 9, 100)])
            got = dev.transform(torch.from_numpy(probe)).cpu().numpy()
            assert (got == ref.transform(probe)).all()                # bit-exact rows, incl. first-occurrence order


def test_model_with_growing_vocab_and_hashed_feature():
    """raw ids end to end: a hashed feature and a vocabulary feature that grows while training."""
    from oracle import embedding as oe
    from oracle import hashing as oh
    from oracle.vocab import Vocab
    from torchctr_b200.models import DNN
    fc = [{"name": "h", "type": "sparse", "num_embeddings": 101, "emb_dim": 16, "raw_ids": True, "hash_buckets": 101, "seed": 3},
          {"name": "v", "type": "sparse", "num_embeddings": 1, "emb_dim": 16, "raw_ids": True, "islist": True},
          {"name": "x", "type": "dense"}]
    torch.manual_seed(0)
    model = no_dropout(DNN(fc, [16])).cuda().train()
    gen = torch.Generator().manual_seed(3)
    ref_vocab = Vocab()
    for step in range(3):
        feats = {"h": torch.randint(0, 2 ** 33, (64, 1), generator=gen),
                 "v": torch.randint(1000 * step, 1000 * step + 300, (64, 7), generator=gen),
                 "dense_features": torch.randn(64, 1, generator=gen)}
        feats["v"][torch.rand(64, 7, generator=gen) < 0.3] = -100
        logits = model(feats)
        n = ref_vocab.fit(feats["v"][feats["v"] >= 0].numpy())
        assert model.embeddings["v"].num_embeddings == n == model.embeddings["v"].weight.shape[0]
        # recompute the tower input on the CPU from the oracle's rows
        hrows = torch.from_numpy(oh.hash_bucket_ids(feats["h"].numpy(), 101, 3)).long()
        vrows = torch.from_numpy(ref_vocab.transform(feats["v"].numpy())).long()
        vrows[feats["v"] < 0] = -100
        x = torch.cat([oe.pooled_lookup(hrows, model.embeddings["h"].weight.detach().cpu()),
                       oe.pooled_lookup(vrows, model.embeddings["v"].weight.detach().cpu()), feats["dense_features"]], 1)
        with torch.no_grad():
            ref_logits = torch.nn.functional.linear(x.cuda(), model.tower[0].weight, model.tower[0].bias)
            got_first = model._first_linear(model._lookup(feats, feats["dense_features"], False), model.tower[0])
        close(got_first, ref_logits, TOWER_RTOL)
        logits.sum().backward()


@pytest.mark.parametrize("opt_name", ["adagrad", "adam"])
def test_graphed_train_step_equals_eager(opt_name):
    """Whole-step CUDA graph replay == eager launches (same kernels), incl. Adam's step-dependent
    scalars, which reach the kernels through the device hyper-parameter tensor."""
    from torchctr_b200.graph import GraphedTrainStep
    from torchctr_b200.models import DeepFM
    gen = torch.Generator().manual_seed(17)
    fc, f0, l0 = _criteo_like(gen, 1024, 5, 300, 16, 3)
    batches = [(f0, l0)]
    for _ in range(5):
        feats = {k: (torch.randint(0, 300, v.shape, generator=gen) if v.dtype == torch.int64
                     else torch.randn(v.shape, generator=gen)) for k, v in f0.items()}
        batches.append((feats, (torch.rand(1024, 1, generator=gen) < 0.25).float()))
    torch.manual_seed(1)
    a = no_dropout(DeepFM(fc, [32, 16])).cuda().train()
    b = no_dropout(DeepFM(fc, [32, 16])).cuda().train()
    b.load_state_dict(a.state_dict())

    def make_opt(m):
        if opt_name == "adagrad":
            o = torch.optim.Adagrad(m.dense_parameters(), lr=0.05)
            m.bind_optimizer(o, kind="adagrad")
        else:
            o = torch.optim.Adam(m.dense_parameters(), lr=0.01, capturable=True)
            m.bind_optimizer(o, kind="adam")
        return o

    oa, ob = make_opt(a), make_opt(b)
    # the constructor runs `warmup` real steps on the example batch before capturing
    graphed = GraphedTrainStep(b, ob, batches[0], warmup=1)
    oa.zero_grad(); a.training_step(batches[0], 0).backward(); oa.step()
    pinned = [({k: v.pin_memory() for k, v in f.items()}, l.pin_memory()) for f, l in batches]
    for i, batch in enumerate(batches):
        oa.zero_grad()
        la = a.training_step(batch, i)
        la.backward()
        oa.step()
        if i == 1:                                   # the i32 wire block (ids narrowed on the host, widened on the device)
            lb = graphed(graphed.pack(pinned[i], "cpu", ids="i32"))
        elif i == 2:                                 # the input-prefetch path: copy on a side stream, then one D2D move
            graphed.prefetch(pinned[i])
            lb = graphed(pinned[i])
        elif i == 3:                                 # prefetch of a wire block
            wire = graphed.pack(pinned[i], "cpu", ids="i32")
            assert wire.numel() < graphed.pack(pinned[i], "cpu").numel()
            graphed.prefetch(wire)
            lb = graphed(wire)
        elif i == 4:                                 # full-width packed block, resident on the device
            lb = graphed(graphed.pack(batch))
        elif i == 5:                                 # wire block resident on the device
            lb = graphed(graphed.pack(batch, ids="i32"))
        else:
            lb = graphed(batch)
        close(lb, la, 1e-5)
    sa, sb = a.state_dict(), b.state_dict()
    for k in sa:
        if sa[k].dtype.is_floating_point:
            close(sb[k], sa[k], 1e-5)
    big = ({k: (v + 2 ** 31 if v.dtype == torch.int64 else v) for k, v in batches[0][0].items()}, batches[0][1])
    with pytest.raises(ValueError):                  # ids that do not fit 32 bits cannot take the wire layout
        graphed.pack(big, "cpu", ids="i32")


@pytest.mark.parametrize("pooling", ["sum", "mean"])
def test_cfg5_long_sequences_with_growing_vocab(pooling):
    """BASELINE config 5 on one GPU: a variable-length sequence feature (up to 200 ids / row, Zipf keys from an
    unbounded key space, fresh keys every step) pooled through a vocabulary that grows while training, fused
    Adagrad on the touched rows.  Rows are bit-exact against the oracle vocabulary, pooled vectors / updated rows 1e-5."""
    from oracle import embedding as oe
    from oracle.vocab import Vocab
    from torchctr_b200.models import DNN
    B, L, D = 2048, 200, 16
    fc = [{"name": "seq", "type": "sparse", "num_embeddings": 1, "emb_dim": D, "raw_ids": True, "islist": True,
           "pooling": pooling, "vocab_capacity": 1 << 12},
          {"name": "x", "type": "dense"}]
    torch.manual_seed(0)
    model = no_dropout(DNN(fc, [16])).cuda().train()
    opt = torch.optim.Adagrad(model.dense_parameters(), lr=0.05)
    model.bind_optimizer(opt, kind="adagrad")
    captured = {}
    run_tower = model._run_tower

    def spy(x, *a, **k):                                   # the tower input and its gradient
        captured["x"] = x.detach().clone()
        x.register_hook(lambda g: captured.__setitem__("gx", g.detach().clone()))
        return run_tower(x, *a, **k)
    model._run_tower = spy
    rng = np.random.default_rng(11)
    gen = torch.Generator().manual_seed(5)
    ref_vocab = Vocab()
    table = model.embeddings["seq"]
    table_ref = table.weight.detach().cpu().clone()
    acc_ref = torch.zeros_like(table_ref)
    for step in range(3):
        keys = (rng.zipf(1.05, size=(B, L)) % (10 ** 7)).astype(np.int64) * 31 + step * 10 ** 9     # fresh key range per step
        lens = rng.integers(1, L + 1, size=B)
        keys[np.arange(L)[None, :] >= lens[:, None]] = -100
        feats = {"seq": torch.from_numpy(keys), "dense_features": torch.randn(B, 1, generator=gen)}
        labels = (torch.rand(B, 1, generator=gen) < 0.25).float()
        n = ref_vocab.fit(keys[keys >= 0])
        rows = torch.from_numpy(ref_vocab.transform(keys)).long()
        rows[torch.from_numpy(keys) < 0] = -100
        opt.zero_grad()
        loss = model.training_step((feats, labels), step)  # forward: grows the vocabulary and the table, then pools
        assert table.num_embeddings == n == table.weight.shape[0]
        if table_ref.shape[0] < n:                         # grown rows are drawn by the kernel (N(0, 0.01)): adopt them
            grown = table.weight.detach().cpu()[table_ref.shape[0]:n]
            assert 0.005 < float(grown.std()) < 0.02
            table_ref = torch.cat([table_ref, grown])
            acc_ref = torch.cat([acc_ref, torch.zeros_like(grown)])
        close(captured["x"][:, :D], oe.pooled_lookup(rows, table_ref, pooling), 1e-5)
        loss.backward()                                    # fused Adagrad on the touched rows
        opt.step()
        urows, grads = oe.unique_row_grads(rows, captured["gx"][:, :D].cpu(), n, pooling)
        acc_ref[urows] += grads * grads
        table_ref[urows] -= 0.05 * grads / (acc_ref[urows].sqrt() + 1e-10)
        # Adagrad's first update of a row is lr * g / (|g| + eps): with mean pooling over up to 200 ids the gradients
        # are ~1e-8, a few eps (1e-10), and the ratio amplifies their 1e-6 relative error -- a conditioning of the
        # optimizer, not of the kernel (sum pooling holds 1e-5)
        close(table.weight.detach().cpu(), table_ref, 1e-5 if pooling == "sum" else 3e-4)
        close(table.opt_state0.detach().cpu()[:n], acc_ref, 5e-5)      # sums of squares: twice the relative error of the gradients


def test_model_on_a_device_that_is_not_current():
    """every launch runs under the device (and on the stream) of its tensors, not of torch's current device."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    from oracle import models as om
    from torchctr_b200.models import DeepFM
    gen = torch.Generator().manual_seed(2)
    fc, feats, labels = _criteo_like(gen, 256, 4, 50, 16, 3)
    torch.manual_seed(0)
    ref = no_dropout(om.OracleDeepFM(fc, [32, 16])).train()
    ours = no_dropout(DeepFM(fc, [32, 16])).to("cuda:1").train()
    ours.load_state_dict(ref.state_dict())
    assert torch.cuda.current_device() == 0
    opt_ref = torch.optim.SGD(ref.parameters(), lr=0.1)
    opt = torch.optim.SGD(ours.parameters(), lr=0.1)
    ours.bind_optimizer(opt)
    l_ref = ref.training_step((feats, labels), 0); l = ours.training_step((feats, labels), 0)
    close(l, l_ref, TOWER_RTOL)
    l_ref.backward(); l.backward(); opt_ref.step(); opt.step()
    torch.cuda.synchronize("cuda:1")
    for k, v in ref.state_dict().items():
        close(ours.state_dict()[k], v, 5e-4)


@pytest.mark.parametrize("which", ["deepfm", "dnn"])
def test_blocked_gradient_handover_equals_row_major(which):
    """The tower's first block hands dL/dx to the lookup column-blocked (nn.embedding.BlockedGrad: ctr_linear_fwd_blocked ->
    ctr_group_t.grad_blocked).  Same numbers in another layout: parameters after three steps are bit-identical to the
    row-major handover, and the blocked path is the one that ran."""
    from torchctr_b200.models import DNN, DeepFM
    from torchctr_b200.nn import embedding
    gen = torch.Generator().manual_seed(23)
    B = 1024
    fc, f0, l0 = _criteo_like(gen, B, 6, 500, 16, 3)
    batches = [(f0, l0)]
    for _ in range(2):
        feats = {k: (torch.randint(0, 500, v.shape, generator=gen) if v.dtype == torch.int64
                     else torch.randn(v.shape, generator=gen)) for k, v in f0.items()}
        batches.append((feats, (torch.rand(B, 1, generator=gen) < 0.25).float()))
    batches = [({k: v.cuda() for k, v in f.items()}, l.cuda()) for f, l in batches]
    old = embedding.BLOCKED_GRAD
    states = []
    try:
        for on in (False, True):
            embedding.BLOCKED_GRAD = on
            torch.manual_seed(4)
            m = (DeepFM if which == "deepfm" else DNN)(fc, [64, 32]).cuda().train()
            opt = torch.optim.Adagrad(m.dense_parameters(), lr=0.05)
            m.bind_optimizer(opt, kind="adagrad")
            n0 = embedding.blocked_backwards
            for i, batch in enumerate(batches):
                opt.zero_grad(set_to_none=True)
                m.training_step(batch, i).backward()
                opt.step()
            torch.cuda.synchronize()
            assert embedding.blocked_backwards - n0 == (len(batches) if on else 0)
            states.append({k: v.clone() for k, v in m.state_dict().items()})
        for k in states[0]:
            assert torch.equal(states[0][k], states[1][k]), k
    finally:
        embedding.BLOCKED_GRAD = old
