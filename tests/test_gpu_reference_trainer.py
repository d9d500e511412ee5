"""The drop-in, proven with the reference's OWN Trainer (missing #4 / next #7).

``torchctr.trainer.Trainer`` (unmodified, imported from ``baseline/_ref``) drives ``torchctr_b200.models.DNN`` exactly as
it drives ``torchctr.models.DNN``: same constructor call, the optimizer handed to the Trainer only (no ``bind_optimizer``:
the tables find their optimizer themselves), ``fit`` with per-epoch evaluation, ``save_ckpt`` / ``load_ckpt``
(``torch.load(weights_only=True)``, strict ``load_state_dict``).  Checked against the reference model trained by the same
Trainer on the CPU: loss traces, and checkpoints moving in BOTH directions -- ours into the reference model + torch.optim
and back -- with the next training step identical, which also proves that the fused Adagrad state of the tables is
checkpointed (it lives in ``optimizer.state_dict()`` in torch's own format).
Tower GEMMs run in the exact (3xTF32) mode, as in every parity test.
"""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

LR = 0.05


@pytest.fixture()
def ref():
    from baseline import refshim
    if not refshim.available():
        pytest.skip("baseline/_ref (the unmodified reference) is not installed: python baseline/install_ref.py")
    torch.backends.cuda.matmul.allow_tf32 = False
    return refshim.load_reference()


def _feat_configs():
    fc = [{"name": f"c{i}", "type": "sparse", "num_embeddings": 40 + 13 * i, "emb_dim": 16} for i in range(8)]
    fc.append({"name": "hist", "type": "sparse", "num_embeddings": 97, "emb_dim": 16, "islist": True})
    fc += [{"name": f"d{i}", "type": "dense"} for i in range(4)]
    return fc


def _batches(gen, n, B, fc):
    out = []
    for _ in range(n):
        feats = {}
        for c in fc:
            if c["type"] != "sparse":
                continue
            if c.get("islist"):
                ids = torch.randint(0, c["num_embeddings"], (B, 12), generator=gen)
                lens = torch.randint(0, 13, (B, 1), generator=gen)
                ids[torch.arange(12)[None, :] >= lens] = -100                  # the collate's padding value (dataset.py:9)
                feats[c["name"]] = ids
            else:
                feats[c["name"]] = torch.randint(0, c["num_embeddings"], (B, 1), generator=gen)
        feats["dense_features"] = torch.randn(B, 4, generator=gen)
        out.append((feats, (torch.rand(B, 1, generator=gen) < 0.3).float()))
    return out


def _no_dropout(m):
    for x in m.modules():
        if isinstance(x, torch.nn.Dropout):
            x.p = 0.0
    return m


def _pair(ref, fc):
    from torchctr_b200.models import DNN
    torch.manual_seed(0)
    theirs = _no_dropout(ref.models.DNN(fc, [32, 16]))
    ours = _no_dropout(DNN(fc, [32, 16])).cuda()
    ours.load_state_dict(theirs.state_dict())              # strict: identical keys
    return theirs, ours


def _trainer(ref, model, path, epochs, trace):
    opt = torch.optim.Adagrad(model.parameters(), lr=LR)                       # the whole model, tables included
    return ref.trainer.Trainer(model, optimizer=opt, max_epochs=epochs, save_ckpt_path=str(path), use_accelerate=False,
                               log_steps=10 ** 6, callback_train_epoch_end=lambda rets: trace.extend(rets),
                               callback_eval_epoch_end=lambda rets: trace.extend(rets))


def _compare_traces(ours, theirs, n_train, n_eval):
    """Training losses to 2e-4; evaluation losses to 5e-3.  A Linear bias in front of BatchNorm has a mathematically zero
    gradient; what arrives is rounding noise, and Adagrad turns noise of ANY size into steps of +-lr.  In training mode
    BatchNorm subtracts the bias again (losses agree to 1e-7), but running_mean lags behind a bias that jitters by +-0.05
    per step, so evaluation-mode outputs of the reference ITSELF depend on that noise at the 1e-3 level."""
    per_epoch = n_train + n_eval
    for i, (a, b) in enumerate(zip(ours, theirs)):
        tol = 2e-4 if i % per_epoch < n_train else 5e-3
        assert abs(a - b) <= tol * max(1.0, abs(b)), (i, ours, theirs)


def test_reference_trainer_fit_and_checkpoints_both_ways(ref, tmp_path):
    fc = _feat_configs()
    gen = torch.Generator().manual_seed(3)
    train, evalb = _batches(gen, 6, 512, fc), _batches(gen, 2, 512, fc)
    theirs, ours = _pair(ref, fc)
    t_trace, o_trace = [], []
    t_tr = _trainer(ref, theirs, tmp_path / "theirs", 2, t_trace)
    o_tr = _trainer(ref, ours, tmp_path / "ours", 2, o_trace)
    t_tr.fit(train, evalb)
    o_tr.fit(train, evalb)                                                     # unmodified Trainer.fit over our model
    assert ours._lookup.binding is not None and ours._lookup.binding.kind == "adagrad"      # found its optimizer
    assert len(o_trace) == len(t_trace) == 2 * (6 + 2)
    _compare_traces(o_trace, t_trace, 6, 2)
    # tables never got a dense gradient, and moved only where the batch touched them
    assert all(p.grad is None for n, p in ours.named_parameters() if n.startswith("embeddings"))
    sd_o, sd_t = ours.state_dict(), theirs.state_dict()
    assert list(sd_o) == list(sd_t)                                            # the reference's keys, nothing else
    for k in sd_t:
        if sd_t[k].dtype.is_floating_point and not k.endswith(("0.bias", "4.bias", "running_mean")):   # see _compare_traces
            # 12 Adagrad steps from N(0, 1) tables: Adagrad's g / sqrt(sum g^2) turns round-off sized gradients into steps of
            # order lr, so two fp32 implementations drift apart by ~1e-3 (measured: 1e-6 after 3 steps, 5e-4 after 12) while
            # the losses above stay within 2e-4; the parameters are only checked for gross errors here
            err = (sd_o[k].cpu() - sd_t[k]).abs()
            scale = max(1.0, float(sd_t[k].abs().max()))
            assert float(err.mean()) <= 1e-3 * scale and float(err.max()) <= 5e-2 * scale, (k, float(err.max()))

    # ---- our checkpoint -> the reference model + torch.optim.Adagrad, through Trainer.load_ckpt (weights_only, strict)
    ckpt = torch.load(os.path.join(str(tmp_path / "ours"), "checkpoint.000012.ckpt"), weights_only=True)
    assert ckpt["model.feat_configs"] == fc and ckpt["global_steps"] == 12
    # both directions: each checkpoint directory is read by a fresh reference model AND a fresh model of ours through the
    # unmodified Trainer.load_ckpt; from the identical restored state one more epoch must give the same losses
    # (exclude_keys: the reference's load_ckpt calls None.load_state_dict when no lr_scheduler was configured, trainer.py:476-480)
    for src in ("ours", "theirs"):
        theirs2, ours2 = _pair(ref, fc)
        t2_trace, o2_trace = [], []
        t2 = _trainer(ref, theirs2, tmp_path / src, 3, t2_trace)
        o2 = _trainer(ref, ours2, tmp_path / src, 3, o2_trace)
        t2.load_ckpt(str(tmp_path / src), exclude_keys=["lr_scheduler"])
        o2.load_ckpt(str(tmp_path / src), exclude_keys=["lr_scheduler"])
        assert t2.num_epoch == o2.num_epoch == 2 and t2.global_steps == o2.global_steps == 12
        for k, v in theirs2.state_dict().items():                              # restored bit-for-bit on both sides
            assert torch.equal(v, ours2.state_dict()[k].cpu()), (src, k)
        # the Adagrad sums of the tables arrived too (torch's own state format, whoever wrote the checkpoint)
        s_theirs = t2.optimizer.state[theirs2.embeddings["c0"].weight]["sum"]
        s_ours = o2.optimizer.state[ours2.embeddings["c0"].weight]["sum"]
        assert float(s_theirs.abs().max()) > 0 and torch.equal(s_theirs, s_ours.cpu()), src
        t2.max_epochs = o2.max_epochs = 3                                      # load_ckpt also restored max_epochs = 2
        t2.fit(train, evalb)
        o2.fit(train, evalb)
        assert len(o2_trace) == len(t2_trace) == 8
        _compare_traces(o2_trace, t2_trace, 6, 2)


def test_tables_outside_the_torch_optimizer_checkpoint_separately(ref, tmp_path):
    """The bench's arrangement (dense parameters in the torch optimizer, tables fused): model.state_dict() still has only the
    reference's keys, and the fused state travels through table_optimizer_state_dict()."""
    fc = _feat_configs()
    gen = torch.Generator().manual_seed(5)
    train = _batches(gen, 3, 256, fc)
    _, ours = _pair(ref, fc)
    opt = torch.optim.Adagrad(ours.dense_parameters(), lr=LR)
    ours.bind_optimizer(opt, kind="adagrad")
    ours.train()
    for k, b in enumerate(train):
        opt.zero_grad(); ours.training_step(b, k).backward(); opt.step()
    theirs = ref.models.DNN(fc, [32, 16])
    theirs.load_state_dict({k: v.cpu() for k, v in ours.state_dict().items()})            # strict, unmodified reference
    extra = ours.table_optimizer_state_dict()
    path = tmp_path / "extra.pt"
    torch.save(extra, path)
    back = torch.load(path, weights_only=True)
    assert set(back) == {f"embeddings.{c['name']}" for c in fc if c["type"] == "sparse"}
    _, again = _pair(ref, fc)
    again.load_state_dict(ours.state_dict())
    opt2 = torch.optim.Adagrad(again.dense_parameters(), lr=LR)
    opt2.load_state_dict(opt.state_dict())
    again.bind_optimizer(opt2, kind="adagrad")
    again.load_table_optimizer_state_dict(back)
    again.train()
    la = again.training_step(train[0], 0); lo = ours.training_step(train[0], 0)
    opt.zero_grad(); opt2.zero_grad(); la.backward(); lo.backward(); opt.step(); opt2.step()
    torch.cuda.synchronize()
    assert abs(la.item() - lo.item()) <= 1e-6 * max(1.0, abs(lo.item()))
    for k, v in ours.state_dict().items():
        if k.startswith("embeddings"):                      # the fused Adagrad continued from the restored sums
            assert float((v - again.state_dict()[k]).abs().max()) <= 1e-5 * float(v.abs().max()), k
