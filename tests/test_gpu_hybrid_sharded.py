"""Hybrid placement (torchctr_b200.parallel.hybrid): small tables replicated (dense gradient all-reduce + dense update),
large tables row-sharded over peer memory, DeepFM's first-order tables and FM term fused into the one lookup and the
sweeps.  Ranks run as THREADS of one process on one GPU (ThreadTransport); the oracle is plain torch autograd on the CPU
over the full tables and the GLOBAL batch (what the reference's replicas + DDP compute, torchctr/trainer.py:128-130)."""
import threading

import pytest
import torch

pytestmark = pytest.mark.gpu

RTOL = 1e-5


def _close(got, ref, rtol=RTOL):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    if got.shape != ref.shape or ref.numel() == 0:
        return got.shape == ref.shape
    return float((got - ref).abs().max()) <= rtol * max(float(ref.abs().max()), 1e-30)


def _problem(world, D, seed=11, hashed=False):
    gen = torch.Generator().manual_seed(seed)
    Vs = [37, 3000, 2, 1000, 5000, 101, 64]            # replicate_max_rows = 200 -> tables 1, 3, 4 are sharded
    B = 257
    full = [torch.randn(v, D, generator=gen) for v in Vs]
    full1 = [torch.randn(v, 1, generator=gen) for v in Vs]
    ids_all, raw_all, gx_all, ge_all = [], [], [], []
    stride = (len(Vs) * D + 3 + 3) // 4 * 4
    for r in range(world):
        ids, raw = [], []
        for f, v in enumerate(Vs):
            t = torch.randint(0, v, (B, 1), generator=gen)
            t[: B // 3, 0] = torch.randint(0, min(v, 3), (B // 3,), generator=gen)       # hot rows shared by all ranks
            rid = t.clone()
            if hashed and f in (1, 5):
                from oracle import hashing as oh
                rid = torch.randint(0, 2 ** 40, (B, 1), generator=gen)
                rid[: B // 3, 0] = torch.randint(0, 4, (B // 3,), generator=gen)
                t = torch.from_numpy(oh.hash_bucket_ids(rid.numpy(), v, 7 + f)).long()
            ids.append(t)
            raw.append(rid)
        ids_all.append(ids)
        raw_all.append(raw)
        gx_all.append(torch.randn(B, stride, generator=gen))
        ge_all.append(torch.randn(B, generator=gen))
    dense = torch.randn(B, 3, generator=gen)
    return Vs, B, D, full, full1, ids_all, raw_all, gx_all, ge_all, dense, hashed


def _oracle_step(Vs, D, W, W1, acc, acc1, ids_all, gx_all, ge_all, dense, twins, fm, kind, lr):
    """Forward of every rank + the optimizer step on the summed gradients, with torch autograd on the CPU."""
    Wp = [w.clone().requires_grad_(True) for w in W]
    W1p = [w.clone().requires_grad_(True) for w in W1]
    outs, loss = [], 0.0
    for r, ids in enumerate(ids_all):
        v = [Wp[f][ids[f][:, 0]] for f in range(len(Vs))]
        x = torch.cat(v + [dense], 1)
        extra = torch.zeros(x.shape[0])
        if twins:
            extra = extra + sum(W1p[f][ids[f][:, 0], 0] for f in range(len(Vs)))
        if fm:
            s = sum(v)
            extra = extra + 0.5 * (s * s - sum(t * t for t in v)).sum(1)
        outs.append((x.detach(), extra.detach()))
        loss = loss + (x * gx_all[r][:, :x.shape[1]]).sum()
        if twins or fm:
            loss = loss + (extra * ge_all[r]).sum()
    loss.backward()
    for ws, ps, accs in ((W, Wp, acc), (W1, W1p, acc1)):
        for f in range(len(Vs)):
            g = ps[f].grad
            if g is None:
                continue
            if kind == "sgd":
                ws[f] -= lr * g
            else:
                accs[f] += g * g
                ws[f] -= lr * g / (accs[f].sqrt() + 1e-10)
    return outs


def _rank_main(rank, world, shared, prob, kind, twins, fm, errors, steps, hot_rows=0):
    try:
        from torchctr_b200.nn.embedding import EmbeddingTable
        from torchctr_b200.parallel.hybrid import HybridShardedTables
        from torchctr_b200.parallel.peer import ThreadTransport
        Vs, B, D, full, full1, ids_all, raw_all, gx_all, ge_all, dense, hashed = prob
        dev = torch.device("cuda", 0)
        tr = ThreadTransport(shared, rank, dev)
        names = [f"f{i}" for i in range(len(Vs))]
        kw = [dict(index_kind="hash", hash_seed=7 + f) if hashed and f in (1, 5) else {} for f in range(len(Vs))]
        tabs = [EmbeddingTable(v, D, _weight=w.clone(), **k) for v, w, k in zip(Vs, full, kw)]
        tabs1 = [EmbeddingTable(v, 1, _weight=w.clone(), **k) for v, w, k in zip(Vs, full1, kw)] if twins else None
        st = HybridShardedTables(names, tabs, tabs1, tr, dev, fm=fm, replicate_max_rows=200, hot_rows=hot_rows).train()
        assert sorted({st.parts[p][0] for p in st.sh}) == [1, 3, 4]
        if hot_rows:      # large direct-id tables: a replicated head [0, hot_rows) + a sharded tail; hashed ones stay whole
            split = [f for f in (1, 3, 4) if not (hashed and f == 1)]
            assert sorted(st.parts[p][0] for p in st.rp if st.parts[p][3] == "window") == split
        else:
            assert sorted(st.parts[p][0] for p in st.rp) == [0, 2, 5, 6]
        lr = 0.5
        opt = torch.optim.SGD(list(st.shards), lr=lr) if kind == "sgd" else torch.optim.Adagrad(list(st.shards), lr=lr)
        st.bind_optimizer(opt, kind=kind)
        W = [w.clone() for w in full]
        W1 = [w.clone() for w in full1]
        acc = [torch.zeros_like(w) for w in full]
        acc1 = [torch.zeros_like(w) for w in full1]
        feats = {n: t for n, t in zip(names, raw_all[rank])}
        for step in range(steps):
            # (threads cannot meet at a barrier inside autograd's backward: drive the two halves of the Function directly)
            ids_dev = st._ids_in_order(feats)
            x, extra = st._forward(ids_dev, dense.to(dev))
            outs = _oracle_step(Vs, D, W, W1, acc, acc1, ids_all, gx_all, ge_all, dense, twins, fm, kind, lr)
            ref_x, ref_e = outs[rank]
            width = ref_x.shape[1]
            assert _close(x[:, :width], ref_x), f"forward x step {step}"
            assert float(x[:, width:].abs().sum()) == 0.0
            if twins or fm:
                assert _close(extra, ref_e), f"forward extra step {step}: {float((extra.cpu() - ref_e).abs().max()):.3e}"
            else:
                assert extra is None
            gx = torch.zeros(B, x.shape[1])
            gx[:, :width] = gx_all[rank][:, :width]
            st._backward(gx.to(dev), ge_all[rank].to(dev) if (twins or fm) else None)
            torch.cuda.synchronize()
            for w, ws in enumerate((W, W1) if twins else (W,)):
                got = st.export_full_tables(w)               # collective: every table gathered back to [V, D]
                for f in range(len(Vs)):
                    rows, expect = got[f], ws[f]
                    assert rows.shape == expect.shape, (rows.shape, expect.shape)
                    # element-wise Adagrad moves an element by lr g / sqrt(sum g^2): where an element's gradients all but
                    # cancel (g = gy + c (fm_sum - v) is a sum of O(10) terms) the step is ill-conditioned in g, so the bound
                    # is 1e-5 of the table's scale plus the step's sensitivity to an fp32 rounding error of the gradient
                    tol = torch.full_like(expect, RTOL * float(expect.abs().max()))
                    if kind == "adagrad":
                        tol = tol + lr * 2e-5 / ((acc, acc1)[w][f].sqrt() + 1e-12)
                    err = (rows - expect).abs()
                    if bool((err > tol).any()):
                        bad = torch.nonzero((err > tol).any(1)).reshape(-1)[:8].tolist()
                        raise AssertionError(f"update width {w} table {f} step {step}: max abs err {float(err.max()):.3e}, rows {bad}, "
                                             f"errs {[round(float(err[i].max()), 6) for i in bad]}")
            assert int(st.status.item()) == 0
        # checkpoints in the reference's unsharded format
        if kind == "adagrad":
            for w in range(2 if twins else 1):
                s0, _ = st.export_full_optimizer_state(w)
                for f, a in enumerate((acc, acc1)[w]):
                    assert _close(s0[f], a), f"optimizer state width {w} table {f}"
        for w, ws in enumerate((W, W1) if twins else (W,)):
            fullw = st.export_full_tables(w)
            for f in range(len(Vs)):
                assert _close(fullw[f], ws[f], rtol=1e-3), f"export width {w} table {f}"      # (values were checked row by row above)
            st.load_full_tables([t * 0.5 for t in fullw], w)
            again = st.export_full_tables(w)
            for f in range(len(Vs)):
                assert torch.equal(again[f], fullw[f] * 0.5), f"load width {w} table {f}"
    except BaseException as e:          # noqa: BLE001 -- reported by the main thread
        import traceback
        errors.append((rank, repr(e)[:600], traceback.format_exc()[-1200:]))
        try:
            shared.barrier.abort()
        except Exception:
            pass


@pytest.mark.parametrize("world,kind,D,twins,fm,hashed,hot_rows", [
    (2, "sgd", 16, True, True, False, 0), (3, "sgd", 16, True, True, False, 100), (3, "adagrad", 16, True, True, False, 0),
    (4, "adagrad", 16, True, True, True, 100), (2, "adagrad", 32, False, False, False, 100), (3, "sgd", 64, False, False, True, 0),
    (2, "adagrad", 16, True, False, False, 0), (2, "sgd", 16, True, True, False, 100)])
def test_hybrid_sharded_threads(world, kind, D, twins, fm, hashed, hot_rows):
    import faulthandler
    import sys
    from torchctr_b200.parallel.peer import ThreadTransport
    faulthandler.dump_traceback_later(300, exit=True, file=sys.stderr)
    shared = ThreadTransport.Shared(world)
    prob = _problem(world, D, hashed=hashed)
    errors = []
    threads = [threading.Thread(target=_rank_main, args=(r, world, shared, prob, kind, twins, fm, errors, 2, hot_rows)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=240)
    faulthandler.cancel_dump_traceback_later()
    assert not errors, errors
    assert all(not t.is_alive() for t in threads)


def _model_rank(rank, world, shared, fc, state, batches, out, errors):
    try:
        from torchctr_b200.models import DeepFM
        from torchctr_b200.parallel import shard_model
        from torchctr_b200.parallel.hybrid import HybridShardedTables
        from torchctr_b200.parallel.peer import ThreadTransport
        dev = torch.device("cuda", 0)
        tr = ThreadTransport(shared, rank, dev)
        model = DeepFM(fc, [32, 16])
        model.load_state_dict(state)
        model = shard_model(model, transport=tr, device=dev, replicate_max_rows=100, hot_rows=32).to(dev).eval()
        assert isinstance(model._sharded, HybridShardedTables) and model._sharded.sh and model._sharded.rp
        assert any(part[3] == "window" for part in model._sharded.parts)          # hot rows of the large tables replicated
        with torch.no_grad():
            out[rank] = model(batches[rank]).cpu()
        sd = model.full_state_dict()
        assert set(sd) == set(state), set(sd) ^ set(state)
        for k, v in state.items():
            assert torch.equal(sd[k].cpu(), v), k
    except BaseException as e:          # noqa: BLE001
        import traceback
        errors.append((rank, repr(e)[:600], traceback.format_exc()[-1200:]))
        try:
            shared.barrier.abort()
        except Exception:
            pass


def test_hybrid_deepfm_model_forward_and_full_state_dict():
    """``shard_model`` puts a DeepFM on the hybrid placement (tables above ``replicate_max_rows`` sharded, the rest
    replicated); eval logits equal the unsharded model's, ``full_state_dict()`` returns the unsharded keys and values."""
    from torchctr_b200.models import DeepFM
    from torchctr_b200.parallel.peer import ThreadTransport
    world = 2
    gen = torch.Generator().manual_seed(4)
    Vs = [50, 400, 77, 1000, 9]
    fc = [{"name": f"c{i}", "type": "sparse", "num_embeddings": v, "emb_dim": 16} for i, v in enumerate(Vs)]
    fc += [{"name": f"d{i}", "type": "dense"} for i in range(3)]
    batches = []
    for r in range(world):
        feats = {f"c{i}": torch.randint(0, v, (200, 1), generator=gen) for i, v in enumerate(Vs)}
        feats["dense_features"] = torch.randn(200, 3, generator=gen)
        batches.append(feats)
    torch.manual_seed(0)
    ref = DeepFM(fc, [32, 16]).cuda().eval()
    state = {k: v.detach().cpu().clone() for k, v in ref.state_dict().items()}
    with torch.no_grad():
        expect = [ref(b).cpu() for b in batches]
    shared = ThreadTransport.Shared(world)
    out, errors = {}, []
    threads = [threading.Thread(target=_model_rank, args=(r, world, shared, fc, state, batches, out, errors)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=120)
    assert not errors, errors
    for r in range(world):
        assert _close(out[r], expect[r], rtol=2e-3), r            # (the tower GEMMs run in TF32 unless the exact mode is set)
