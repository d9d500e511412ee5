"""Row-sharded tables over peer memory (torchctr_b200.parallel.peer): every kernel of the multi-rank path --
sharded lookup, owner bucketing, owner-side pull + sort + fused update with device-side counts -- run here with
the ranks as THREADS of one process on one GPU (ThreadTransport), checked against the single-table oracle on the
global batch.  The 2-GPU run of the same module over CUDA IPC is tests/test_sharded.py::test_peer_two_gpus."""
import threading

import pytest
import torch

pytestmark = pytest.mark.gpu

RTOL = 1e-5     # relative to the largest reference magnitude (fp32 sums of up to ~100 gradients per hot row)


def _close(got, ref):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    if got.shape != ref.shape or ref.numel() == 0:
        return got.shape == ref.shape
    return float((got - ref).abs().max()) <= RTOL * max(float(ref.abs().max()), 1e-30)


def _problem(world, seed=5, hashed=False):
    """``hashed``: tables 1 and 4 take RAW ids and bucket them with murmur3 inside the kernels (BASELINE config 4:
    hash-embedding tables row-sharded over the GPUs); the oracle works on the bucketed rows."""
    gen = torch.Generator().manual_seed(seed)
    Vs, Ls, D, B = [37, 101, 2, 1000, 5000], [1, 6, 1, 3, 1], 16, 300
    full16 = [torch.randn(v, D, generator=gen) for v in Vs]
    full1 = [torch.randn(v, 1, generator=gen) for v in Vs]
    ids_all, g16_all, g1_all = [], [], []
    for r in range(world):
        ids = []
        for v, L in zip(Vs, Ls):
            t = torch.randint(0, v, (B, L), generator=gen)
            if L > 1:
                t[torch.rand(B, L, generator=gen) < 0.3] = -100
            t[:, 0] = torch.randint(0, min(v, 3), (B,), generator=gen)       # hot rows
            ids.append(t)
        ids_all.append(ids)
        g16_all.append(torch.randn(B, len(Vs) * D + 4, generator=gen))
        g1_all.append(torch.randn(B, 8, generator=gen))
    dense = torch.randn(B, 3, generator=gen)
    raw_all = None
    if hashed:
        from oracle import hashing as oh
        raw_all = []
        for r in range(world):
            raw = [t.clone() for t in ids_all[r]]
            for f in (1, 4):
                rid = torch.randint(0, 2 ** 40, ids_all[r][f].shape, generator=gen)
                rid[:, 0] = torch.randint(0, 5, (B,), generator=gen)                   # hot raw ids
                rid[ids_all[r][f] < 0] = -100
                rows = torch.from_numpy(oh.hash_bucket_ids(rid.clamp(min=0).numpy(), Vs[f], 7 + f)).long()
                rows[rid < 0] = -100
                raw[f] = rid
                ids_all[r][f] = rows                                                   # what the oracle indexes with
            raw_all.append(raw)
    return Vs, Ls, D, B, full16, full1, ids_all, g16_all, g1_all, dense, raw_all


def _rank_main(rank, world, shared, prob, kind, errors, steps, dedup=False):
    try:
        from oracle import embedding as oe
        from torchctr_b200.nn.embedding import EmbeddingTable
        from torchctr_b200.parallel.peer import PeerShardedTables, ThreadTransport, owned_rows
        Vs, Ls, D, B, full16, full1, ids_all, g16_all, g1_all, dense, raw_all = prob
        hashed = raw_all is not None
        dev = torch.device("cuda", 0)
        tr = ThreadTransport(shared, rank, dev)
        names = [f"f{i}" for i in range(len(Vs))]
        kw = [dict(index_kind="hash", hash_seed=7 + f) if hashed and f in (1, 4) else {} for f in range(len(Vs))]
        tabs16 = [EmbeddingTable(v, D, _weight=w.clone(), **k) for v, w, k in zip(Vs, full16, kw)]
        tabs1 = [EmbeddingTable(v, 1, _weight=w.clone(), **k) for v, w, k in zip(Vs, full1, kw)]
        st = PeerShardedTables(names, [tabs16, tabs1], tr, dev, dedup=dedup).train()
        opt = torch.optim.SGD(list(st.shards), lr=0.5) if kind == "sgd" else torch.optim.Adagrad(list(st.shards), lr=0.5)
        st.bind_optimizer(opt, kind=kind)
        feats = {n: t for n, t in zip(names, raw_all[rank] if hashed else ids_all[rank])}
        B = ids_all[rank][0].shape[0]                    # ranks may hold batches of different sizes (uneven last batch)
        dense = dense[:B]
        w16 = [w.clone() for w in full16]
        w1 = [w.clone() for w in full1]
        acc16 = [torch.zeros_like(w) for w in full16]
        acc1 = [torch.zeros_like(w) for w in full1]
        for step in range(steps):
            # the autograd engine runs every CUDA backward of a process on ONE thread per device, so ranks that are
            # threads cannot meet at a barrier inside backward(): drive the two halves of the Function directly
            ids_dev = [feats[n].to(dev) for n in names]
            x, first = st._forward(ids_dev, dense.to(dev))
            ref_x = torch.cat([oe.pooled_lookup(i, w) for i, w in zip(ids_all[rank], w16)] + [dense], 1)
            ref_1 = torch.cat([oe.pooled_lookup(i, w) for i, w in zip(ids_all[rank], w1)], 1)
            width = ref_x.shape[1]
            assert x.shape[1] % 4 == 0 and float(x[:, width:].abs().sum()) == 0.0
            err16 = (x[:, :width].cpu() - ref_x).abs()
            per_feature = [round(float(err16[:, f * D:(f + 1) * D].max()), 6) for f in range(len(Vs))]
            assert _close(x[:, :width], ref_x), \
                f"forward D=16 step {step}: max abs err per table {per_feature}, scale {float(ref_x.abs().max()):.2f}"
            assert _close(first[:, :len(Vs)], ref_1), f"forward D=1 step {step}"
            g16 = torch.zeros(B, x.shape[1])
            g16[:, :len(Vs) * D] = g16_all[rank][:, :len(Vs) * D]
            g1 = g1_all[rank][:, :first.shape[1]].contiguous()
            st._backward((g16.to(dev), g1.to(dev)))
            torch.cuda.synchronize()
            # oracle: the optimizer on the touched rows with the gradients of ALL ranks
            for ws, accs, gouts, Dw in ((w16, acc16, g16_all, D), (w1, acc1, g1_all, 1)):
                for f, v in enumerate(Vs):
                    dg = torch.zeros(v, Dw)
                    for r in range(world):
                        dg += oe.dense_table_grad(ids_all[r][f], gouts[r][:, f * Dw:(f + 1) * Dw], v)
                    if kind == "sgd":
                        ws[f] -= 0.5 * dg
                    else:
                        touched = torch.zeros(v, dtype=torch.bool)
                        for r in range(world):
                            flat = ids_all[r][f].reshape(-1)
                            touched[flat[flat >= 0]] = True
                        accs[f] += dg * dg
                        upd = 0.5 * dg / (accs[f].sqrt() + 1e-10)
                        ws[f][touched] -= upd[touched]
            for w, ws in enumerate((w16, w1)):
                for f, v in enumerate(Vs):
                    fr, rows = st.local_rows_of(w, f)
                    expect = ws[f][fr::world]
                    assert rows.shape == expect.shape, (rows.shape, expect.shape)
                    assert _close(rows, expect), \
                        f"update width {w} table {f} step {step}: max abs err {float((rows.cpu() - expect).abs().max()):.3e}"
            assert int(st.status.item()) == 0
        if kind == "adagrad":                            # the fused-update state gathers to the oracle's accumulators too
            for w, accs in enumerate((acc16, acc1)):
                s0, s1 = st.export_full_optimizer_state(w)
                assert s1 is None
                for f in range(len(Vs)):
                    assert _close(s0[f], accs[f]), f"optimizer state width {w} table {f}"
                st.load_full_optimizer_state(([t * 2 for t in s0], None), w)
                again0, _ = st.export_full_optimizer_state(w)
                assert all(torch.equal(a, t * 2) for a, t in zip(again0, s0))
        # checkpoint round trip in the reference's unsharded format: export == the oracle's tables, load restores them
        for w, ws in enumerate((w16, w1)):
            full = st.export_full_tables(w)
            for f in range(len(Vs)):
                assert _close(full[f], ws[f]), f"export width {w} table {f}"
            st.load_full_tables([t * 0.5 for t in full], w)
            again = st.export_full_tables(w)
            for f in range(len(Vs)):
                assert torch.equal(again[f], full[f] * 0.5), f"load width {w} table {f}"
    except BaseException as e:          # noqa: BLE001 -- reported by the main thread
        errors.append((rank, repr(e)[:600]))
        try:
            shared.barrier.abort()
        except Exception:
            pass


@pytest.mark.parametrize("world,kind,hashed,dedup", [(2, "sgd", False, False), (3, "sgd", False, False), (4, "adagrad", False, False),
                                                     (3, "adagrad", True, False), (2, "sgd", False, True), (4, "adagrad", True, True)])
def test_peer_sharded_threads(world, kind, hashed, dedup):
    import faulthandler
    import sys
    from torchctr_b200.parallel.peer import ThreadTransport
    faulthandler.dump_traceback_later(300, exit=True, file=sys.stderr)      # a stuck rank must not hang the suite
    shared = ThreadTransport.Shared(world)
    prob = _problem(world, hashed=hashed)
    errors = []
    threads = [threading.Thread(target=_rank_main, args=(r, world, shared, prob, kind, errors, 2, dedup)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=240)
    faulthandler.cancel_dump_traceback_later()
    assert not errors, errors
    assert all(not t.is_alive() for t in threads)


def test_uneven_batches_share_buffers_sized_for_the_largest():
    """the ranks agree on the buffer geometry collectively (sized for the largest batch any rank holds), a
    smaller batch -- the uneven last batch of an epoch -- reuses them, and a batch that does not fit raises instead of
    entering a collective alone."""
    import faulthandler
    import sys
    from torchctr_b200.parallel.peer import ThreadTransport
    faulthandler.dump_traceback_later(300, exit=True, file=sys.stderr)
    world = 3
    shared = ThreadTransport.Shared(world)
    prob = list(_problem(world))
    for r in range(world):                               # rank r holds B - 37 r samples
        n = prob[3] - 37 * r
        prob[6][r] = [t[:n] for t in prob[6][r]]
        prob[7][r] = prob[7][r][:n]
        prob[8][r] = prob[8][r][:n]
    errors = []
    threads = [threading.Thread(target=_rank_main, args=(r, world, shared, tuple(prob), "adagrad", errors, 2, False)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=240)
    faulthandler.cancel_dump_traceback_later()
    assert not errors, errors


def _native_init_rank(rank, world, shared, Vs, D, seed, out, errors):
    try:
        from torchctr_b200.nn.embedding import EmbeddingTable
        from torchctr_b200.parallel.peer import PeerShardedTables, ThreadTransport
        dev = torch.device("cuda", 0)
        tr = ThreadTransport(shared, rank, dev)
        tabs = [EmbeddingTable(v, D, device="meta") for v in Vs]                  # declared, never allocated
        tabs1 = [EmbeddingTable(v, 1, device="meta") for v in Vs]
        st = PeerShardedTables([f"f{i}" for i in range(len(Vs))], [tabs, tabs1], tr, dev, init_seed=seed)
        out[rank] = (st.export_full_tables(0), st.export_full_tables(1))
    except BaseException as e:          # noqa: BLE001
        errors.append((rank, repr(e)[:600]))
        try:
            shared.barrier.abort()
        except Exception:
            pass


@pytest.mark.parametrize("world", [2, 3, 8])
def test_shard_native_init_equals_unsharded_init(world):
    """BASELINE config 4 (2^30 rows x 64: no rank can hold a full table): tables declared on the meta device are created
    shard by shard with the counter-based generator keyed by the GLOBAL row, so ANY world size gives the table an unsharded
    model gets from ``counter_init_`` with the same seed -- bit for bit."""
    from torchctr_b200.nn.embedding import EmbeddingTable
    from torchctr_b200.parallel.peer import ThreadTransport
    Vs, D, seed = [37, 1000, 5, 4099], 64, 1234
    shared = ThreadTransport.Shared(world)
    out, errors = {}, []
    threads = [threading.Thread(target=_native_init_rank, args=(r, world, shared, Vs, D, seed, out, errors)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=120)
    assert not errors, errors
    for w, Dw in enumerate((D, 1)):
        for f, v in enumerate(Vs):
            ref = EmbeddingTable(v, Dw).cuda().counter_init_(EmbeddingTable.counter_seed(seed, f, w)).weight.detach().cpu()
            assert 0.8 < float(ref.std()) < 1.2 or v * Dw < 200
            for r in range(world):
                assert torch.equal(out[r][w][f], ref), (w, f, r)


def _sharded_model_rank(rank, world, shared, fc, state, feats, out, errors):
    try:
        from torchctr_b200.models import DeepFM
        from torchctr_b200.parallel import shard_model
        from torchctr_b200.parallel.peer import ThreadTransport
        dev = torch.device("cuda", 0)
        tr = ThreadTransport(shared, rank, dev)
        model = DeepFM(fc, [32, 16])
        model.load_state_dict(state)
        model = shard_model(model, transport=tr, device=dev).to(dev).eval()
        with torch.no_grad():
            out[rank] = model(feats).cpu()
        sd = model.full_state_dict()                     # collective: the reference-format state dict, tables gathered
        assert set(sd) == set(state), set(sd) ^ set(state)
        for k, v in state.items():
            assert torch.equal(sd[k].cpu(), v), k
    except BaseException as e:          # noqa: BLE001
        errors.append((rank, repr(e)[:600]))
        try:
            shared.barrier.abort()
        except Exception:
            pass


def test_sharded_deepfm_model_forward_and_full_state_dict():
    """Model level: ``shard_model`` on a DeepFM (two table widths sharing the ids, FM + first-order terms outside the fused
    single-GPU path) gives the logits of the unsharded model, and ``full_state_dict()`` returns exactly the unsharded keys
    and values."""
    from torchctr_b200.models import DeepFM
    from torchctr_b200.parallel.peer import ThreadTransport
    world = 2
    gen = torch.Generator().manual_seed(4)
    fc = [{"name": f"c{i}", "type": "sparse", "num_embeddings": 50 + 9 * i, "emb_dim": 16} for i in range(5)]
    fc += [{"name": f"d{i}", "type": "dense"} for i in range(3)]
    feats = {f"c{i}": torch.randint(0, 50 + 9 * i, (200, 1), generator=gen) for i in range(5)}
    feats["dense_features"] = torch.randn(200, 3, generator=gen)
    torch.manual_seed(0)
    ref = DeepFM(fc, [32, 16]).cuda().eval()
    state = {k: v.detach().cpu().clone() for k, v in ref.state_dict().items()}
    with torch.no_grad():
        expect = ref(feats).cpu()
    shared = ThreadTransport.Shared(world)
    out, errors = {}, []
    threads = [threading.Thread(target=_sharded_model_rank, args=(r, world, shared, fc, state, feats, out, errors)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=120)
    assert not errors, errors
    for r in range(world):
        assert _close(out[r], expect), r


def _vocab_rank(rank, world, shared, keys_all, ids_all, gouts, D, errors, report):
    try:
        from oracle import embedding as oe
        from oracle.vocab import Vocab
        from torchctr_b200.nn import VocabIndex
        from torchctr_b200.nn.embedding import EmbeddingTable
        from torchctr_b200.parallel.peer import PeerShardedTables, ThreadTransport
        dev = torch.device("cuda", 0)
        tr = ThreadTransport(shared, rank, dev)
        torch.manual_seed(0)
        seq = EmbeddingTable(1, D, index_kind="vocab", vocab=VocabIndex(capacity=256))         # one row so far: OOV
        seq.vocab_max_rows = 4096
        cat = EmbeddingTable(50, D)
        st = PeerShardedTables(["seq", "cat"], [[seq, cat]], tr, dev, init_seed=7).train()
        opt = torch.optim.SGD(list(st.shards), lr=0.5)
        st.bind_optimizer(opt, kind="sgd")
        ref_vocab = Vocab()
        full_seq = st.export_full_tables(0)[0].clone()                   # [1, D]
        full_cat = st.export_full_tables(0)[1].clone()
        for step in range(len(keys_all)):
            feats = {"seq": keys_all[step][rank], "cat": ids_all[step][rank]}
            st.grow_vocabularies(feats)                                  # collective: the GLOBAL batch's new keys, rank order
            glob = torch.cat([keys_all[step][r] for r in range(world)], 0).numpy()
            n = ref_vocab.fit(glob[glob >= 0])
            assert st.live_rows[0] == n, (st.live_rows[0], n)
            exported = st.export_full_tables(0)[0]
            assert exported.shape == (n, D)
            assert torch.equal(exported[:full_seq.shape[0]], full_seq)   # existing rows untouched by growth
            grown = exported[full_seq.shape[0]:]
            if grown.numel():
                assert 0.003 < float(grown.std()) < 0.03                 # N(0, 0.01), as DynamicEmbedding draws new rows
            full_seq = exported.clone()
            rows = {r: torch.from_numpy(ref_vocab.transform(keys_all[step][r].numpy())).long() for r in range(world)}
            for r in range(world):
                rows[r][keys_all[step][r] < 0] = -100
            x, = st._forward([feats["seq"].to(dev), feats["cat"].to(dev)], None)
            ref_x = torch.cat([oe.pooled_lookup(rows[rank], full_seq), oe.pooled_lookup(ids_all[step][rank], full_cat)], 1)
            assert _close(x[:, :2 * D], ref_x), f"forward step {step}"
            g = gouts[step][rank]
            st._backward((g.to(dev),))
            torch.cuda.synchronize()
            dg_seq = sum(oe.dense_table_grad(rows[r], gouts[step][r][:, :D], n) for r in range(world))
            dg_cat = sum(oe.dense_table_grad(ids_all[step][r], gouts[step][r][:, D:2 * D], 50) for r in range(world))
            full_seq = full_seq - 0.5 * dg_seq
            full_cat = full_cat - 0.5 * dg_cat
            got = st.export_full_tables(0)
            assert _close(got[0], full_seq) and _close(got[1], full_cat), f"update step {step}"
            full_seq, full_cat = got[0].clone(), got[1].clone()
        report[rank] = st.live_rows[0]
    except BaseException as e:          # noqa: BLE001
        errors.append((rank, repr(e)[:700]))
        try:
            shared.barrier.abort()
        except Exception:
            pass


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_growing_vocabulary(world):
    """BASELINE config 5 at N > 1: a variable-length sequence feature over an unbounded key space, pooled through a vocabulary
    that grows while training, with the table row-sharded.  The key -> row map is replicated and fitted on the all-gathered
    global batch, so rows are bit-exact those of ``oracle.vocab`` on the concatenated batch (transformer.py:451-498); grown
    rows appear on their owners only; pooled vectors and the SGD update match the single-table oracle."""
    from torchctr_b200.parallel.peer import ThreadTransport
    gen = torch.Generator().manual_seed(9)
    B, L, D, steps = 96, 7, 16, 3
    keys_all, ids_all, gouts = [], [], []
    for step in range(steps):
        ks, ids, gs = [], [], []
        for r in range(world):
            k = torch.randint(0, 400, (B, L), generator=gen) * 7919 + step * 10 ** 6      # fresh key range every step + repeats
            k[:, 0] = torch.randint(0, 5, (B,), generator=gen) * 7919                       # hot keys shared by all ranks
            k[torch.rand(B, L, generator=gen) < 0.3] = -100
            ks.append(k)
            ids.append(torch.randint(0, 50, (B, 1), generator=gen))
            gs.append(torch.randn(B, 2 * D, generator=gen))
        keys_all.append(ks); ids_all.append(ids); gouts.append(gs)
    shared = ThreadTransport.Shared(world)
    errors, report = [], {}
    threads = [threading.Thread(target=_vocab_rank, args=(r, world, shared, keys_all, ids_all, gouts, D, errors, report)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=240)
    assert not errors, errors
    assert len(set(report.values())) == 1 and report[0] > 300
