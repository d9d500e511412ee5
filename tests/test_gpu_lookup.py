"""GPU parity tests of the lookup / backward / optimizer kernels, through the C ABI
(torchctr_b200.ops -> libctr_b200.so), against the CPU oracle and the golden fixtures.

Tolerances (BASELINE.json north_star): bit-exact for hashed / remapped indices, dedup and vocab
growth; <= 1e-5 relative (fp32, relative to max|ref|) for pooled embeddings, gradients and updated rows.
"""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

RTOL = 1e-5


def close(got, ref, rtol=RTOL):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    assert got.shape == ref.shape, (got.shape, ref.shape)
    scale = max(float(ref.abs().max()), 1e-30)
    err = float((got - ref).abs().max())
    assert err <= rtol * scale, f"max abs err {err:.3e} > {rtol} * {scale:.3e}"


def make_ids(gen, B, L, V, pad_frac=0.3, hot=0):
    ids = torch.randint(0, V, (B, L), generator=gen)
    if L > 1:
        lens = torch.randint(0, L + 1, (B,), generator=gen)
        ids[torch.arange(L).unsqueeze(0) >= lens.unsqueeze(1)] = -100
    elif pad_frac:
        ids[torch.rand(B, 1, generator=gen) < pad_frac * 0.2] = -100
    if hot:
        ids[:, 0] = torch.randint(0, hot, (B,), generator=gen)
    return ids


def run_fwd(specs_cpu, B, dense=None, dev="cuda"):
    """specs_cpu: list of dict(ids, table, pooling, id_weight, index_kind, seed, num_rows).  Returns (out, layout)."""
    from torchctr_b200 import ops
    col, cols = 0, []
    for s in specs_cpu:
        cols.append(col)
        col += s["table"].shape[1]
    dense_col, width = col, col + (0 if dense is None else dense.shape[1])
    stride = (width + 3) // 4 * 4
    out = torch.full((B, stride), float("nan"), device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    feats = []
    for s, c in zip(specs_cpu, cols):
        feats.append(ops.FeatureSpec(
            ids=s["ids"].to(dev), table=s["table"].to(dev).contiguous(), num_rows=s["table"].shape[0],
            D=s["table"].shape[1], out_col=c, pooling=s.get("pooling", "sum"), index_kind=s.get("index_kind", "direct"),
            hash_seed=s.get("seed", 0), id_weight=None if s.get("id_weight") is None else s["id_weight"].to(dev),
            bag_scale=torch.empty(B, device=dev) if s.get("pooling") == "mean" else None))
    call = ops.make_group(feats, B, out, stride, dense=None if dense is None else dense.to(dev), dense_col=dense_col,
                          zero_from=width if stride > width else -1, status=status)
    ops.emb_pool_fwd(call)
    torch.cuda.synchronize()
    return out, cols, width, feats, int(status.item())


@pytest.mark.parametrize("D,L", [(16, 1), (16, 12), (16, 50), (16, 200), (9, 1), (9, 7), (32, 1), (32, 33), (64, 1),
                                 (64, 20), (1, 1), (1, 5), (128, 3), (4, 2), (8, 1)])
def test_pooled_forward_matches_oracle(D, L):
    from oracle import embedding as oe
    gen = torch.Generator().manual_seed(100 * D + L)
    B, V = 301, 97
    ids = make_ids(gen, B, L, V)
    ids[3] = -100                                       # an empty bag -> zero vector
    table = torch.randn(V, D, generator=gen)
    for pooling in ("sum", "mean"):
        out, cols, width, _, status = run_fwd([dict(ids=ids, table=table, pooling=pooling)], B)
        assert status == 0
        close(out[:, :D], oe.pooled_lookup(ids, table, pooling))
        assert float(out[:, width:].abs().sum()) == 0.0     # padding columns are zero
        assert float(out[3, :D].abs().sum()) == 0.0


def test_forward_group_with_dense_and_weights():
    from oracle import embedding as oe
    gen = torch.Generator().manual_seed(7)
    B = 1000
    dims = [16, 16, 9, 16, 1, 32]
    Ls = [1, 1, 1, 12, 1, 40]
    specs, refs = [], []
    for D, L in zip(dims, Ls):
        V = 50 + D
        ids = make_ids(gen, B, L, V)
        table = torch.randn(V, D, generator=gen)
        w = torch.rand(B, L, generator=gen) if L > 1 else None
        specs.append(dict(ids=ids, table=table, id_weight=w, pooling="mean" if D == 32 else "sum"))
        refs.append(oe.pooled_lookup(ids, table, "mean" if D == 32 else "sum", per_id_weight=w))
    dense = torch.randn(B, 5, generator=gen)
    out, cols, width, _, status = run_fwd(specs, B, dense=dense)
    assert status == 0
    ref = torch.cat(refs + [dense], dim=1)
    close(out[:, :width], ref)
    assert out.shape[1] % 4 == 0 and float(out[:, width:].abs().sum()) == 0.0


def test_l1_lookup_is_bit_exact_copy_of_rows():
    """Full BASELINE config-2 size: 26 single-id features, B=65536, D=16 -- a pure row copy."""
    gen = torch.Generator().manual_seed(11)
    B, F, D = 65536, 26, 16
    specs = []
    for f in range(F):
        V = [1000, 50000, 2_000_000][f % 3]
        specs.append(dict(ids=torch.randint(0, V, (B, 1), generator=gen), table=torch.randn(V, D, generator=gen)))
    dense = torch.randn(B, 13, generator=gen)
    out, cols, width, feats, status = run_fwd(specs, B, dense=dense)
    assert status == 0 and width == 429 and out.shape[1] == 432
    for f, c in zip(feats, cols):
        assert torch.equal(out[:, c:c + D], f.table[f.ids[:, 0]])
    assert torch.equal(out[:, 416:429].cpu(), dense)


def test_out_of_range_id_sets_status_not_memory():
    gen = torch.Generator().manual_seed(5)
    ids = torch.randint(0, 10, (64, 3), generator=gen)
    ids[5, 1] = 10                                           # == num_rows -> IndexError upstream
    out, _, _, _, status = run_fwd([dict(ids=ids, table=torch.randn(10, 16, generator=gen))], 64)
    assert status & 1


def test_hash_bucket_matches_sklearn_golden(golden_dir):
    from torchctr_b200 import ops
    with open(os.path.join(golden_dir, "hash_golden.json")) as f:
        rows = [r for r in json.load(f) if r["is_int"]]
    for seed in sorted({r["seed"] for r in rows}):
        for buckets in sorted({r["buckets"] for r in rows}):
            sel = [r for r in rows if r["seed"] == seed and r["buckets"] == buckets]
            ids = torch.tensor([r["v"] for r in sel], dtype=torch.int64, device="cuda")
            got = ops.hash_bucket(ids, buckets, seed).cpu().numpy()
            assert (got == np.array([r["bucket"] for r in sel], dtype=np.int32)).all()


def test_hash_bucket_matches_oracle_on_random_ids():
    from oracle import hashing as oh
    from torchctr_b200 import ops
    rng = np.random.default_rng(3)
    ids = np.concatenate([rng.integers(0, 2 ** 31, 200000), rng.integers(-2 ** 62, 2 ** 62, 50000),
                          np.array([0, 9, 10, 99, 100, 2 ** 31 - 1, -1, 2 ** 63 - 1, -2 ** 63, 10 ** 18, 10 ** 9, 999999999])])
    for seed, buckets in ((0, 1000003), (12345, 2 ** 31 - 1), (2 ** 32 - 1, 100)):
        got = ops.hash_bucket(torch.from_numpy(ids).cuda(), buckets, seed).cpu().numpy()
        assert (got == oh.hash_bucket_ids(ids, buckets, seed)).all()
    with pytest.raises(OverflowError):
        ops.hash_bucket(torch.zeros(1, dtype=torch.int64, device="cuda"), 10, 2 ** 32)


def test_hashed_lookup_fuses_hash_bucket():
    from oracle import embedding as oe
    from oracle import hashing as oh
    gen = torch.Generator().manual_seed(21)
    B, L, V, D = 500, 6, 1009, 16
    raw = torch.randint(0, 2 ** 40, (B, L), generator=gen)
    raw[torch.rand(B, L, generator=gen) < 0.3] = -100
    table = torch.randn(V, D, generator=gen)
    rows = torch.from_numpy(oh.hash_bucket_ids(raw.clamp(min=0).numpy(), V, 7)).long()
    rows[raw < 0] = -100
    out, *_ = run_fwd([dict(ids=raw, table=table, index_kind="hash", seed=7)], B)
    close(out[:, :D], oe.pooled_lookup(rows, table))


# ------------------------------------------------------------------------------------------------
def run_bwd(specs_cpu, B, grad_out, opt_kind="none", opt_kw=None, states=None, dev="cuda"):
    from torchctr_b200 import ops
    col, cols = 0, []
    for s in specs_cpu:
        cols.append(col)
        col += s["table"].shape[1]
    feats = []
    for i, (s, c) in enumerate(zip(specs_cpu, cols)):
        st = states[i] if states else (None, None)
        feats.append(ops.FeatureSpec(
            ids=s["ids"].to(dev), table=s["table"].to(dev).contiguous(), num_rows=s["table"].shape[0],
            D=s["table"].shape[1], out_col=c, pooling=s.get("pooling", "sum"), index_kind=s.get("index_kind", "direct"),
            hash_seed=s.get("seed", 0), id_weight=None if s.get("id_weight") is None else s["id_weight"].to(dev),
            state0=None if st[0] is None else st[0].to(dev), state1=None if st[1] is None else st[1].to(dev),
            bag_scale=None if s.get("bag_scale") is None else s["bag_scale"].to(dev)))
    g = grad_out.to(dev).contiguous()
    call = ops.make_group(feats, B, g, g.shape[1])
    ws = torch.empty(ops.emb_bwd_workspace_bytes(call) + 256, dtype=torch.uint8, device=dev)
    ops.emb_bwd_plan(call, ws)
    S = sum(s["ids"].numel() for s in specs_cpu)
    dmax = max(s["table"].shape[1] for s in specs_cpu)
    uf = torch.full((S,), -1, dtype=torch.int32, device=dev)
    ur = torch.full((S,), -1, dtype=torch.int32, device=dev)
    rg = torch.zeros(S, dmax, device=dev)
    nu = torch.zeros(1, dtype=torch.int64, device=dev)
    ops.emb_bwd_apply(call, ws, ops.make_opt(opt_kind, **(opt_kw or {})), uf, ur, rg, nu)
    torch.cuda.synchronize()
    U = int(nu.item())
    return feats, uf[:U].cpu(), ur[:U].cpu(), rg[:U].cpu(), U


@pytest.mark.parametrize("B,hot", [(257, 0), (5000, 3), (20000, 1)])
def test_backward_unique_rows_and_grads_match_oracle(B, hot):
    from oracle import embedding as oe
    gen = torch.Generator().manual_seed(B)
    dims, Ls = [16, 16, 9, 1], [1, 12, 3, 1]
    specs, col = [], 0
    for D, L in zip(dims, Ls):
        V = 40 + 3 * D
        specs.append(dict(ids=make_ids(gen, B, L, V, hot=hot), table=torch.randn(V, D, generator=gen)))
        col += D
    gout = torch.randn(B, (col + 3) // 4 * 4, generator=gen)
    feats, uf, ur, rg, U = run_bwd(specs, B, gout)
    c = 0
    total = 0
    for i, s in enumerate(specs):
        D = s["table"].shape[1]
        rows, grads = oe.unique_row_grads(s["ids"], gout[:, c:c + D], s["table"].shape[0])
        sel = uf == i
        assert torch.equal(ur[sel].long(), rows)            # dedup is bit-exact, sorted
        close(rg[sel][:, :D], grads)
        total += rows.numel()
        c += D
    assert U == total


def test_backward_mean_and_weights():
    from oracle import embedding as oe
    gen = torch.Generator().manual_seed(99)
    B, L, V, D = 700, 9, 60, 16
    ids = make_ids(gen, B, L, V)
    w = torch.rand(B, L, generator=gen)
    table = torch.randn(V, D, generator=gen)
    gout = torch.randn(B, D, generator=gen)
    scale = oe.bag_scale(ids, "mean")
    feats, uf, ur, rg, U = run_bwd([dict(ids=ids, table=table, pooling="mean", id_weight=w, bag_scale=scale)], B, gout)
    rows, grads = oe.unique_row_grads(ids, gout, V, "mean", per_id_weight=w)
    assert torch.equal(ur.long(), rows)
    close(rg, grads)


def test_optimizers_match_torch_golden(golden_dir):
    """Three steps of torch.optim.{Adagrad, SparseAdam, SGD} recorded from real torch (optim_golden.pt)."""
    g = torch.load(os.path.join(golden_dir, "optim_golden.pt"))
    V, D = g["w0"].shape
    for kind, key, kw in (("adagrad", "adagrad", dict(lr=0.1, eps=1e-10)),
                          ("adam", "sparse_adam", dict(lr=0.01, eps=1e-8, betas=(0.9, 0.999))),
                          ("sgd", "sgd", dict(lr=0.1))):
        w = g["w0"].clone().cuda()
        s0 = torch.zeros(V, D, device="cuda")
        s1 = torch.zeros(V, D, device="cuda")
        for step, (ids, gout) in enumerate(g["steps"], start=1):
            spec = dict(ids=ids, table=w)
            feats, *_ = run_bwd([spec], ids.shape[0], gout, kind, dict(kw, step=step), states=[(s0, s1)])
            w, s0, s1 = feats[0].table, feats[0].state0, feats[0].state1
        close(w, g[key]["w"])
        if kind == "adagrad":
            close(s0, g[key]["sum"])
        if kind == "adam":
            close(s0, g[key]["exp_avg"])
            close(s1, g[key]["exp_avg_sq"])


def test_rowwise_adagrad_matches_oracle():
    from oracle import embedding as oe
    from oracle import optim as oo
    gen = torch.Generator().manual_seed(4)
    B, L, V, D = 3000, 4, 50, 16
    ids = make_ids(gen, B, L, V, hot=2)
    table = torch.randn(V, D, generator=gen)
    gout = torch.randn(B, D, generator=gen)
    st = torch.rand(V, generator=gen)
    feats, *_ = run_bwd([dict(ids=ids, table=table)], B, gout, "rowwise_adagrad", dict(lr=0.05, eps=1e-10),
                        states=[(st.clone(), None)])
    rows, grads = oe.unique_row_grads(ids, gout, V)
    w_ref, s_ref = table.clone(), st.clone()
    oo.rowwise_adagrad_rows(w_ref, s_ref, rows, grads, 0.05, 1e-10)
    close(feats[0].table, w_ref)
    close(feats[0].state0, s_ref)


def test_full_size_sgd_equals_index_add():
    """Config-2 size (26 x 65536 ids, Zipf-like): fused SGD == table - lr * index_add(grad) on the GPU."""
    gen = torch.Generator().manual_seed(2)
    B, F, D = 65536, 26, 16
    specs = []
    for f in range(F):
        V = [1000, 50000, 2_000_000][f % 3]
        u = torch.rand(B, 1, generator=gen)
        ids = (V ** u - 1).long().clamp(0, V - 1)             # log-uniform: hot head, long tail
        specs.append(dict(ids=ids, table=torch.randn(V, D, generator=gen)))
    gout = torch.randn(B, F * D, generator=gen)
    feats, uf, ur, rg, U = run_bwd(specs, B, gout, "sgd", dict(lr=0.5))
    g = gout.cuda()
    nuniq = 0
    for f, (s, ft) in enumerate(zip(specs, feats)):
        ref = s["table"].cuda().index_add(0, s["ids"].cuda()[:, 0], g[:, f * D:(f + 1) * D], alpha=-0.5)
        close(ft.table, ref, rtol=2e-5)                       # atomics in index_add: order differs
        nuniq += int(torch.unique(s["ids"]).numel())
    assert U == nuniq


@pytest.mark.parametrize("D", [1, 3, 16, 32, 64, 128])
@pytest.mark.parametrize("pattern", ["one_row", "distinct", "blocks", "zipf"])
def test_sweep_handles_every_run_length(D, pattern):
    """The segmented sweep must not care how the sorted positions split into runs: one giant run that crosses
    every warp range, no repeats at all, runs of ragged lengths straddling range boundaries, Zipf."""
    from oracle import embedding as oe
    gen = torch.Generator().manual_seed(1000 + D)
    B, L = 6000, 3
    V = 2 * B * L
    if pattern == "one_row":
        ids = torch.full((B, L), 7, dtype=torch.int64)
    elif pattern == "distinct":
        ids = torch.randperm(V, generator=gen)[: B * L].reshape(B, L)
    elif pattern == "blocks":
        lens = torch.randint(1, 700, (200,), generator=gen)
        vals = torch.repeat_interleave(torch.randperm(V, generator=gen)[:200], lens)[: B * L]
        ids = torch.cat([vals, torch.randint(0, V, (B * L - vals.numel(),), generator=gen)])
        ids = ids[torch.randperm(B * L, generator=gen)].reshape(B, L)
    else:
        u = torch.rand(B, L, generator=gen)
        ids = (V ** u - 1).long().clamp(0, V - 1)
    ids[torch.rand(B, L, generator=gen) < 0.1] = -100
    table = torch.randn(V, D, generator=gen)
    gout = torch.randn(B, (D + 3) // 4 * 4, generator=gen)
    feats, uf, ur, rg, U = run_bwd([dict(ids=ids, table=table)], B, gout, "sgd", dict(lr=0.25))
    rows, grads = oe.unique_row_grads(ids, gout[:, :D], V)
    assert U == rows.numel() and torch.equal(ur.long(), rows)
    close(rg[:, :D], grads, rtol=2e-5)
    ref = table.clone()
    ref[rows] -= 0.25 * grads
    close(feats[0].table, ref, rtol=2e-5)


def test_sort_only_plan_counts_unique_rows_and_updates():
    """CTR_PLAN_NO_RUNS (what a fused training step uses): same update, num_unique still reported."""
    from torchctr_b200 import ops
    gen = torch.Generator().manual_seed(77)
    B, V, D = 30000, 5000, 16
    u = torch.rand(B, 1, generator=gen)
    ids = (V ** u - 1).long().clamp(0, V - 1)
    table = torch.randn(V, D, generator=gen)
    gout = torch.randn(B, D, generator=gen)
    t = table.cuda()
    spec = ops.FeatureSpec(ids=ids.cuda(), table=t, num_rows=V, D=D, out_col=0)
    call = ops.make_group([spec], B, gout.cuda(), D)
    ws = torch.empty(ops.emb_bwd_workspace_bytes(call) + 256, dtype=torch.uint8, device="cuda")
    ops.emb_bwd_plan(call, ws, runs=False)
    nu = torch.zeros(1, dtype=torch.int64, device="cuda")
    ops.emb_bwd_apply(call, ws, ops.make_opt("sgd", lr=0.5), num_unique=nu)
    ref = table.cuda().index_add(0, ids.cuda()[:, 0], gout.cuda(), alpha=-0.5)
    close(t, ref, rtol=2e-5)
    assert int(nu.item()) == int(torch.unique(ids).numel())


@pytest.mark.parametrize("D,F,opt_kind", [(16, 26, "adagrad"), (16, 5, "sgd"), (32, 7, "adam"), (64, 3, "rowwise_adagrad"), (16, 9, "adagrad")])
def test_fused_twin_and_fm_terms_match_the_separate_kernels(D, F, opt_kind):
    """ctr_group_t.extra: the first-order (twin) tables and the FM second-order term folded into the single-id lookup
    and into the fused update must equal the oracle's definitions: extra[b] = sum_f w1_f[id] + FM2(v_b), row gradients
    = grad_out slice + g_extra[b] * (sum_f v_bf - v_bf), twin gradients = sum of g_extra over the run -- then the same
    optimizer arithmetic as torch.optim on the touched rows (1e-5)."""
    from oracle import embedding as oe
    from oracle import optim as oo
    from oracle.models import fm_second_order
    from torchctr_b200 import ops
    gen = torch.Generator().manual_seed(D * F)
    B = 3000
    Vs = [37 + 11 * i if i % 2 else 5000 + i for i in range(F)]
    hashed = F == 9                                             # raw ids through the in-kernel murmur3 as well
    tables = [torch.randn(V, D, generator=gen) for V in Vs]
    twins = [torch.randn(V, generator=gen) for V in Vs]
    if hashed:
        raw = [torch.randint(0, 2 ** 40, (B, 1), generator=gen) for _ in Vs]
        from oracle import hashing as oh
        rows = [torch.from_numpy(oh.hash_bucket_ids(r.numpy(), V, 7)).long() for r, V in zip(raw, Vs)]
    else:
        # Zipf-like: long runs of the hot rows cross many sweep ranges
        rows = [((V ** torch.rand(B, 1, generator=gen)) - 1).long().clamp_(0, V - 1) for V in Vs]
        raw = rows
    stride = (F * D + 3 + 3) // 4 * 4
    dense = torch.randn(B, 3, generator=gen)
    dev = "cuda"
    t_dev = [t.clone().to(dev) for t in tables]
    w_dev = [t.clone().to(dev) for t in twins]
    kind_states = {"sgd": 0, "adagrad": 1, "rowwise_adagrad": 1, "adam": 2}[opt_kind]
    def state_like(t, rowwise):
        return torch.zeros(t.shape[0] if rowwise else t.shape, device=dev)
    s0 = [state_like(t, opt_kind == "rowwise_adagrad") for t in t_dev] if kind_states >= 1 else [None] * F
    s1 = [torch.zeros_like(t) for t in t_dev] if kind_states == 2 else [None] * F
    ws0 = [torch.zeros_like(t) for t in w_dev] if kind_states >= 1 else [None] * F
    ws1 = [torch.zeros_like(t) for t in w_dev] if kind_states == 2 else [None] * F
    ids_dev = [r.to(dev) for r in raw]
    out = torch.empty(B, stride, device=dev)
    extra = torch.empty(B, device=dev)
    fm_sum = torch.empty(B, D, device=dev)
    def specs(with_state):
        return [ops.FeatureSpec(ids=ids_dev[i], table=t_dev[i], num_rows=Vs[i], D=D, out_col=i * D,
                                index_kind="hash" if hashed else "direct", hash_seed=7, twin_table=w_dev[i],
                                state0=s0[i] if with_state else None, state1=s1[i] if with_state else None,
                                twin_state0=ws0[i] if with_state else None, twin_state1=ws1[i] if with_state else None)
                for i in range(F)]
    fwd = ops.make_group(specs(False), B, out, stride, dense=dense.to(dev), dense_col=F * D, zero_from=F * D + 3,
                         extra=extra, fm_sum=fm_sum, fm=True)
    ops.emb_pool_fwd(fwd)
    pooled = [t[r[:, 0]] for t, r in zip(tables, rows)]
    ref_out = torch.zeros(B, stride)
    ref_out[:, :F * D] = torch.cat(pooled, 1)
    ref_out[:, F * D:F * D + 3] = dense
    ref_extra = sum(w[r[:, 0]] for w, r in zip(twins, rows)) + fm_second_order(torch.stack(pooled, 1))[:, 0]
    assert torch.equal(out.cpu(), ref_out)                       # single-id lookups are copies: bit-exact
    close(extra.cpu(), ref_extra, 1e-5)
    close(fm_sum.cpu(), torch.stack(pooled, 1).sum(1), 1e-5)
    # backward + fused update
    gout = torch.randn(B, stride, generator=gen) * 0.1
    gextra = torch.randn(B, generator=gen) * 0.1
    bwd = ops.make_group(specs(True), B, gout.to(dev), stride, extra=gextra.to(dev), fm_sum=fm_sum, fm=True)
    ws = torch.empty(ops.emb_bwd_workspace_bytes(bwd) + 256, dtype=torch.uint8, device=dev)
    ops.emb_bwd_plan(bwd, ws, runs=False)
    lr = 0.05
    ops.emb_bwd_apply(bwd, ws, ops.make_opt(opt_kind, lr=lr, eps=1e-8 if opt_kind == "adam" else 1e-10, step=1))
    torch.cuda.synchronize()
    ssum = torch.stack(pooled, 1).sum(1)
    for i in range(F):
        slot_grad = gout[:, i * D:(i + 1) * D] + gextra[:, None] * (ssum - pooled[i])
        g_rows = torch.zeros(Vs[i], D).index_add_(0, rows[i][:, 0], slot_grad)
        g_twin = torch.zeros(Vs[i]).index_add_(0, rows[i][:, 0], gextra)
        touched = torch.zeros(Vs[i], dtype=torch.bool)
        touched[rows[i][:, 0]] = True
        for tab, grad, got, wide in ((tables[i], g_rows, t_dev[i], True), (twins[i][:, None], g_twin[:, None], w_dev[i][:, None], False)):
            ref = tab.clone()
            g = grad[touched]
            if opt_kind == "sgd":
                ref[touched] -= lr * g
            elif opt_kind == "adagrad" or (opt_kind == "rowwise_adagrad" and not wide):
                ref[touched] -= lr * g / (g.pow(2).sqrt() + 1e-10)
            elif opt_kind == "rowwise_adagrad":
                ref[touched] -= lr * g / (g.pow(2).mean(1, keepdim=True).sqrt() + 1e-10)
            else:                                                   # lazy Adam, step 1: m = (1-b1) g, v = (1-b2) g^2
                m, v = 0.1 * g, 0.001 * g * g
                ref[touched] -= (lr * (1 - 0.999) ** 0.5 / (1 - 0.9)) * m / (v.sqrt() + 1e-8)
            close(got.cpu(), ref, 2e-5 if opt_kind != "sgd" else 1e-5)
            assert torch.equal(got.cpu()[~touched], tab[~touched])      # untouched rows do not move
