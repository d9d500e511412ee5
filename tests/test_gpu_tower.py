"""Fused tower block (BatchNorm1d training statistics + ReLU + Dropout, and their backward) against torch modules:
the reference tower is [Linear, BatchNorm1d, ReLU, Dropout(0.5)] x k (torchctr/models/dnn.py:39-45)."""
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


def _close(got, ref, rtol=1e-5):
    got, ref = got.detach().float().cpu(), ref.detach().float().cpu()
    assert got.shape == ref.shape, (got.shape, ref.shape)
    err = float((got - ref).abs().max())
    assert err <= rtol * max(float(ref.abs().max()), 1e-30), f"max abs err {err:.3e} vs scale {float(ref.abs().max()):.3e}"


@pytest.mark.parametrize("B,N", [(257, 64), (4096, 128), (65536, 256), (1000, 48), (3, 8), (5000, 1024)])
def test_bn_relu_forward_backward_match_torch(B, N):
    from torchctr_b200 import ops
    gen = torch.Generator().manual_seed(B + N)
    z = (torch.randn(B, N, generator=gen) * 2 + torch.randn(N, generator=gen)).cuda()
    gy = torch.randn(B, N, generator=gen).cuda()
    bn = nn.BatchNorm1d(N).cuda().train()
    with torch.no_grad():
        bn.weight.copy_(torch.rand(N, generator=gen) + 0.5)
        bn.bias.copy_(torch.randn(N, generator=gen) * 0.3)
    rm, rv, nbt = bn.running_mean.clone(), bn.running_var.clone(), bn.num_batches_tracked.clone()
    zr = z.clone().requires_grad_(True)
    y_ref = torch.relu(bn(zr))
    y_ref.backward(gy)
    mean, rstd = ops.bn_stats(z, bn.eps, bn.momentum, rm, rv, nbt)
    y = ops.bn_relu_dropout_fwd(z, mean, rstd, bn.weight.data, bn.bias.data, 0.0, None, 0)
    _close(y, y_ref)
    _close(rm, bn.running_mean)
    _close(rv, bn.running_var)
    assert int(nbt) == int(bn.num_batches_tracked) == 1
    gz, dgamma, dbeta, dbias = ops.bn_relu_dropout_bwd(gy, z, mean, rstd, bn.weight.data, bn.bias.data, 0.0, None, 0)
    _close(gz, zr.grad, rtol=2e-5)
    _close(dgamma, bn.weight.grad, rtol=2e-5)
    _close(dbeta, bn.bias.grad, rtol=2e-5)
    # column sums of gz vanish analytically (BatchNorm removes the mean): only rounding noise is left, as in torch
    assert float(dbias.abs().max()) <= 1e-3 * max(float(gz.abs().sum(0).max()), 1e-30)


def test_dropout_mask_is_bernoulli_consistent_and_reseeded():
    from torchctr_b200 import ops
    gen = torch.Generator().manual_seed(9)
    B, N, p = 8192, 256, 0.5
    z = torch.randn(B, N, generator=gen).cuda()
    gy = torch.randn(B, N, generator=gen).cuda()
    gamma, beta = torch.ones(N, device="cuda"), torch.full((N,), 3.0, device="cuda")     # + 3: almost every unit active
    seed = torch.tensor([1234], dtype=torch.int64, device="cuda")
    mean, rstd = ops.bn_stats(z, 1e-5, 0.1)
    y0 = ops.bn_relu_dropout_fwd(z, mean, rstd, gamma, beta, 0.0, None, 0)
    y = ops.bn_relu_dropout_fwd(z, mean, rstd, gamma, beta, p, seed, 7)
    active = y0 > 0
    kept = (y != 0) & active
    frac = float(kept.sum()) / float(active.sum())
    assert abs(frac - (1 - p)) < 0.005, frac
    _close(y, torch.where(kept, y0 / (1 - p), torch.zeros_like(y0)))
    # per-column keep rates are balanced too (the generator must not correlate with the layout)
    col = kept.float().mean(0)
    assert float(col.min()) > 0.45 and float(col.max()) < 0.55
    # same seed / layer -> same mask; other layer or next step -> a different one
    assert torch.equal(y, ops.bn_relu_dropout_fwd(z, mean, rstd, gamma, beta, p, seed, 7))
    assert not torch.equal(y, ops.bn_relu_dropout_fwd(z, mean, rstd, gamma, beta, p, seed, 8))
    seed2 = seed + 1
    assert not torch.equal(y, ops.bn_relu_dropout_fwd(z, mean, rstd, gamma, beta, p, seed2, 7))
    # backward sees the same mask: compare with torch autograd through an explicit mask
    zr = z.clone().requires_grad_(True)
    g_t, b_t = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    zh = (zr - zr.mean(0)) / torch.sqrt(zr.var(0, unbiased=False) + 1e-5)
    y_ref = torch.relu(zh * g_t + b_t) * kept.float() / (1 - p)
    y_ref.backward(gy)
    gz, dgamma, dbeta, _ = ops.bn_relu_dropout_bwd(gy, z, mean, rstd, gamma, beta, p, seed, 7)
    _close(gz, zr.grad, rtol=3e-5)
    _close(dgamma, g_t.grad, rtol=3e-5)
    _close(dbeta, b_t.grad, rtol=3e-5)


def test_tower_block_autograd_matches_torch_modules():
    """The fused node against nn.Linear + nn.BatchNorm1d + ReLU (Dropout p = 0), exact fp32 matmul mode."""
    from torchctr_b200.nn.tower import tower_block
    torch.backends.cuda.matmul.allow_tf32 = False
    gen = torch.Generator().manual_seed(3)
    B, K, N = 3000, 148, 64
    x = torch.randn(B, K, generator=gen).cuda()
    lin, bn, drop = nn.Linear(K, N).cuda(), nn.BatchNorm1d(N).cuda().train(), nn.Dropout(0.0)
    lin2, bn2 = nn.Linear(K, N).cuda(), nn.BatchNorm1d(N).cuda().train()
    lin2.load_state_dict(lin.state_dict())
    gy = torch.randn(B, N, generator=gen).cuda()
    xr = x.clone().requires_grad_(True)
    y_ref = torch.relu(bn2(lin2(xr)))
    y_ref.backward(gy)
    xf = x.clone().requires_grad_(True)
    seed = torch.zeros(1, dtype=torch.int64, device="cuda")
    y = tower_block(xf, lin, bn, drop, seed, 0)
    y.backward(gy)
    _close(y, y_ref, rtol=2e-5)
    _close(xf.grad, xr.grad, rtol=5e-5)
    _close(lin.weight.grad, lin2.weight.grad, rtol=5e-5)
    _close(bn.weight.grad, bn2.weight.grad, rtol=5e-5)
    _close(bn.bias.grad, bn2.bias.grad, rtol=5e-5)
    _close(bn.running_var, bn2.running_var)


@pytest.mark.parametrize("B,H,with_extra", [(65536, 64, True), (1000, 64, False), (257, 128, True), (33, 8, True)])
def test_logit_bce_head_matches_torch(B, H, with_extra):
    """Last Linear(H, 1) + extra logit terms + mean BCE-with-logits (dnn.py:46,68,75) as one node vs torch ops."""
    import torch.nn.functional as F
    from torchctr_b200.nn.head import head_eligible, logit_bce
    gen = torch.Generator().manual_seed(B + H)
    h = torch.randn(B, H, generator=gen).cuda().requires_grad_(True)
    lin = nn.Linear(H, 1).cuda()
    extra = (torch.randn(B, 1, generator=gen) * 2).cuda().requires_grad_(True) if with_extra else None
    labels = (torch.rand(B, 1, generator=gen) < 0.25).float().cuda()
    assert head_eligible(h, lin, labels)
    loss = logit_bce(h, lin, extra, labels)
    (loss * 0.5).backward()                     # a non-trivial upstream gradient (multi-GPU scales by 1 / world)
    got = (loss.detach(), h.grad.clone(), lin.weight.grad.clone(), lin.bias.grad.clone(), None if extra is None else extra.grad.clone())
    h.grad = lin.weight.grad = lin.bias.grad = None
    if extra is not None:
        extra.grad = None
    z = lin(h) + (extra if extra is not None else 0)
    ref = F.binary_cross_entropy_with_logits(z, labels)
    (ref * 0.5).backward()
    want = (ref.detach(), h.grad, lin.weight.grad, lin.bias.grad, None if extra is None else extra.grad)
    for g, r in zip(got, want):
        if r is not None:
            _close(g, r, rtol=2e-5)


@pytest.mark.parametrize("B,H,ne", [(65536, 64, 13), (999, 32, 32), (300, 128, 1)])
def test_logit_bce_head_second_linear_term(B, H, ne):
    """The head with DeepFM's Linear(Nd, 1) over the dense block folded in vs torch ops."""
    import torch.nn.functional as F
    from torchctr_b200.nn.head import logit_bce, second_term_eligible
    gen = torch.Generator().manual_seed(B + H + ne)
    h = torch.randn(B, H, generator=gen).cuda().requires_grad_(True)
    lin, lin_e = nn.Linear(H, 1).cuda(), nn.Linear(ne, 1).cuda()
    extra = (torch.randn(B, 1, generator=gen) * 2).cuda().requires_grad_(True)
    xe = torch.randn(B, ne, generator=gen).cuda()
    labels = (torch.rand(B, 1, generator=gen) < 0.25).float().cuda()
    assert second_term_eligible(xe, lin_e)
    loss = logit_bce(h, lin, extra, labels, xe=xe, linear_e=lin_e)
    (loss * 0.5).backward()
    params = [h, lin.weight, lin.bias, extra, lin_e.weight, lin_e.bias]
    got = [loss.detach()] + [p.grad.clone() for p in params]
    for p in params:
        p.grad = None
    ref = F.binary_cross_entropy_with_logits(lin(h) + extra + lin_e(xe), labels)
    (ref * 0.5).backward()
    for g, r in zip(got, [ref.detach()] + [p.grad for p in params]):
        _close(g, r, rtol=2e-5)


def test_fused_adagrad_equals_torch_adagrad():
    from torchctr_b200.optim import FusedAdagrad
    gen = torch.Generator().manual_seed(1)
    shapes = [(256, 432), (256,), (128, 256), (1,), (64, 128), (7, 3)]
    p1 = [torch.randn(*s, generator=gen).cuda().requires_grad_(True) for s in shapes]
    p2 = [p.detach().clone().requires_grad_(True) for p in p1]
    o1 = FusedAdagrad(p1, lr=0.05, eps=1e-10)
    o2 = torch.optim.Adagrad(p2, lr=0.05, eps=1e-10)
    for step in range(4):
        for a, b in zip(p1, p2):
            g = torch.randn(a.shape, generator=gen).cuda()
            a.grad, b.grad = g.clone(), g.clone()
        if step == 2:
            p1[3].grad = p2[3].grad = None          # a parameter without gradient is skipped, like torch does
        o1.step()
        o2.step()
    for a, b in zip(p1, p2):
        _close(a, b, rtol=1e-6)
    for a, b in zip(p1, p2):
        _close(o1.state[a]["sum"], o2.state[b]["sum"], rtol=1e-6)
        assert float(o1.state[a]["step"]) == float(o2.state[b]["step"])
    sd = o1.state_dict()
    o3 = torch.optim.Adagrad([p.detach().clone().requires_grad_(True) for p in p1], lr=0.05)
    o3.load_state_dict(sd)                          # same state layout as torch.optim.Adagrad


@pytest.mark.parametrize("precision", ["tf32", "tf32x3"])
def test_deferred_weight_gradients_equal_the_single_stream_schedule(precision):
    """nn/tower.py issues the weight / bias gradients of the tower and of the logit head on a second stream (joined when the
    backward pass ends) when autograd takes the gradient tensors over as they are.  Same kernels, same order of every sum:
    the gradients and the updated parameters are bit-identical to the single-stream schedule -- also when ``.grad`` is already
    set (accumulation), where everything stays on the one stream."""
    from torchctr_b200.models import DeepFM
    from torchctr_b200.nn import set_matmul_precision, tower
    gen = torch.Generator().manual_seed(5)
    B, F, D, nd = 2048, 5, 16, 3           # 5 * 16 + 3 = 83 -> padded to 84: the first layer's weight is padded inside the node
    fc = [{"name": f"c{i}", "type": "sparse", "num_embeddings": 200 + i, "emb_dim": D} for i in range(F)]
    fc += [{"name": f"d{i}", "type": "dense"} for i in range(nd)]
    batches = []
    for _ in range(3):
        feats = {f"c{i}": torch.randint(0, 200 + i, (B, 1), generator=gen).cuda() for i in range(F)}
        feats["dense_features"] = torch.randn(B, nd, generator=gen).cuda()
        batches.append((feats, (torch.rand(B, 1, generator=gen) < 0.25).float().cuda()))
    set_matmul_precision(precision)
    old = tower.defer_weight_grads
    try:
        results = []
        for defer, flat in ((False, False), (True, False), (True, True)):
            # flat: every dense .grad is a view of one buffer (the multi-GPU all-reduce layout): the second stream adds into it
            tower.defer_weight_grads = defer
            torch.manual_seed(3)
            m = DeepFM(fc, [64, 32]).cuda().train()
            opt = torch.optim.Adagrad(m.dense_parameters(), lr=0.05)
            m.bind_optimizer(opt, kind="adagrad")
            if flat:
                m.enable_flat_dense_grads()
            grads = []
            for i, batch in enumerate(batches):
                if i != 2 and flat:
                    m.zero_dense_grads()
                elif i != 2:
                    opt.zero_grad(set_to_none=True)      # step 2 accumulates on top of step 1's gradients
                m.training_step(batch, i).backward()
                grads.append({k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None})
                opt.step()
            torch.cuda.synchronize()
            assert not tower._deferred, "the second stream was not joined at the end of backward"
            results.append((grads, {k: v.clone() for k, v in m.state_dict().items()}))
        g0, s0 = results[0]
        for g1, s1 in results[1:]:
            for a, b in zip(g0, g1):
                assert a.keys() == b.keys() and len(a) >= 10
                for k in a:
                    assert torch.equal(a[k], b[k]), k
            for k in s0:
                assert torch.equal(s0[k], s1[k]), k
    finally:
        tower.defer_weight_grads = old
        set_matmul_precision(None)


@pytest.mark.parametrize("model_name", ["DNN", "DeepFM", "DCNv2"])
def test_first_block_prepared_next_to_the_lookup_is_bit_identical(model_name):
    """``CTRModelBase._prepare_tower`` issues the first block's dropout-seed snapshot, zero-padded weight and its transpose on the
    second stream BEFORE the lookup (``nn.tower.prepare_first_block``).  Same values, same kernels downstream: losses and
    parameters after three steps (dropout on) are bit-identical to the in-line preparation, and the prepared path is the one
    that ran from the second step on."""
    import torchctr_b200.models as models
    from torchctr_b200.nn import tower
    gen = torch.Generator().manual_seed(9)
    B, F, D, nd = 1024, 5, 16, 3
    fc = [{"name": f"c{i}", "type": "sparse", "num_embeddings": 300 + i, "emb_dim": D} for i in range(F)]
    fc += [{"name": f"d{i}", "type": "dense"} for i in range(nd)]
    batches = []
    for _ in range(3):
        feats = {f"c{i}": torch.randint(0, 300 + i, (B, 1), generator=gen).cuda() for i in range(F)}
        feats["dense_features"] = torch.randn(B, nd, generator=gen).cuda()
        batches.append((feats, (torch.rand(B, 1, generator=gen) < 0.25).float().cuda()))
    old = tower.prepare_early
    real = tower._TowerBlockFn.forward
    taken = []

    def spy(ctx, *args):
        taken.append(args[-1] is not None)           # the ``prepared`` argument
        return real(ctx, *args)
    try:
        tower._TowerBlockFn.forward = staticmethod(spy)
        results = []
        for early in (False, True):
            tower.prepare_early = early
            taken.clear()
            torch.manual_seed(3)
            m = getattr(models, model_name)(fc, [64, 32]).cuda().train()
            opt = torch.optim.Adagrad(m.dense_parameters(), lr=0.05)
            m.bind_optimizer(opt, kind="adagrad")
            losses = []
            for i, batch in enumerate(batches):
                opt.zero_grad(set_to_none=True)
                loss = m.training_step(batch, i)
                loss.backward()
                opt.step()
                losses.append(loss.detach().clone())
            torch.cuda.synchronize()
            assert not tower._deferred
            # two fused blocks per step; the first one is prepared from the second step on (the padding is known by then)
            assert taken == ([False, False] * 3 if not early else [False, False, True, False, True, False]), taken
            results.append((losses, {k: v.clone() for k, v in m.state_dict().items()}))
        (l0, s0), (l1, s1) = results
        for a, b in zip(l0, l1):
            assert torch.equal(a, b)
        for k in s0:
            assert torch.equal(s0[k], s1[k]), k
    finally:
        tower._TowerBlockFn.forward = real
        tower.prepare_early = old
