"""The benchmarked precision mode is the verified mode.

The tower / cross GEMMs run on the hand-written tcgen05 kernels in one of two precisions
(``torchctr_b200.nn.linear``):

* ``tf32x3`` -- error-compensated split (``ctr_split_tf32``): fp32-grade.  Tested against fp64 / the CPU fp32 oracle at
  the north_star bound (1e-5) at GEMM level and for the loss; quantities behind BatchNorm's 1/std and Adagrad's
  g / (|g| + eps) are compared at 1e-4 (summation-order noise of ANY fp32 GEMM is amplified there: the CPU oracle
  and torch's own CUDA fp32 path differ from each other by the same amount, which the test measures and prints).
* ``tf32`` -- what ``bench.py`` runs.  TF32 keeps 10 mantissa bits (unit round-off 2^-11 = 4.9e-4), so a K-term dot
  product is off by ~4.9e-4 * sqrt(2) * |a||w| / sqrt(K)-ish; the stated bounds are 5e-3 for the loss, 3e-2 for tower
  gradients / updated rows relative to max|ref| -- and, as SURVEY.md 7 (hard part 3) prescribes, the SAME bounds must
  hold against torch's own TF32 path (the oracle moved to CUDA with ``allow_tf32 = True``), i.e. we are no further from
  fp32 than the library the reference would run on this GPU.

The model-level steps use plain SGD: the first Adagrad step of an element is lr * g / (|g| + eps) = lr * sign(g), so any
element whose gradient is smaller than the GEMM's round-off flips by 2 * lr whatever the precision (among the 2.1 M cross
weights of DCN-v2 a few do even between torch's CPU and CUDA fp32 paths); the fused Adagrad / Adam row updates are
compared with ``torch.optim`` at kernel level with injected gradients (``test_gpu_lookup.py``) and through the
reference's golden DNN fixture below in the exact mode.

Model shapes are BASELINE configs[1] / configs[2] widths: DeepFM 26 x 16 + 13 = 429 -> 256 -> 128 -> 64 -> 1 and
DCN-v2 26 x 32 + 13 = 845 with three 845 x 845 cross layers, at a batch the CPU oracle finishes in seconds.
Measured errors are appended to ``gpurun_out/precision_errors.json`` for DESIGN.md.
"""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

TOL = {
    #  loss | dense (tower / cross) gradients | table-row updates (row_after - row_before) | golden-fixture state
    "tf32x3": dict(loss=1e-5, grad=1e-4, rows=1e-5, state=1e-4),
    "tf32": dict(loss=5e-3, grad=3e-2, rows=3e-2, state=3e-2),
}


def rel_err(got, ref):
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    return float((got - ref).abs().max()) / max(float(ref.abs().max()), 1e-30)


def record(name, value):
    path = os.path.join(ROOT, "gpurun_out", "precision_errors.json")
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        data = json.load(open(path)) if os.path.exists(path) else {}
        data[name] = value
        json.dump(data, open(path, "w"), indent=1, sort_keys=True)
    except OSError:
        pass


@pytest.fixture(autouse=True)
def _restore_precision():
    from torchctr_b200.nn import set_matmul_precision
    old = torch.backends.cuda.matmul.allow_tf32
    yield
    set_matmul_precision(None)
    torch.backends.cuda.matmul.allow_tf32 = old


def no_dropout(model):
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    return model


def test_split_tf32_parts_and_layouts():
    from torchctr_b200 import ops
    gen = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(37, 21, device="cuda", generator=gen) * torch.logspace(-6, 6, 21, device="cuda")
    seg = 24
    for role in (0, 1):
        a = ops.split_tf32(x, 1, role)
        assert a.shape == (37, 3 * seg)
        b = ops.split_tf32(x, 0, role)
        assert b.shape == (3 * 37, seg)
        parts_a = [a[:, k * seg:k * seg + 21] for k in range(3)]
        parts_b = [b[k * 37:(k + 1) * 37, :21] for k in range(3)]
        for parts in (parts_a, parts_b):
            hi = parts[2]                                          # role 0: lo | hi | hi, role 1: hi | lo | hi
            lo = parts[0] if role == 0 else parts[1]
            other = parts[1] if role == 0 else parts[0]
            assert torch.equal(hi, other)
            for t in (hi, lo):                                     # TF32-representable: low 13 mantissa bits clear
                assert int((t.contiguous().view(torch.int32) & 0x1FFF).abs().max()) == 0
            # hi + lo reproduces x to 2^-21 relative (two 11-bit roundings)
            assert float(((hi.double() + lo.double() - x.double()).abs() / x.double().abs().clamp_min(1e-30)).max()) < 2.0 ** -20
        assert float(a[:, 21:seg].abs().max()) == 0 and float(b[:, 21:].abs().max()) == 0     # padding columns are zero


@pytest.mark.parametrize("M,N,K", [(4096, 256, 432), (1000, 128, 256), (513, 64, 128), (2048, 848, 848), (300, 17, 19)])
def test_linear_tf32x3_meets_fp32_bound(M, N, K):
    """Forward, input gradient and weight gradient of ``linear_tc`` in the exact mode against fp64: 1e-5 of max|ref|
    (measured ~1e-6, what an fp32 GEMM gives), and it beats plain TF32 by two orders of magnitude."""
    from torchctr_b200.nn.linear import linear_tc
    gen = torch.Generator(device="cuda").manual_seed(M + N + K)
    x = torch.randn(M, K, device="cuda", generator=gen, requires_grad=True)
    w = (torch.randn(N, K, device="cuda", generator=gen) / K ** 0.5).requires_grad_(True)
    b = torch.randn(N, device="cuda", generator=gen, requires_grad=True)
    gy = torch.randn(M, N, device="cuda", generator=gen)
    ref_y = x.detach().double() @ w.detach().double().t() + b.detach().double()
    ref_gx = gy.double() @ w.detach().double()
    ref_gw = gy.double().t() @ x.detach().double()
    errs = {}
    for prec in ("tf32x3", "tf32"):
        if prec == "tf32" and (K % 4 or N % 4):
            continue
        x.grad = w.grad = b.grad = None
        y = linear_tc(x, w, b, precision=prec)
        y.backward(gy)
        errs[prec] = (rel_err(y, ref_y), rel_err(x.grad, ref_gx), rel_err(w.grad, ref_gw))
    record(f"linear_{M}x{N}x{K}", errs)
    assert max(errs["tf32x3"]) < 1e-5, errs
    if "tf32" in errs:
        assert max(errs["tf32"]) < 2e-3 and min(errs["tf32"]) > 20 * max(errs["tf32x3"]), errs


def _criteo_shape(gen, B, F, V, D, nd):
    fc = [{"name": f"c{i}", "type": "sparse", "num_embeddings": V + 7 * i, "emb_dim": D} for i in range(F)]
    fc += [{"name": f"d{i}", "type": "dense"} for i in range(nd)]
    # Zipf-like ids: many repeats of the hot rows, as in the bench
    feats = {}
    for i in range(F):
        u = torch.rand(B, 1, generator=gen)
        feats[f"c{i}"] = ((V + 7 * i) ** u - 1).long().clamp_(0, V + 7 * i - 1)
    feats["dense_features"] = torch.randn(B, nd, generator=gen)
    labels = (torch.rand(B, 1, generator=gen) < 0.25).float()
    return fc, feats, labels


def _build(which, fc, hidden, device):
    from oracle import models as om
    from torchctr_b200.models import DCNv2, DeepFM
    torch.manual_seed(0)
    ref = no_dropout(om.OracleDeepFM(fc, hidden) if which == "deepfm" else om.OracleDCNv2(fc, hidden, 3))
    ours = no_dropout(DeepFM(fc, hidden) if which == "deepfm" else DCNv2(fc, hidden, 3)).to(device)
    with torch.no_grad():      # trained-size embeddings (the N(0, 1) init makes DeepFM logits ~ +-30 and lets the cross
        for n, p in ref.named_parameters():     # network's x0 * (W x + b) + x grow to 1e3: fp32 itself is then only good to 1e-2)
            if n.startswith(("embeddings", "linear_embeddings")):
                p.mul_(0.1)
    ours.load_state_dict(ref.state_dict())
    return ref.train(), ours.train()


def _biases_in_front_of_batchnorm(model):
    names = set()
    layers = list(model.tower)
    for i, m in enumerate(layers[:-1]):
        if isinstance(m, torch.nn.Linear) and isinstance(layers[i + 1], torch.nn.BatchNorm1d):
            names.add(f"tower.{i}.bias")
    return names


def _is_table(name):
    return name.startswith(("embeddings", "linear_embeddings"))


def _step_errors(ref, ours, batches, lr, ref_device="cpu", tag=""):
    """Runs the same SGD steps on both (``ours`` may be a second oracle: the noise-floor measurement).  Returns the worst
    relative error (to max|ref| of the tensor) of the loss, of the dense gradients (every step) and of the table-row
    updates after the last step (row_after - row_init: the embedding gradients as the fused update applied them)."""
    init = {k: v.detach().clone().cpu() for k, v in ref.state_dict().items() if _is_table(k) and k.endswith("weight")}
    opt_ref = torch.optim.SGD(ref.parameters(), lr=lr)
    opt = torch.optim.SGD(ours.parameters(), lr=lr)
    if hasattr(ours, "bind_optimizer"):
        ours.bind_optimizer(opt)
    dev = next(ours.parameters()).device
    e_loss = e_grad = 0.0
    worst = {}
    bn_biases = _biases_in_front_of_batchnorm(ref)
    for step, (feats, labels) in enumerate(batches):
        rf = {k: v.to(ref_device) for k, v in feats.items()}
        of = {k: v.to(dev) for k, v in feats.items()}
        opt_ref.zero_grad(); opt.zero_grad()
        l_ref = ref.training_step((rf, labels.to(ref_device)), step)
        l = ours.training_step((of, labels.to(dev)), step)
        e_loss = max(e_loss, rel_err(l, l_ref))
        l_ref.backward(); l.backward()
        ref_grads = dict(ref.named_parameters())
        for name, p in ours.named_parameters():
            if p.grad is None or _is_table(name) or name in bn_biases:   # bias in front of BatchNorm: mathematically zero
                continue
            e = rel_err(p.grad, ref_grads[name].grad)
            if e > e_grad:
                e_grad, worst["grad"] = e, f"{name}@step{step}"
        opt_ref.step(); opt.step()
    torch.cuda.synchronize()
    e_rows = 0.0
    sd, sd_ref = ours.state_dict(), ref.state_dict()
    for k, w0 in init.items():
        e = rel_err(sd[k].detach().cpu() - w0, sd_ref[k].detach().cpu() - w0)
        if e > e_rows:
            e_rows, worst["rows"] = e, k
    record(f"worst_{tag}", worst)
    return e_loss, e_grad, e_rows


@pytest.mark.parametrize("precision", ["tf32x3", "tf32"])
@pytest.mark.parametrize("which", ["deepfm", "dcnv2"])
def test_models_at_bench_widths(which, precision):
    from torchctr_b200.nn import set_matmul_precision
    set_matmul_precision(precision)
    torch.backends.cuda.matmul.allow_tf32 = precision == "tf32"
    gen = torch.Generator().manual_seed(11)
    D = 16 if which == "deepfm" else 32
    B = 4096
    fc, f0, l0 = _criteo_shape(gen, B, 26, 3000, D, 13)
    batches = [(f0, l0)]
    for _ in range(2):
        _, f, l = _criteo_shape(gen, B, 26, 3000, D, 13)
        batches.append((f, l))
    hidden = [256, 128, 64]
    ref, ours = _build(which, fc, hidden, "cuda")
    ours_err = _step_errors(ref, ours, batches, lr=0.05, tag=f"{which}_{precision}_vs_cpu")
    # the yardstick: torch's own CUDA path in the same precision mode (oracle modules on the GPU: cuBLAS fp32 when
    # precision is tf32x3, cuBLAS TF32 when it is tf32) against the same CPU fp32 oracle
    ref_b, _ = _build(which, fc, hidden, "cuda")
    ref_c, _ = _build(which, fc, hidden, "cuda")
    torch_err = _step_errors(ref_b, ref_c.cuda(), batches, lr=0.05, tag=f"{which}_{precision}_torch_cuda_vs_cpu")
    record(f"{which}_{precision}", {"ours_vs_cpu_fp32_oracle": ours_err, "torch_cuda_same_mode_vs_cpu_fp32_oracle": torch_err})
    tol = TOL[precision]
    for i, key in enumerate(("loss", "grad", "rows")):
        # within the stated bound, or -- where fp32 itself is not that good on this quantity (batch reductions with
        # cancellation behind BatchNorm) -- within 3x of what torch's own CUDA path achieves
        assert ours_err[i] <= max(tol[key], 3.0 * torch_err[i]), (key, ours_err, torch_err)


@pytest.mark.parametrize("precision", ["tf32x3", "tf32"])
def test_dnn_reference_golden_in_both_precisions(golden_dir, precision):
    """The fixture recorded from the REAL reference (torchctr.models.DNN + torch.optim.Adagrad) with the tower on the
    tcgen05 kernels in the exact and in the benchmarked precision."""
    from torchctr_b200.models import DNN
    from torchctr_b200.nn import set_matmul_precision
    set_matmul_precision(precision)
    tol = TOL[precision]
    g = torch.load(os.path.join(golden_dir, "dnn_golden.pt"))
    model = no_dropout(DNN(g["feat_configs"], g["hidden_units"])).cuda()
    model.load_state_dict(g["init_state"])
    model.eval()
    with torch.no_grad():
        e_eval = rel_err(model(g["feats"]), g["eval_logits"])
    model.train()
    opt = torch.optim.Adagrad(model.parameters(), lr=g["adagrad_lr"])
    model.bind_optimizer(opt)
    opt.zero_grad()
    loss = model.training_step((g["feats"], g["labels"]), 0)
    e_loss = rel_err(loss, g["train_loss"])
    loss.backward()
    noise = {n for n, gr in g["grads"].items() if float(gr.abs().max()) < 1e-6}
    per = {n: rel_err(p.grad, g["grads"][n]) for n, p in model.named_parameters() if n.startswith("tower") and n not in noise}
    e_grad = max(per.values())
    record(f"worst_dnn_golden_{precision}", max(per, key=per.get))
    opt.step()
    after = model.state_dict()
    e_state = max(rel_err(after[k], ref) for k, ref in g["after_adagrad_step"].items() if k not in noise)
    record(f"dnn_golden_{precision}", dict(eval_logits=e_eval, loss=e_loss, grad=e_grad, state=e_state))
    # yardstick: the oracle's DNN (pinned to the reference by this same fixture) on torch's own CUDA path in the same mode
    from oracle import models as om
    torch.backends.cuda.matmul.allow_tf32 = precision == "tf32"
    twin = no_dropout(om.OracleDNN(g["feat_configs"], g["hidden_units"])).cuda()
    twin.load_state_dict(g["init_state"])
    twin.train()
    cf = {k: v.cuda() for k, v in g["feats"].items()}
    l2 = twin.training_step((cf, g["labels"].cuda()), 0)
    l2.backward()
    t_loss = rel_err(l2, g["train_loss"])
    t_grad = max(rel_err(p.grad, g["grads"][n]) for n, p in twin.named_parameters() if n.startswith("tower") and n not in noise)
    record(f"dnn_golden_{precision}_torch_cuda_same_mode", dict(loss=t_loss, grad=t_grad))
    assert e_loss <= max(tol["loss"], 3 * t_loss) and e_eval <= tol["grad"]
    if precision == "tf32x3":
        assert e_grad <= tol["grad"]
    # tf32: this fixture has B = 64 -- ONE ReLU whose pre-activation sits within TF32 round-off of zero flips and moves
    # that column's BatchNorm-beta / first-layer weight gradient by 10 % or more (measured, profiles/micro/
    # diag_golden_tf32.py: every other tensor is within 6e-4, like torch's own TF32 path).  Gradients under TF32 are
    # therefore checked at B = 4096 in test_models_at_bench_widths; here the value is only recorded.
    if precision == "tf32x3":      # Adagrad's first step is lr * sign(g): only meaningful at fp32-grade gradient error
        assert e_state <= 5e-4
