"""Diagnostic (GPU): per-parameter gradient error of the golden DNN case in TF32 -- ours vs torch's cuBLAS TF32 --
and the three GEMMs of the first tower block on the very same tensors against fp64."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import models as om
from torchctr_b200 import ops
from torchctr_b200.models import DNN
from torchctr_b200.nn import set_matmul_precision

g = torch.load(os.path.join(ROOT, "tests/golden/dnn_golden.pt"))
def nd(m):
    for x in m.modules():
        if isinstance(x, torch.nn.Dropout): x.p = 0.0
    return m
def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max()) / max(float(b.abs().max()), 1e-30)
def grads(model, feats, labels):
    model.zero_grad(); l = model.training_step((feats, labels), 0); l.backward()
    return {n: p.grad for n, p in model.named_parameters() if n.startswith("tower")}, l
cf = {k: v.cuda() for k, v in g["feats"].items()}
res = {}
for mode in ("tf32x3", "tf32"):
    set_matmul_precision(mode)
    m = nd(DNN(g["feat_configs"], g["hidden_units"])).cuda(); m.load_state_dict(g["init_state"]); m.train()
    res["ours_" + mode], _ = grads(m, g["feats"], g["labels"])
torch.backends.cuda.matmul.allow_tf32 = True
t = nd(om.OracleDNN(g["feat_configs"], g["hidden_units"])).cuda(); t.load_state_dict(g["init_state"]); t.train()
res["torch_tf32"], _ = grads(t, cf, g["labels"].cuda())
for n in g["grads"]:
    if n.startswith("tower"):
        print(f"{n:16s}", {k: f"{rel(v[n], g['grads'][n]):.2e}" for k, v in res.items()})
# GEMM-level on random tensors of the same shapes
gen = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(64, 60, device="cuda", generator=gen) * 3; w = torch.randn(32, 60, device="cuda", generator=gen) / 8
gz = torch.randn(64, 32, device="cuda", generator=gen)
print("fwd   ", rel(ops.linear_fwd(x, w), x.double() @ w.double().t()), rel(torch.nn.functional.linear(x, w), x.double() @ w.double().t()))
print("dgrad ", rel(ops.linear_fwd(gz, w.t().contiguous()), gz.double() @ w.double()), rel(gz @ w, gz.double() @ w.double()))
print("wgrad ", rel(ops.linear_wgrad(gz, x), gz.double().t() @ x.double()), rel(gz.t() @ x, gz.double().t() @ x.double()))
