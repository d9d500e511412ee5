import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from torchctr_b200 import ops
def rel(a, b):
    return float((a.double() - b).abs().max()) / max(float(b.abs().max()), 1e-30)
gen = torch.Generator(device="cuda").manual_seed(0)
for M in (64, 200):
    for N in (16, 32, 60):
        for K in (16, 32, 48, 60):
            a = torch.randn(M, K, device="cuda", generator=gen); w = torch.randn(N, K, device="cuda", generator=gen)
            b = torch.randn(N, device="cuda", generator=gen)
            ref = a.double() @ w.double().t() + b.double()
            print("fwd", M, N, K, f"{rel(ops.linear_fwd(a, w, b), ref):.2e}", end=" | ")
            g = torch.randn(M, N, device="cuda", generator=gen)
            if N % 4 == 0 and K % 4 == 0:
                print("wgrad", f"{rel(ops.linear_wgrad(g, a), g.double().t() @ a.double()):.2e}", end=" | ")
                # column sums of a zero-column-sum matrix through the GEMM (what BatchNorm backward feeds the next dgrad)
                g0 = g - g.mean(0, keepdim=True)
                out = ops.linear_fwd(g0, w.t().contiguous())
                refo = g0.double() @ w.double()
                print("colsum", f"{float((out.double().sum(0) - refo.sum(0)).abs().max()) / float(refo.abs().sum(0).max()):.2e}")
            else:
                print()
