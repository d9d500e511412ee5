import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from torchctr_b200 import ops
M, N, K = (int(x) for x in (sys.argv[1:4] if len(sys.argv) > 3 else (65536, 256, 432)))
A = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") / K ** 0.5; b = torch.randn(N, device="cuda")
out = torch.empty(M, N, device="cuda")
for _ in range(6):
    ops.linear_fwd(A, W, b, 0, out)
torch.cuda.synchronize()
print("ok")
