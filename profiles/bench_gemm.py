"""Micro-benchmark of the tcgen05 TF32 GEMM (ctr_linear_fwd) against torch/cuBLAS TF32 on the tower and
cross-layer shapes of the BASELINE configs.  Run on a B200: python profiles/bench_gemm.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchctr_b200 import ops  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = True
shapes = [(65536, 256, 432), (65536, 128, 256), (65536, 64, 128), (65536, 432, 256), (65536, 848, 848)]
res = []
for M, N, K in shapes:
    A = torch.randn(M, K, device="cuda")
    W = torch.randn(N, K, device="cuda") / K ** 0.5
    b = torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda")

    def time(fn, iters=30):
        for _ in range(5):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    t_ours = time(lambda: ops.linear_fwd(A, W, b, 0, out))
    t_torch = time(lambda: torch.nn.functional.linear(A, W, b))
    fl = 2.0 * M * N * K
    res.append({"M": M, "N": N, "K": K, "ours_us": 1e3 * t_ours, "torch_tf32_us": 1e3 * t_torch,
                "ours_TFLOPs": fl / t_ours / 1e9, "torch_TFLOPs": fl / t_torch / 1e9,
                "min_hbm_us": 1e6 * 4 * (M * K + M * N + N * K) / 6.5395e12})
    print(json.dumps(res[-1]))
