"""How fast is the weight-gradient contraction gW[N, K] = sum_b gz[b, N] x[b, K] through torch / cuBLAS (TF32)
in its possible formulations?  Decides whether a hand-written split-K tcgen05 kernel is worth it."""
import torch
torch.backends.cuda.matmul.allow_tf32 = True
B = 65536


def t(fn, iters=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(iters):
        fn()
    e1.record(); torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / iters


for N, K in ((256, 432), (128, 256), (64, 128)):
    gz = torch.randn(B, N, device="cuda"); x = torch.randn(B, K, device="cuda")
    gzT = gz.t().contiguous(); xT = x.t().contiguous()
    res = {
        "gz.t() @ x": t(lambda: gz.t() @ x),
        "(x.t() @ gz).t()": t(lambda: (x.t() @ gz).t()),
        "gzT @ x (A contiguous)": t(lambda: gzT @ x),
        "gzT @ xT.t() (both K-major)": t(lambda: gzT @ xT.t()),
        "transposes only": t(lambda: (gz.t().contiguous(), x.t().contiguous())),
        "bf16 gz.t() @ x": t(lambda: gz.bfloat16().t() @ x.bfloat16()),
        "min HBM us": 1e6 * 4 * B * (N + K) / 6.5e12,
    }
    print(N, K, {k: round(v, 1) for k, v in res.items()})
