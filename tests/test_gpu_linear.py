"""tcgen05 GEMM (K6) against torch: TF32 inputs / fp32 accumulate, so the tolerance is TF32's
(10-bit mantissa): 2e-3 of max|ref| against an fp32 reference, and it must agree with torch's own TF32
path to the same bound."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel_err(got, ref):
    return float((got - ref).abs().max()) / max(float(ref.abs().max()), 1e-30)


@pytest.mark.parametrize("M,N,K,lda_pad", [(128, 128, 32, 0), (1000, 256, 432, 0), (4096, 128, 256, 0), (777, 64, 128, 0),
                                           (300, 848, 848, 0), (65536, 256, 429, 3), (513, 16, 64, 0)])
@pytest.mark.parametrize("act", [0, 1])
def test_linear_fwd_matches_torch(M, N, K, lda_pad, act):
    from torchctr_b200 import ops
    gen = torch.Generator(device="cuda").manual_seed(M + N + K)
    buf = torch.zeros(M, K + lda_pad, device="cuda")
    buf[:, :K] = torch.randn(M, K, device="cuda", generator=gen)
    A = buf[:, :K] if lda_pad else buf
    Wfull = torch.zeros(N, K + lda_pad, device="cuda")
    Wfull[:, :K] = torch.randn(N, K, device="cuda", generator=gen) / K ** 0.5
    W = Wfull[:, :K] if lda_pad else Wfull
    bias = torch.randn(N, device="cuda", generator=gen)
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = A.double() @ W.double().t() + bias.double()
    if act:
        ref = ref.clamp(min=0)
    got = ops.linear_fwd(A, W, bias, act)
    torch.cuda.synchronize()
    assert rel_err(got.double(), ref) < 2e-3
    torch.backends.cuda.matmul.allow_tf32 = True
    tf32 = torch.nn.functional.linear(A.contiguous(), W.contiguous(), bias)
    torch.backends.cuda.matmul.allow_tf32 = False
    if act:
        tf32 = tf32.clamp(min=0)
    assert rel_err(got, tf32) < 2e-3


def test_linear_tc_autograd_matches_torch():
    from torchctr_b200.nn.linear import linear_tc
    gen = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn(2048, 432, device="cuda", generator=gen, requires_grad=True)
    w = (torch.randn(256, 432, device="cuda", generator=gen) / 20).requires_grad_(True)
    b = torch.randn(256, device="cuda", generator=gen, requires_grad=True)
    gy = torch.randn(2048, 256, device="cuda", generator=gen)
    y = linear_tc(x, w, b)
    y.backward(gy)
    got = (y.detach(), x.grad.clone(), w.grad.clone(), b.grad.clone())
    x.grad = w.grad = b.grad = None
    yr = torch.nn.functional.linear(x, w, b)
    yr.backward(gy)
    for g, r in zip(got, (yr.detach(), x.grad, w.grad, b.grad)):
        assert rel_err(g, r) < 2e-3


@pytest.mark.parametrize("B,N,K", [(65536, 256, 432), (65536, 128, 256), (65536, 64, 128), (1000, 64, 128), (70, 48, 20),
                                   (4097, 300, 520), (31, 8, 4)])
def test_wgrad_matches_fp32_matmul(B, N, K):
    """dW = gz^T x on tcgen05 with MN-major operands and a batch split: TF32 inputs (10-bit mantissa) against fp32."""
    from torchctr_b200 import ops
    gen = torch.Generator().manual_seed(B + N + K)
    gz = torch.randn(B, N, generator=gen).cuda()
    x = torch.randn(B, K, generator=gen).cuda()
    got = ops.linear_wgrad(gz, x)
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = gz.double().t() @ x.double()
    # TF32 operands (10-bit mantissa): same bound as the forward GEMM tests, 2e-3 of max|ref|
    assert got.shape == (N, K)
    assert rel_err(got.double(), ref) < 2e-3
    # strided views: rows padded to a multiple of 4 floats, as the lookup output is
    xp = torch.zeros(B, (K + 7) // 4 * 4 + 4, device="cuda")
    xp[:, :K] = x
    got2 = ops.linear_wgrad(gz, xp[:, :K])
    assert torch.equal(got, got2)


@pytest.mark.parametrize("M,N,K,D", [(4096, 416, 256, 16), (1000, 128, 64, 32), (300, 192, 40, 64), (513, 64, 128, 16)])
def test_linear_column_blocked_output_equals_row_major(M, N, K, D):
    """ctr_linear_fwd_blocked: the same numbers as ctr_linear_fwd, columns [j D, (j + 1) D) stored as their own contiguous
    [M, D] matrix -- the layout the embedding update reads dL/d(pooled output) in (ctr_group_t.grad_blocked)."""
    from torchctr_b200 import ops
    gen = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=gen).cuda()
    w = torch.randn(N, K, generator=gen).cuda()
    ref = ops.linear_fwd(a, w)
    buf = torch.full((M * N + 64,), float("nan"), device="cuda")
    ops.linear_fwd(a, w, out=buf, out_block=D)
    got = buf[:M * N].view(N // D, M, D).permute(1, 0, 2).reshape(M, N)
    assert torch.equal(got, ref)
    assert torch.isnan(buf[M * N:]).all()
    with pytest.raises(ValueError):
        ops.linear_fwd(a[:100], w, out=buf, out_block=D)          # needs M > 128 (the CTA-pair kernel)
