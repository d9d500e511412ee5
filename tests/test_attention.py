"""f4: ``target_attention`` as a fused masked-softmax pooling kernel, against the REAL reference
(tests/golden/attention_golden.pt: outputs + gradients of torchctr.nn.functional.target_attention) and, for the fixed mask
semantics the reference intends but does not implement (functional.py:63), against the oracle."""
import os

import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "attention_golden.pt")


def close(a, b, tol=1e-5):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    assert float((a - b).abs().max()) <= tol * max(float(b.abs().max()), 1e-30), float((a - b).abs().max())


def test_oracle_matches_reference_golden_and_mask_is_ignored_upstream():
    from oracle.attention import target_attention
    for c in torch.load(GOLDEN):
        assert c["mask_ignored_by_reference"]                        # the latent bug, recorded from the reference itself
        for mask in (None, c["mask"]):
            t = c["target"].clone().requires_grad_(True); k = c["cand"].clone().requires_grad_(True)
            out = target_attention(t, k, mask)
            out.backward(c["gout"])
            close(out, c["out"], 1e-6); close(t.grad, c["gtarget"], 1e-6); close(k.grad, c["gcand"], 1e-6)


@pytest.mark.gpu
def test_kernel_matches_reference_golden():
    from torchctr_b200.nn.functional import target_attention
    for c in torch.load(GOLDEN):
        for mask in (None, c["mask"]):
            t = c["target"].cuda().requires_grad_(True); k = c["cand"].cuda().requires_grad_(True)
            out = target_attention(t, k, None if mask is None else mask.cuda())      # default: the reference's behaviour
            out.backward(c["gout"].cuda())
            close(out, c["out"]); close(t.grad, c["gtarget"]); close(k.grad, c["gcand"])


@pytest.mark.gpu
@pytest.mark.parametrize("B,N,E", [(300, 200, 16), (1000, 50, 32), (64, 7, 128), (513, 33, 4), (10, 1, 8), (5, 0, 16)])
def test_kernel_with_the_mask_applied_matches_oracle(B, N, E):
    from oracle.attention import target_attention as ref_fn
    from torchctr_b200.nn.functional import target_attention
    gen = torch.Generator().manual_seed(B + N + E)
    tgt = torch.randn(B, E, generator=gen); cand = torch.randn(B, N, E, generator=gen)
    lens = torch.randint(0, N + 1, (B, 1), generator=gen)             # some rows have no valid candidate at all
    mask = (torch.arange(N)[None, :] < lens).float()
    gout = torch.randn(B, E, generator=gen)
    tr = tgt.clone().requires_grad_(True); kr = cand.clone().requires_grad_(True)
    ref = ref_fn(tr, kr, mask, honor_mask=True)
    ref.backward(gout)
    t = tgt.cuda().requires_grad_(True); k = cand.cuda().requires_grad_(True)
    out = target_attention(t, k, mask.cuda(), honor_mask=True)
    out.backward(gout.cuda())
    close(out, ref, 2e-5); close(t.grad, tr.grad, 2e-5)
    if N:
        close(k.grad, kr.grad, 2e-5)
        assert float(k.grad.cpu()[mask == 0].abs().max() if (mask == 0).any() else 0.0) == 0.0      # masked candidates get no gradient
