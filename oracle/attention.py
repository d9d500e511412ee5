"""Oracle (test infrastructure): ``target_attention`` restated from ``/root/reference/torchctr/nn/functional.py:46-74``.

Pinned by ``tests/golden/attention_golden.pt`` (outputs and gradients of the imported reference, with and without a mask).
``honor_mask=False`` follows the reference literally -- ``attn_scores.masked_fill(mask == 0, -inf)`` at ``:63`` returns a new
tensor that is dropped, so the mask does nothing; ``honor_mask=True`` is the evident intent."""
import torch
import torch.nn.functional as F


def target_attention(target_emb, candidate_embs, mask=None, honor_mask=False):
    scores = torch.matmul(candidate_embs, target_emb.unsqueeze(-1)).squeeze(-1)      # :60
    if mask is None or not honor_mask:
        weights = F.softmax(scores, dim=1)                                             # :63 is a no-op upstream, :66
    else:
        dead = (mask == 0).all(dim=1, keepdim=True)                                    # a fully masked row pools to zero
        scores = scores.masked_fill(mask == 0, float("-inf"))                          # :63, as intended
        scores = torch.where(dead, torch.zeros_like(scores), scores)
        weights = F.softmax(scores, dim=1) * (~dead).to(scores.dtype)
    return (weights.unsqueeze(-1) * candidate_embs).sum(dim=1)                         # :69-72
