"""Oracle definitions of the feature-level style-transfer operators (AdaIN + Gram loss).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

PARITY UNPINNED BY THE REFERENCE: the reference has no AdaIN and no Gram-matrix loss -- at the
site where these operators are inserted (train_and_test.py:552-561) it runs a WaveGlow flow
plus NoiseTransfer.  The definitions below are the frozen spec from
SURVEY.md section 8c / appendix A3-A4; they are pinned against torch autograd in fp64 by
tests/test_oracle.py, and the tensors they act on (``target_feature`` and
``source_shape_changed_feature``, both [B, C, L]) are the ones the reference produces at
train_and_test.py:547-550.
"""
from __future__ import annotations

import torch

ADAIN_EPS = 1e-5


def row_stats(x: torch.Tensor):
    """Per (b, c) row over L: mean and unbiased variance (the Welford kernel's contract).
    The same reduction along dim 0 is the reference's batch mean, widgets.py:155-159."""
    return x.mean(-1), x.var(-1, unbiased=True)


def adain(content: torch.Tensor, style: torch.Tensor, eps: float = ADAIN_EPS) -> torch.Tensor:
    """out = (content - mu_c) / sigma_c * sigma_s + mu_s, sigma = sqrt(var_unbiased + eps);
    rows paired by (b, c) as the reference pairs batches (train_and_test.py:540-541)."""
    mc, vc = row_stats(content)
    ms, vs = row_stats(style)
    sc, ss = torch.sqrt(vc + eps), torch.sqrt(vs + eps)
    return (content - mc[..., None]) / sc[..., None] * ss[..., None] + ms[..., None]


def adain_backward(dy, content, style, eps: float = ADAIN_EPS):
    """Closed form (appendix A3): returns (dcontent, dstyle)."""
    L = content.shape[-1]
    mc, vc = row_stats(content)
    ms, vs = row_stats(style)
    sc, ss = torch.sqrt(vc + eps), torch.sqrt(vs + eps)
    xh = (content - mc[..., None]) / sc[..., None]
    sh = (style - ms[..., None]) / ss[..., None]
    s1 = dy.sum(-1, keepdim=True)
    s2 = (dy * xh).sum(-1, keepdim=True)
    dcontent = (ss / sc)[..., None] * (dy - s1 / L - xh * s2 / (L - 1))
    dstyle = s1 / L + sh * s2 / (L - 1)
    return dcontent, dstyle


def gram(x: torch.Tensor) -> torch.Tensor:
    """G(x) = x x^T / (C L) per sample -> [B, C, C]."""
    _, C, L = x.shape
    return torch.bmm(x, x.transpose(1, 2)) / (C * L)


def gram_style_loss(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """mean over B*C*C of (G(a) - G(b))^2."""
    return ((gram(a) - gram(b)) ** 2).mean()


def gram_style_loss_backward(a, b):
    """Closed form (appendix A4) for upstream gradient 1: da = 4 D a / (B C^3 L), db = -4 D b / (B C^3 L)."""
    B, C, L = a.shape
    D = gram(a) - gram(b)
    k = 4.0 / (B * C * C * C * L)
    return k * torch.bmm(D, a), -k * torch.bmm(D, b)
