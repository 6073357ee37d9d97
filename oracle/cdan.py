"""Oracle of the conditional-adversarial (C-DAN) consumer of the transferred features, on CPU.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Functional restatement (state dicts + explicit gradient-reversal nodes instead of tensor hooks) of

* ``RandomLayer``                    C_DAN.py:11-25   fixed Gaussian projections, product of the projected views
* ``Entropy`` / ``calc_coeff``       C_DAN.py:32-45, widgets.py:12-13
* ``AdversarialNetworkforCDAN``      widgets.py:81-131 critic MLP with a gradient-reversal input and a warm-up schedule
* ``CDAN``                           C_DAN.py:49-82   entropy-weighted Wasserstein-style distance

**Pinned** against the unmodified reference classes imported from /root/reference: ``oracle/make_golden.py`` wrote
``tests/golden/cdan_small.npz`` (seeded inputs, the reference's random matrices / critic parameters, loss, every
gradient); ``tests/test_oracle.py`` checks this file against it.  Dropout (p = 0.2 in the critic) draws from the
device's generator, so seeded CPU and CUDA runs cannot agree on its mask: the pinned cases run the critic with
p = 0 (training mode, so the schedule advances) and in eval mode.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Sequence

import torch
import torch.nn.functional as F

RANDOM_DIM = 1024            # RandomLayer(output_dim=1024); train_and_test.py:74-76
ENTROPY_EPS = 1e-5           # C_DAN.py:34


def calc_coeff(iter_num, high=1.0, low=0.0, alpha=100.0, max_iter=20.0) -> float:
    """widgets.py:12-13 with the critic's own constants (widgets.py:108-112)."""
    return float(2.0 * (high - low) / (1.0 + math.exp(-alpha * iter_num / max_iter)) - (high - low) + low)


class _Reverse(torch.autograd.Function):
    """identity forward, ``-coeff * g`` backward (what ``register_hook(grl_hook(coeff))`` does, C_DAN.py:38-41)."""

    @staticmethod
    def forward(ctx, x, coeff):
        ctx.coeff = coeff
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return -ctx.coeff * g, None


def init_random_layer(dims: Sequence[int], output_dim: int = RANDOM_DIM) -> List[torch.Tensor]:
    """C_DAN.py:16: one ``torch.randn(d, output_dim)`` per view, in order, from the global CPU generator."""
    return [torch.randn(d, output_dim) for d in dims]


def random_layer(mats: List[torch.Tensor], views: List[torch.Tensor]) -> torch.Tensor:
    """C_DAN.py:20-25: prod_i (view_i @ R_i), the first factor scaled by output_dim ** (-1 / n_views)."""
    outs = [v @ m for v, m in zip(views, mats)]
    res = outs[0] / math.pow(float(mats[0].shape[1]), 1.0 / len(outs))
    for o in outs[1:]:
        res = res * o
    return res


def init_ad_net(in_feature: int, hidden: int) -> "OrderedDict[str, torch.Tensor]":
    """widgets.py:96-106: three nn.Linear (each draws its default init), then ``apply(init_weights)`` re-draws the
    weights with xavier_normal_ in registration order and zeroes the biases."""
    lins = [torch.nn.Linear(in_feature, hidden), torch.nn.Linear(hidden, hidden), torch.nn.Linear(hidden, 1)]
    sd = OrderedDict()
    for i, lin in enumerate(lins, 1):
        torch.nn.init.xavier_normal_(lin.weight)
        sd[f"ad_layer{i}.weight"] = lin.weight.detach().clone()
        sd[f"ad_layer{i}.bias"] = torch.zeros_like(lin.bias)
    return sd


class AdNetState:
    """The critic's warm-up schedule (widgets.py:107-119): ``iter_num`` advances once per training-mode call and
    saturates at ``max_iter``; ``coeff`` is the value of the LAST call (CDAN reads it after both calls)."""

    def __init__(self):
        self.iter_num = -1
        self.max_iter = 20.0
        self.coeff = 0.001

    def advance(self, training: bool) -> float:
        if training:
            self.iter_num += 1
        if self.iter_num >= self.max_iter:
            self.iter_num = self.max_iter
        self.coeff = calc_coeff(self.iter_num, 1.0, 0.0, 100.0, self.max_iter)
        return self.coeff


def ad_net_forward(sd: Dict[str, torch.Tensor], state: AdNetState, x: torch.Tensor, training: bool = True,
                   dropout_p: float = 0.0) -> torch.Tensor:
    """widgets.py:113-130."""
    coeff = state.advance(training)
    h = _Reverse.apply(x * 1.0, coeff)
    h = F.relu(F.linear(h, sd["ad_layer1.weight"], sd["ad_layer1.bias"]))
    h = F.dropout(h, dropout_p, training)
    h = F.relu(F.linear(h, sd["ad_layer2.weight"], sd["ad_layer2.bias"]))
    h = F.dropout(h, dropout_p, training)
    return F.linear(h, sd["ad_layer3.weight"], sd["ad_layer3.bias"])


def entropy(p: torch.Tensor) -> torch.Tensor:
    """C_DAN.py:32-37 (input already soft-maxed)."""
    return torch.sum(-p * torch.log(p + ENTROPY_EPS), dim=1)


def cdan(feat_t: torch.Tensor, feat_s2t: torch.Tensor, logits_t: torch.Tensor, logits_s2t: torch.Tensor,
         ad_sd: Dict[str, torch.Tensor], ad_state: AdNetState, mats: List[torch.Tensor], training: bool = True,
         dropout_p: float = 0.0) -> torch.Tensor:
    """C_DAN.py:49-82 (random-layer branch): distance_target - distance_generated.

    Note the shapes the reference really multiplies: the weights are [B], the critic outputs [B, 1] (the
    ``.view(-1, 1)`` results at lines 74/76 are discarded), so ``weight * out`` broadcasts to [B, B] and the sum is
    ``sum(w) * sum(out)`` with ``sum(w) == 1`` up to rounding."""
    ft, fs = torch.flatten(feat_t, 1), torch.flatten(feat_s2t, 1)
    pt, ps = F.softmax(logits_t, dim=1), F.softmax(logits_s2t, dim=1)
    out_t = ad_net_forward(ad_sd, ad_state, random_layer(mats, [ft, pt]), training, dropout_p)
    out_s = ad_net_forward(ad_sd, ad_state, random_layer(mats, [fs, ps]), training, dropout_p)
    coeff = ad_state.coeff
    ht = _Reverse.apply(entropy(pt), coeff)
    hs = _Reverse.apply(entropy(ps), coeff)
    wt = 1.0 + torch.exp(-ht)
    ws = 1.0 + torch.exp(-hs)
    wt = wt / torch.sum(wt).detach()
    ws = ws / torch.sum(ws).detach()
    return torch.sum(wt * out_t) - torch.sum(ws * out_s)
