"""Oracle of the GradNorm arithmetic of the reference's joint stage, torch/numpy on CPU.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates train_and_test.py:646-766 -- what happens between "the losses of this batch exist" and "every optimizer has
stepped" -- for one side (target or source; the reference runs the same code twice with different loss lists).
**Pinned** against the unmodified source: ``oracle/make_golden.py`` executes lines 500-511 and 646-766 of
``/root/reference/train_and_test.py`` over the reference's own OS-CNN modules for two consecutive batches and stores
the losses, per-loss norms, weight gradients, balanced weights and parameter gradients in
``tests/golden/gradnorm_small.npz``; ``tests/test_oracle_gradnorm.py`` compares.

Two behaviours of the reference that a restatement must keep:

* ``torch.autograd.grad(loss_i, shared.parameters())`` returns the gradient of the *unmasked* big Conv1d weight: it is
  non-zero on the masked taps (OS_CNN.py:68-71 masks ``weight.data``, not the gradient -- SURVEY F4) and those entries
  are part of every ``torch.norm``.
* the graph is "cleared" by zeroing ``.data`` of the balanced weights and calling ``loss_total.backward()`` a second
  time (train_and_test.py:727-741): the balanced part contributes nothing then, but the un-balanced remainder
  (c_cdan * cdan + c_fd * discriminator + c_t * t_sl + c_s * s_sl) is differentiated AGAIN, so what the optimizers see
  is  grad(loss_total) + grad(remainder).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np
import torch

ALPHA = 3                      # train_and_test.py:511
TOTAL_T, TOTAL_S = 7.0, 8.0    # :753-757: the balanced weights are renormalised to these sums
INIT_T, INIT_S = (2.0, 5.0), (2.0, 2.0, 4.0)     # :501-505
LR_T, LR_S = 0.0002, 0.001     # :506-507 (Adam)


def sigmoid_np(x: np.ndarray) -> np.ndarray:
    return 1 / (1 + np.exp(-x))


def remainder_coefficients(cur_epoch: int) -> Tuple[float, float, float, float]:
    """(cdan, feature discriminator, t_sl, s_sl) multipliers of the un-balanced losses, train_and_test.py:668-675."""
    if cur_epoch < 12:
        return 3, 3, 2, 2
    if cur_epoch < 24:
        return 2, 3, 1.8, 1.5
    if cur_epoch < 50:
        return 1.5, 2, 1.8, 1.8
    return 1.5, 1.5, 2.5, 2.5


def side_norms(losses: Sequence[torch.Tensor], weights: torch.Tensor, shared_params: List[torch.Tensor]) -> torch.Tensor:
    """train_and_test.py:685-690: norms[i] = sum over the shared block's parameter tensors of ||w_i * dL_i/dp||_2
    (differentiable in w)."""
    norms = []
    for i, li in enumerate(losses):
        g = torch.autograd.grad(li, shared_params, retain_graph=True)
        norms.append(torch.cat([torch.norm(torch.mul(weights[i], gp)).unsqueeze(0) for gp in g]).sum())
    return torch.stack(norms)


def weight_gradient(norms: torch.Tensor, weights: torch.Tensor, loss_values: np.ndarray, initial: np.ndarray,
                    alpha: float = ALPHA) -> Tuple[torch.Tensor, np.ndarray]:
    """train_and_test.py:693-715: target = mean(norms) * (relative inverse training rate) ** alpha, treated as a constant;
    GradNorm loss = sum |norms - target|; returns (d loss / d weights, target)."""
    ratio = sigmoid_np(loss_values) / initial
    inv_rate = ratio / np.mean(ratio)
    mean_norm = np.mean(norms.detach().numpy())
    target = torch.tensor(mean_norm * (inv_rate ** alpha), requires_grad=False)
    gn_loss = torch.sum(torch.abs(norms - target))
    return torch.autograd.grad(gn_loss, weights)[0], target.numpy()


def renormalize_(weights: torch.Tensor, total: float) -> None:
    """train_and_test.py:752-757: clamp at 0, rescale to a fixed sum."""
    weights.data[:].clamp_(min=0.0)
    weights.data = weights.data * (total / torch.sum(weights.data, dim=0))


def named_losses(mods, xt, yt, xs, ys, style_weight, adain=None, gram_style_loss=None):
    """The nine named losses of train_and_test.py:547-611 for the AdaIN/Gram variant of the transfer site (north_star):
    the two flow losses are replaced by the Gram style loss (target side) and the AdaIN content loss (source side);
    the adversarial / self-supervised terms are stand-ins with the same structure (difference of mean magnitudes, the
    wgan_loss form of widgets.py:57-61 over identity critics, mean squares).  ``mods`` = (FE_t, CL_t, FE_s, CL_s) with
    the reference's module interface: make_golden.py passes the reference's modules, tests/test_gpu_drivers.py the CUDA
    modules together with the CUDA ``adain`` / ``gram_style_loss``."""
    import torch.nn.functional as F
    from . import style as S
    adain = adain or S.adain
    gram_style_loss = gram_style_loss or S.gram_style_loss
    fe_t, cl_t, fe_s, cl_s = mods
    tf = fe_t(xt)
    sf = fe_s(xs)
    s2t = adain(sf, tf)
    logits_t, pooled_t = cl_t(tf)
    cl_t.eval()                                                   # train_and_test.py:584-586
    logits_s2t, pooled_s2t = cl_t(s2t)
    cl_t.train()
    logits_s, pooled_s = cl_s(sf)
    return dict(
        target_nf_loss=style_weight * gram_style_loss(s2t, tf),
        source_nf_loss=torch.mean((s2t - sf) ** 2),
        target_classification_loss=F.cross_entropy(logits_t, yt),
        source_classification_loss=F.cross_entropy(logits_s, ys),
        s2t2s_classification_loss=F.cross_entropy(cl_s.hidden(pooled_s2t), ys),
        cdan_loss=torch.mean(torch.abs(tf)) - torch.mean(torch.abs(s2t)),
        feature_discriminator_s_loss=-torch.mean(pooled_t) - torch.mean(pooled_s2t) + torch.mean(pooled_s),
        t_sl_loss=0.01 * torch.mean(tf ** 2),
        s_sl_loss=0.01 * torch.mean(sf ** 2))


LOSSES_T = ("target_nf_loss", "target_classification_loss")                                    # train_and_test.py:648-650
LOSSES_S = ("source_nf_loss", "source_classification_loss", "s2t2s_classification_loss")      # :651-654
REMAINDER = ("cdan_loss", "feature_discriminator_s_loss", "t_sl_loss", "s_sl_loss")            # :668-675
