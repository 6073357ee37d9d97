"""Oracle of the benchmarked training step, on CPU.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Also the timed CPU baseline of bench.py
(``cpu_baseline`` leg and ``--impl reference``): the same step the CUDA path runs, executed
with torch CPU kernels (oneDNN convolutions), which is what the reference's modules execute
on a CPU host.

Step (cfg2): tf = FE_t(xt); sf = FE_s(xs); ssf = DimensionUnification(sf);
s2t = AdaIN(ssf, tf); L_style = Gram(s2t, tf); logits_t = CL_t(tf); logits_s = CL_s(ssf);
loss = CE_t + CE_s + lambda * L_style; backward; RMSprop with the reference learning rates
(train_and_test.py:97-101).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict

import torch
import torch.nn.functional as F

from . import os_cnn as O
from . import style as S


def init_dimension_unification(src_c, tgt_c, src_l, tgt_l):
    """widgets.py:66-72: nn.Linear(src_l, tgt_l) then nn.Conv1d(src_c, tgt_c, 1)."""
    lin = torch.nn.Linear(src_l, tgt_l)
    conv = torch.nn.Conv1d(src_c, tgt_c, 1)
    sd = OrderedDict()
    sd["length_unification.weight"] = lin.weight.detach().clone()
    sd["length_unification.bias"] = lin.bias.detach().clone()
    sd["channel_unification.weight"] = conv.weight.detach().clone()
    sd["channel_unification.bias"] = conv.bias.detach().clone()
    return sd


def dimension_unification(sd, x):
    """widgets.py:73-78: relu(conv1x1(relu(Linear over L)))."""
    h = F.relu(F.linear(x, sd["length_unification.weight"], sd["length_unification.bias"]))
    return F.relu(F.conv1d(h, sd["channel_unification.weight"], sd["channel_unification.bias"]))


class ModelSet:
    """The five modules of the step, built in the reference's construction order
    (train_and_test.py:47-67) under one seed."""

    LRS = dict(fe_t=0.001, cl_t=0.003, fe_s=0.001, du=0.001, cl_s=0.003)   # train_and_test.py:97-101

    def __init__(self, Ct, Lt, Kt, Cs, Ls, Ks, seed: int = 0):
        torch.manual_seed(seed)
        self.lpl_t, self.lpl_c = O.trainer_layer_lists(Ct, Lt)
        self.lpl_s, _ = O.trainer_layer_lists(Cs, Ls)
        self.fe_t = O.init_extractor(self.lpl_t)
        self.cl_t = O.init_classifier(self.lpl_c, Kt)
        self.fe_s = O.init_extractor(self.lpl_s)
        self.du = init_dimension_unification(O.feature_channels(self.lpl_s), O.feature_channels(self.lpl_t), Ls, Lt)
        self.cl_s = O.init_classifier(self.lpl_c, Ks)    # train_and_test.py:67 reuses the target list
        self.sq = {}                                      # RMSprop square averages

    def groups(self) -> Dict[str, Dict[str, torch.Tensor]]:
        return dict(fe_t=self.fe_t, cl_t=self.cl_t, fe_s=self.fe_s, du=self.du, cl_s=self.cl_s)

    def trainable(self):
        for gname, sd in self.groups().items():
            for k, v in sd.items():
                if v.is_floating_point() and "running_" not in k:
                    yield gname, k, v

    def set_requires_grad(self):
        for _, _, v in self.trainable():
            v.requires_grad_(True)
            v.grad = None


def step_forward(ms: ModelSet, xt, yt, xs, ys, style_weight: float = 1.0, training: bool = True):
    tf = O.extractor_forward(ms.fe_t, ms.lpl_t, xt, training)
    sf = O.extractor_forward(ms.fe_s, ms.lpl_s, xs, training)
    ssf = dimension_unification(ms.du, sf)
    s2t = S.adain(ssf, tf)
    l_style = S.gram_style_loss(s2t, tf)
    logits_t, _ = O.classifier_forward(ms.cl_t, ms.lpl_c, tf, training)
    logits_s, _ = O.classifier_forward(ms.cl_s, ms.lpl_c, ssf, training)
    ce_t = F.cross_entropy(logits_t, yt)
    ce_s = F.cross_entropy(logits_s, ys)
    loss = ce_t + ce_s + style_weight * l_style
    return dict(loss=loss, ce_t=ce_t, ce_s=ce_s, l_style=l_style, logits_t=logits_t, logits_s=logits_s,
                tf=tf, ssf=ssf, s2t=s2t)


def rmsprop_update(ms: ModelSet, alpha: float = 0.99, eps: float = 1e-8):
    """torch.optim.RMSprop defaults (train_and_test.py:97-101): v = a v + (1-a) g^2; p -= lr g/(sqrt(v)+eps)."""
    with torch.no_grad():
        for gname, k, p in ms.trainable():
            if p.grad is None:
                continue
            key = (gname, k)
            v = ms.sq.setdefault(key, torch.zeros_like(p))
            v.mul_(alpha).addcmul_(p.grad, p.grad, value=1 - alpha)
            p.addcdiv_(p.grad, v.sqrt().add_(eps), value=-ModelSet.LRS[gname])
            p.grad = None


def train_step(ms: ModelSet, xt, yt, xs, ys, style_weight: float = 1.0):
    ms.set_requires_grad()
    out = step_forward(ms, xt, yt, xs, ys, style_weight, training=True)
    out["loss"].backward()
    rmsprop_update(ms)
    return float(out["loss"].detach())
