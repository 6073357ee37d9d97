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

from . import cdan as CD
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

        self._init_extra(Ct, Lt, Kt)

    def _init_extra(self, Ct, Lt, Kt):
        pass

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
            p.addcdiv_(p.grad, v.sqrt().add_(eps), value=-ms.LRS[gname])
            clamp = getattr(ms, "CLAMPS", {}).get(gname, 0.0)
            if clamp > 0:
                p.clamp_(-clamp, clamp)                   # WGAN clipping of the critic, train_and_test.py:763-764
            p.grad = None


def train_step(ms: ModelSet, xt, yt, xs, ys, style_weight: float = 1.0):
    ms.set_requires_grad()
    out = step_forward(ms, xt, yt, xs, ys, style_weight, training=True)
    out["loss"].backward()
    rmsprop_update(ms)
    return float(out["loss"].detach())


# ---- configuration 3: one (source, target) pair with the C-DAN consumer, and several pairs on one target batch ---------

class PairModelSet(ModelSet):
    """The cfg2 modules plus RandomLayer([Cf*Lt, Kt]) and the critic (train_and_test.py:74-76), drawn after them."""

    LRS = dict(ModelSet.LRS, ad_net=0.001)                 # train_and_test.py:105
    CLAMPS = dict(ad_net=0.0005)                           # train_and_test.py:763-764
    CDAN_WEIGHT = 3.0                                      # train_and_test.py:660 (cur_epoch < 12)

    def __init__(self, Ct, Lt, Kt, Cs, Ls, Ks, seed: int = 0, critic_hidden: int = 1024):
        self.critic_hidden = critic_hidden
        super().__init__(Ct, Lt, Kt, Cs, Ls, Ks, seed)

    def _init_extra(self, Ct, Lt, Kt):
        self.mats = CD.init_random_layer([O.feature_channels(self.lpl_t) * Lt, Kt])
        self.ad_net = CD.init_ad_net(1024, self.critic_hidden)
        self.ad_state = CD.AdNetState()

    def groups(self):
        return dict(super().groups(), ad_net=self.ad_net)


def pair_step_forward(ms: PairModelSet, xt, yt, xs, ys, style_weight: float = 1.0, dropout_p: float = 0.0):
    """train_and_test.py:547-603 restricted to the hot path: the cfg2 data flow, then the target classifier on the
    generated features in eval-BatchNorm mode (:584-586; AFTER the training-mode call has updated the running
    statistics) and CDAN (:590-591)."""
    out = step_forward(ms, xt, yt, xs, ys, style_weight, training=True)
    logits_s2t, _ = O.classifier_forward(ms.cl_t, ms.lpl_c, out["s2t"], training=False)
    cdan = CD.cdan(out["tf"], out["s2t"], out["logits_t"], logits_s2t, ms.ad_net, ms.ad_state, ms.mats,
                   training=True, dropout_p=dropout_p)
    out.update(loss=out["loss"] + ms.CDAN_WEIGHT * cdan, cdan=cdan, logits_s2t=logits_s2t)
    return out


def pair_train_step(ms: PairModelSet, xt, yt, xs, ys, style_weight: float = 1.0, dropout_p: float = 0.0):
    ms.set_requires_grad()
    out = pair_step_forward(ms, xt, yt, xs, ys, style_weight, dropout_p)
    out["loss"].backward()
    rmsprop_update(ms)
    return float(out["loss"].detach())


def multi_source_models(target, sources, seed: int = 0, critic_hidden: int = 1024):
    """One PairModelSet per source, drawn one after the other from ONE seed (as MultiSourceModelSet constructs them)."""
    Ct, Lt, Kt = target
    torch.manual_seed(seed)
    sets = []
    for (Cs, Ls, Ks) in sources:
        ms = PairModelSet.__new__(PairModelSet)
        ms.critic_hidden = critic_hidden
        _construct_without_seeding(ms, Ct, Lt, Kt, Cs, Ls, Ks)
        sets.append(ms)
    return sets


def _construct_without_seeding(ms, Ct, Lt, Kt, Cs, Ls, Ks):
    ms.lpl_t, ms.lpl_c = O.trainer_layer_lists(Ct, Lt)
    ms.lpl_s, _ = O.trainer_layer_lists(Cs, Ls)
    ms.fe_t = O.init_extractor(ms.lpl_t)
    ms.cl_t = O.init_classifier(ms.lpl_c, Kt)
    ms.fe_s = O.init_extractor(ms.lpl_s)
    ms.du = init_dimension_unification(O.feature_channels(ms.lpl_s), O.feature_channels(ms.lpl_t), Ls, Lt)
    ms.cl_s = O.init_classifier(ms.lpl_c, Ks)
    ms.sq = {}
    ms._init_extra(Ct, Lt, Kt)


def multi_source_train_step(sets, xt, yt, source_batches, style_weight: float = 1.0, dropout_p: float = 0.0):
    """Sum of the pair losses on one target batch; every pair updates its own modules."""
    return sum(pair_train_step(ms, xt, yt, xs, ys, style_weight, dropout_p) for ms, (xs, ys) in zip(sets, source_batches))
