"""The benchmarked training step of the hot path, host side.

Mirrors the data flow of the reference's joint stage (train_and_test.py:547-603) with the flow-based
transfer replaced by AdaIN + Gram loss as BASELINE.json's north_star prescribes:

    tf  = FE_t(xt)                 target extractor          (OS_CNN_res)
    sf  = FE_s(xs)                 source extractor          (OS_CNN_res)
    ssf = DimensionUnification(sf) source -> target shape    (torch, "next" row)
    s2t = AdaIN(ssf, tf)           feature-level style transfer
    L_style = Gram(s2t, tf)
    logits_t = CL_t(tf); logits_s = CL_s(ssf)                (OS_CNN)
    loss = CE_t + CE_s + style_weight * L_style ; backward ; RMSprop (reference learning rates)

Data parallel: one process per GPU, batch sharded by rank, per-rank BatchNorm statistics (DDP semantics),
one all-reduce over a flat fp32 gradient bucket per step followed by an explicit 1/N scale (RMSprop is not
linear in the gradient, so the scale is NOT folded into the learning rate).
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as TF
from .OS_CNN.OS_CNN import OS_CNN, OS_CNN_res, layer_parameter_list_input_change
from .OS_CNN.OS_CNN_Structure_build import generate_layer_parameter_list
from .widgets import DimensionUnification

MAX_KERNEL_SIZE = 89        # train_and_test.py:40
LEARNING_RATES = dict(fe_t=0.001, cl_t=0.003, fe_s=0.001, du=0.001, cl_s=0.003)    # train_and_test.py:97-101


def trainer_layer_lists(C: int, L: int):
    """Extractor / classifier layer lists exactly as train_and_test.py:38-53 derives them."""
    budgets = [8 * 128 * C, 5 * 128 * 256 + 2 * 256 * 128]
    ext = generate_layer_parameter_list(1, min(int(L / 4), MAX_KERNEL_SIZE), budgets, C)
    cf = sum(p[1] for p in ext[-1])
    return ext, layer_parameter_list_input_change(ext, cf), cf


class StyleTransferModelSet(nn.Module):
    """The five modules of the step, constructed in the reference's order (train_and_test.py:47-67) so that one
    ``torch.manual_seed`` reproduces the reference's initial parameters."""

    def __init__(self, Ct: int, Lt: int, Kt: int, Cs: int, Ls: int, Ks: int):
        super().__init__()
        lpl_t, lpl_c, cf_t = trainer_layer_lists(Ct, Lt)
        lpl_s, _, cf_s = trainer_layer_lists(Cs, Ls)
        self.fe_t = OS_CNN_res(lpl_t)
        self.cl_t = OS_CNN(lpl_c, Kt)
        self.fe_s = OS_CNN_res(lpl_s)
        self.du = DimensionUnification(cf_s, cf_t, Ls, Lt)
        self.cl_s = OS_CNN(lpl_c, Ks)                 # the reference reuses the target's list (train_and_test.py:67)
        self.feature_channels = cf_t

    def forward(self, xt, yt, xs, ys, style_weight: float = 1.0) -> Dict[str, torch.Tensor]:
        tf = self.fe_t(xt)
        sf = self.fe_s(xs)
        ssf = self.du(sf)
        s2t = TF.adain(ssf, tf)
        l_style = TF.gram_style_loss(s2t, tf)
        logits_t, _ = self.cl_t(tf)
        logits_s, _ = self.cl_s(ssf)
        ce_t = F.cross_entropy(logits_t, yt)
        ce_s = F.cross_entropy(logits_s, ys)
        return dict(loss=ce_t + ce_s + style_weight * l_style, ce_t=ce_t, ce_s=ce_s, l_style=l_style,
                    logits_t=logits_t, logits_s=logits_s, tf=tf, ssf=ssf, s2t=s2t)


class FlatGradBucket:
    """All parameter gradients as views into one flat fp32 buffer, so the data-parallel exchange is ONE
    all-reduce per step.  ``p.grad`` is pre-bound to its slice; autograd accumulates in place."""

    def __init__(self, params: List[torch.nn.Parameter]):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, device=self.params[0].device, dtype=torch.float32)
        off = 0
        for p in self.params:
            p.grad = self.flat[off: off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        self.flat.zero_()

    def all_reduce_mean(self, group=None):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.flat.mul_(1.0 / dist.get_world_size(group))


class Trainer:
    """forward + backward + (all-reduce) + RMSprop of the cfg2 step."""

    def __init__(self, model: StyleTransferModelSet, style_weight: float = 1.0, group=None):
        self.model = model
        self.style_weight = style_weight
        self.group = group
        groups = [dict(params=list(getattr(model, name).parameters()), lr=lr) for name, lr in LEARNING_RATES.items()]
        self.bucket = FlatGradBucket([p for g in groups for p in g["params"]])
        self.opt = torch.optim.RMSprop(groups, lr=0.001, foreach=True)

    def broadcast_parameters(self, src: int = 0):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            for t in list(self.model.parameters()) + list(self.model.buffers()):
                dist.broadcast(t.data, src=src, group=self.group)

    def step(self, xt, yt, xs, ys) -> torch.Tensor:
        self.bucket.zero()
        out = self.model(xt, yt, xs, ys, self.style_weight)
        out["loss"].backward()
        self.bucket.all_reduce_mean(self.group)
        self.opt.step()
        return out["loss"].detach()
