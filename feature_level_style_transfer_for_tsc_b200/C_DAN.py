"""Conditional-adversarial (C-DAN) consumer of the transferred features -- drop-in mirror of the reference's
``C_DAN.py`` (same names and call signatures: ``RandomLayer``, ``Entropy``, ``grl_hook``, ``calc_coeff``, ``CDAN``)
for BASELINE configuration 3.

``CDAN(...)`` with a two-view ``RandomLayer`` (the only form the trainer uses, train_and_test.py:74-76,590-591) runs

    y0 = [flatten(target feature); flatten(generated feature)] @ R0      cuBLAS (plain dense GEMM, [2B, C*L] x [C*L, 1024])
    fusion, u = tsc_cdan_fuse_fwd(y0, logits, R1)                         softmax, p @ R1, scale, product, entropy weights
    out = critic(fusion)                                                  three Linear layers (cuBLAS), ONE pass over 2B rows
    loss = tsc_cdan_distance_fwd(u, out)

and the matching two backward kernels, which also apply the reference's three gradient reversals (the hooks of
C_DAN.py:68-69 and widgets.py:121-122) with coefficients read from a small device buffer -- no host synchronisation
(the reference's ``.item()`` at C_DAN.py:72,75 becomes a detached device sum), so the whole loss is CUDA-graph
capturable.  Every other call shape (any number of views, ``random_layer=None``) takes the generic route below, which
is the reference's sequence of torch operators on CUDA tensors.  There is no CPU path.
"""
import math

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .widgets import AdversarialNetworkforCDAN


FUSED = True          # debugging switch: False sends every call through the generic torch-operator route


class RandomLayer(nn.Module):
    """Fixed Gaussian projections of several views, multiplied element-wise (reference lines 11-25).  The matrices are
    drawn on the host from the global generator exactly as the reference draws them (one seed, same values) and are
    plain tensors, not parameters or buffers."""

    def __init__(self, input_dim_list=[], output_dim=1024, with_nvidia=True):
        super(RandomLayer, self).__init__()
        self.input_num = len(input_dim_list)
        self.output_dim = output_dim
        self.random_matrix = [torch.randn(input_dim_list[i], output_dim) for i in range(self.input_num)]
        if with_nvidia:
            for i in range(self.input_num):
                self.random_matrix[i] = self.random_matrix[i].float().cuda()

    def _apply(self, fn, *args, **kwargs):
        # the reference moves the matrices in its constructor (with_nvidia); following .cuda()/.to() as well costs nothing
        self.random_matrix = [fn(m) for m in self.random_matrix]
        return super(RandomLayer, self)._apply(fn, *args, **kwargs)

    @property
    def scale_div(self):
        return math.pow(float(self.output_dim), 1.0 / self.input_num)

    def forward(self, input_list):
        return_list = [torch.mm(input_list[i], self.random_matrix[i]) for i in range(self.input_num)]
        return_tensor = return_list[0] / self.scale_div
        for single in return_list[1:]:
            return_tensor = torch.mul(return_tensor, single)
        return return_tensor


def Entropy(input_):
    """-sum p log(p + 1e-5) over dim 1 of already soft-maxed rows (reference lines 32-37)."""
    epsilon = 1e-5
    entropy = -input_ * torch.log(input_ + epsilon)
    return torch.sum(entropy, dim=1)


def grl_hook(coeff):
    def fun1(grad):
        return -coeff * grad.clone()
    return fun1


def calc_coeff(iter_num, high=1.0, low=0.0, alpha=100.0, max_iter=50.0):
    return float(2.0 * (high - low) / (1.0 + np.exp(-alpha * iter_num / max_iter)) - (high - low) + low)


class _RandomProjectPair(torch.autograd.Function):
    """y0 = [a; b] @ R0 without materialising the concatenation; R0 is a constant."""

    @staticmethod
    def forward(ctx, a, b, r0):
        B = a.shape[0]
        y0 = torch.empty((2 * B, r0.shape[1]), device=a.device, dtype=torch.float32)
        torch.mm(a, r0, out=y0[:B])
        torch.mm(b, r0, out=y0[B:])
        ctx.r0 = r0
        return y0

    @staticmethod
    def backward(ctx, dy0):
        B = dy0.shape[0] // 2
        rt = ctx.r0.t()
        return torch.mm(dy0[:B], rt), torch.mm(dy0[B:], rt), None


class _CdanFuse(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y0, logits, r1, coeff, scale_div):
        fusion, prob, u = ops.cdan_fuse_fwd(y0, logits.contiguous(), r1, scale_div)
        ctx.save_for_backward(y0, prob, r1, u, coeff)
        ctx.scale_div = scale_div
        return fusion, u

    @staticmethod
    def backward(ctx, dfusion, du):
        y0, prob, r1, u, coeff = ctx.saved_tensors
        dy0, dlogits = ops.cdan_fuse_bwd(dfusion.contiguous(), y0, prob, r1, u,
                                         None if du is None else du.contiguous(), coeff, ctx.scale_div)
        return dy0, dlogits, None, None, None


class _CdanDistance(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u, critic_out):
        loss, saved = ops.cdan_distance_fwd(u, critic_out.contiguous())
        ctx.save_for_backward(saved)
        ctx.B = u.numel() // 2
        return loss

    @staticmethod
    def backward(ctx, dloss):
        (saved,) = ctx.saved_tensors
        du, dcritic = ops.cdan_distance_bwd(dloss.contiguous().float(), saved, ctx.B)
        return du, dcritic


def _cdan_fused(input_target, input_g_from_source, prob_target, prob_g_from_source, ad_net, random_layer):
    B = input_target.shape[0]
    r0, r1 = random_layer.random_matrix
    y0 = _RandomProjectPair.apply(input_target, input_g_from_source, r0)
    logits = torch.cat([prob_target, prob_g_from_source], 0)
    coeff = ad_net.reversal_coefficients(calls=2)          # the two critic calls of the reference, then ad_net.coeff
    fusion, u = _CdanFuse.apply(y0, logits, r1, coeff, random_layer.scale_div)
    out = ad_net.critic(fusion)                            # reversal of its input gradient happens in _CdanFuse.backward
    return _CdanDistance.apply(u, out)


def CDAN(input_target, input_g_from_source, prob_target, prob_g_from_source, ad_net: AdversarialNetworkforCDAN,
         random_layer=None):
    """distance_target - distance_generated (reference lines 49-82).  ``prob_*`` are the classifiers' LOGITS (the
    reference soft-maxes them itself)."""
    if not input_target.is_cuda:
        raise RuntimeError("CDAN needs CUDA tensors: the tsc_b200 path has no CPU fallback")
    input_target = torch.flatten(input_target, 1)
    input_g_from_source = torch.flatten(input_g_from_source, 1)
    if (FUSED and random_layer is not None and random_layer.input_num == 2 and prob_target.shape[1] <= 32
            and input_target.shape == input_g_from_source.shape and random_layer.output_dim <= 8192):
        return _cdan_fused(input_target, input_g_from_source, prob_target, prob_g_from_source, ad_net, random_layer)
    prob_target = torch.nn.functional.softmax(prob_target, dim=1)
    prob_g_from_source = torch.nn.functional.softmax(prob_g_from_source, dim=1)
    if random_layer is None:
        fusion_target = torch.bmm(prob_target.unsqueeze(2), input_target.unsqueeze(1))
        target_out = ad_net(fusion_target.view(-1, input_target.size(1) * prob_target.size(1)))
        fusion_source = torch.bmm(prob_g_from_source.unsqueeze(2), input_g_from_source.unsqueeze(1))
        g_source_out = ad_net(fusion_source.view(-1, input_g_from_source.size(1) * prob_g_from_source.size(1)))
    else:
        fusion_target = random_layer.forward([input_target, prob_target])
        target_out = ad_net(fusion_target.view(-1, fusion_target.size(1)))
        fusion_source = random_layer.forward([input_g_from_source, prob_g_from_source])
        g_source_out = ad_net(fusion_source.view(-1, fusion_source.size(1)))
    entropy_target = Entropy(prob_target)
    entropy_g_from_source = Entropy(prob_g_from_source)
    coeff = ad_net.coeff
    entropy_target.register_hook(grl_hook(coeff))
    entropy_g_from_source.register_hook(grl_hook(coeff))
    weight_target = 1.0 + torch.exp(-entropy_target)
    weight_g_from_source = 1.0 + torch.exp(-entropy_g_from_source)
    weight_target = weight_target / torch.sum(weight_target).detach()
    weight_g_from_source = weight_g_from_source / torch.sum(weight_g_from_source).detach()
    distance_target = torch.sum(weight_target * target_out)
    distance_g_from_source = torch.sum(weight_g_from_source * g_source_out)
    return distance_target - distance_g_from_source
