"""Conditional-adversarial (C-DAN) consumer of the transferred features -- the reference's ``C_DAN.py`` surface that the
trainer uses (``RandomLayer([Cf*L, K])`` and ``CDAN(...)``, train_and_test.py:74-76,590-594) on this repository's kernels,
for BASELINE configuration 3.

``CDAN(...)`` runs

    y0 = [flatten(target feature); flatten(generated feature)] @ R0      random projection, [2B, C*L] x [C*L, 1024]
    fusion, u = tsc_cdan_fuse_fwd(y0, logits, R1)                         softmax, p @ R1, scale, product, entropy weights
    out = critic(fusion)                                                  three Linear layers, ONE pass over 2B rows
    loss = tsc_cdan_distance_fwd(u, out)

and the matching two backward kernels, which also apply the reference's three gradient reversals (the hooks of
C_DAN.py:68-71 and widgets.py:121-122) with coefficients read from a small device buffer -- no host synchronisation
(the reference's ``.item()`` at C_DAN.py:75,77 becomes a detached device sum), so the whole loss is CUDA-graph
capturable.  The only call shape is the trainer's: a two-view ``RandomLayer`` (feature, class logits).  Anything else --
``random_layer=None`` (the outer-product form, C_DAN.py:55-59), more views, CPU tensors -- raises: there is no
alternate route.
"""
import math

import torch
import torch.nn as nn

from . import ops
from .widgets import AdversarialNetworkforCDAN


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
        raise RuntimeError("RandomLayer is evaluated inside CDAN(...): the projection, the class-probability product and the "
                           "entropy weights are one fused kernel chain (no stand-alone torch route)")


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
    if random_layer is None or random_layer.input_num != 2:
        raise RuntimeError("CDAN: only the trainer's call shape is built -- a RandomLayer over (feature, class logits), "
                           "train_and_test.py:76,594")
    if prob_target.shape[1] > 32 or random_layer.output_dim > 8192 or input_target.shape != input_g_from_source.shape:
        raise RuntimeError(f"CDAN: unsupported shape (classes {prob_target.shape[1]} > 32, projection width "
                           f"{random_layer.output_dim} > 8192, or feature shapes {tuple(input_target.shape)} != "
                           f"{tuple(input_g_from_source.shape)})")
    return _cdan_fused(input_target, input_g_from_source, prob_target, prob_g_from_source, ad_net, random_layer)
