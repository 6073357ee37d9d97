"""GradNorm joint-stage driver: host mirror of train_and_test.py:499-511 and :646-766 on the CUDA modules.

What the reference does per batch once its named losses exist: balance two groups of losses with learnable weights
(target side ``[2, 5]``, source side ``[2, 2, 4]``), back-propagate the balanced total, measure for every balanced loss
the gradient norm over the parameters of the extractor's last block (``return_last_layer()``), derive the GradNorm
gradient of the weights, "clear the graph" with a second backward, step every optimizer, renormalise the weights and
clip the critics.  Here:

* the per-loss norms are ``torch.autograd.grad`` calls through the ``os_stack`` node under ``functional.dense_wgrad()``
  -- the reference's ``.grad`` of a kernel bank is the *unmasked* gradient and its masked taps are part of
  every ``torch.norm`` -- followed by ONE fused reduction per loss (``tsc_multi_l2norm``: 2 launches instead of
  12 x ``torch.norm`` + ``cat`` + ``sum``);
* the reference's two backward passes (``loss_total.backward(retain_graph=True)``, then -- after zeroing ``.data`` of the
  balanced weights -- ``loss_total.backward()``) accumulate ``grad(total) + grad(remainder)`` into ``.grad``
  (``oracle/grad_norm.py`` explains why); the driver differentiates ``total + remainder`` once, which is the same sum;
* everything the reference pulls to the host (``.cpu().numpy()`` of the losses and norms, 10 floats) is ONE device-to-host
  copy; the weight gradient is computed there in numpy exactly as the reference does and copied back.

PyTorch here is plumbing (tiny weight tensors, Adam over 2-3 scalars); the optimizers of the modules are whatever the
caller passes (``torch.optim`` instances as in train_and_test.py:97-106, or a fused ``train_step.FlatParameters`` step).
"""
from __future__ import annotations

from typing import Callable, Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import functional as TF
from . import ops

ALPHA = 3                                            # train_and_test.py:511


def remainder_coefficients(cur_epoch: int) -> Tuple[float, float, float, float]:
    """(cdan, feature discriminator, t_sl, s_sl) multipliers of the un-balanced losses (train_and_test.py:668-675)."""
    if cur_epoch < 12:
        return 3, 3, 2, 2
    if cur_epoch < 24:
        return 2, 3, 1.8, 1.5
    if cur_epoch < 50:
        return 1.5, 2, 1.8, 1.8
    return 1.5, 1.5, 2.5, 2.5


class GradNormSide:
    """The balanced weights of one side, their Adam optimizer and the first-batch reference losses
    (train_and_test.py:501-510)."""

    def __init__(self, names: Sequence[str], init: Sequence[float], total: float, lr: float, device, capturable: bool = False):
        self.names = tuple(names)
        self.total = float(total)
        self.weights = nn.Parameter(torch.tensor(list(init), dtype=torch.float32, device=device))
        # capturable: Adam keeps its step count on the device, so that its update can be part of a CUDA graph (same formulas)
        self.optimizer = torch.optim.Adam([self.weights], lr=lr, capturable=bool(capturable))
        self.initial: Optional[np.ndarray] = None     # sigmoid of the first batch's losses (:663-666)
        self.initial_dev: Optional[torch.Tensor] = None   # the same on the device (step_device)

    def renormalize_(self) -> None:
        """train_and_test.py:752-757"""
        with torch.no_grad():
            self.weights.clamp_(min=0.0)
            self.weights.mul_(self.total / torch.sum(self.weights))


def target_side(device, capturable: bool = False) -> GradNormSide:
    return GradNormSide(("target_nf_loss", "target_classification_loss"), (2, 5), 7, 0.0002, device, capturable)


def source_side(device, capturable: bool = False) -> GradNormSide:
    return GradNormSide(("source_nf_loss", "source_classification_loss", "s2t2s_classification_loss"), (2, 2, 4), 8, 0.001, device,
                        capturable)


def shared_grad_norm_sums(losses: Sequence[torch.Tensor], shared_params: List[torch.Tensor]) -> torch.Tensor:
    """[sum_p ||d loss_i / d p||_2 for i] (device fp32) over the parameter tensors ``shared_params``, with the
    reference's unmasked weight gradients.  The graph is retained."""
    sums = []
    with TF.dense_wgrad():
        for li in losses:
            g = torch.autograd.grad(li, shared_params, retain_graph=True, allow_unused=True)
            g = [t for t in g if t is not None]
            sums.append(ops.multi_l2norm(g)[len(g)])
    return torch.stack(sums)


def _weight_gradient(w: np.ndarray, sums: np.ndarray, loss_values: np.ndarray, initial: np.ndarray, alpha: float):
    """train_and_test.py:691-715 on the host, float32 like the reference's numpy arrays."""
    norms = (np.abs(w) * sums).astype(np.float32)              # sum_p ||w_i g_p|| = |w_i| sum_p ||g_p||
    ratio = (1 / (1 + np.exp(-loss_values))) / initial
    inv_rate = ratio / np.mean(ratio)
    target = (np.mean(norms) * (inv_rate ** alpha)).astype(np.float32)
    grad = (np.sign(norms - target) * np.sign(w) * sums).astype(np.float32)
    return norms, target, grad


def _weight_gradient_device(w: torch.Tensor, sums: torch.Tensor, loss_values: torch.Tensor, initial: torch.Tensor, alpha: float):
    """``_weight_gradient`` on the device (fp32 tensors of 2-3 elements): no host round trip, capturable in a CUDA graph."""
    norms = w.abs() * sums
    ratio = torch.sigmoid(loss_values) / initial
    inv_rate = ratio / ratio.mean()
    target = norms.mean() * inv_rate.pow(alpha)
    grad = torch.sign(norms - target) * torch.sign(w) * sums
    return norms, target, grad


class JointStageDriver:
    """One call of ``step`` = train_and_test.py:646-766 for one batch.

    shared_t / shared_s : the modules whose parameters GradNorm differentiates against
                          (``extractor.return_last_layer()``, :681-682)
    optimizers          : the module optimizers (``optimizer_list`` + ``optimizer_sl_cpc``, :677-679,746-750); each needs
                          ``zero_grad()`` and ``step()``
    clamps              : [(module, c)] WGAN weight clipping after the step (:763-766)
    """

    def __init__(self, shared_t: nn.Module, shared_s: nn.Module, optimizers: Sequence, clamps: Iterable = (),
                 alpha: float = ALPHA, device="cuda", capturable: bool = False):
        self.shared_t, self.shared_s = shared_t, shared_s
        self.optimizers = list(optimizers)
        self.clamps = list(clamps)
        self.alpha = alpha
        self.t = target_side(device, capturable)
        self.s = source_side(device, capturable)

    def step_device(self, losses: Dict[str, torch.Tensor], coefficients: torch.Tensor) -> Dict[str, torch.Tensor]:
        """``step`` without any host round trip: the remainder multipliers arrive as a device tensor of four floats
        (``remainder_coefficients(cur_epoch)``), the GradNorm weight gradient is computed on the device, and the results stay
        device tensors.  Every operation is capturable, so a whole joint-stage step -- forward included -- can be ONE CUDA
        graph (``GraphedJointStage``); the optimizers passed to the driver must then be capturable too.  The first call
        (outside a capture) fixes the reference losses of train_and_test.py:663-666."""
        t, s = self.t, self.s
        lt = torch.stack([losses[k] for k in t.names])
        ls = torch.stack([losses[k] for k in s.names])
        c = coefficients
        remainder = (c[0] * losses["cdan_loss"] + c[1] * losses["feature_discriminator_s_loss"]
                     + c[2] * losses["t_sl_loss"] + c[3] * losses["s_sl_loss"])
        balanced = torch.sum(t.weights.detach() * lt) + torch.sum(s.weights.detach() * ls)
        for opt in self.optimizers:
            opt.zero_grad()
        sums_t = shared_grad_norm_sums(list(lt.unbind(0)), list(self.shared_t.parameters()))
        sums_s = shared_grad_norm_sums(list(ls.unbind(0)), list(self.shared_s.parameters()))
        if t.initial_dev is None:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("the first joint-stage step fixes the reference losses and must run outside a CUDA graph")
            t.initial_dev = torch.sigmoid(lt.detach()).clone()
            s.initial_dev = torch.sigmoid(ls.detach()).clone()
            t.initial, s.initial = t.initial_dev.cpu().numpy(), s.initial_dev.cpu().numpy()
        norms_t, target_t, grad_t = _weight_gradient_device(t.weights.detach(), sums_t, lt.detach(), t.initial_dev, self.alpha)
        norms_s, target_s, grad_s = _weight_gradient_device(s.weights.detach(), sums_s, ls.detach(), s.initial_dev, self.alpha)
        (balanced + 2.0 * remainder).backward()
        t.weights.grad = grad_t
        s.weights.grad = grad_s
        t.optimizer.step()
        s.optimizer.step()
        for opt in self.optimizers:
            opt.step()
        t.renormalize_()
        s.renormalize_()
        with torch.no_grad():
            for mod, cl in self.clamps:
                for p in mod.parameters():
                    p.clamp_(-cl, cl)
        return dict(norms_t=norms_t, norms_s=norms_s, target_t=target_t, target_s=target_s, grad_w_t=grad_t, grad_w_s=grad_s,
                    loss_t=lt.detach(), loss_s=ls.detach())

    def step(self, losses: Dict[str, torch.Tensor], cur_epoch: int) -> Dict[str, np.ndarray]:
        t, s = self.t, self.s
        lt = torch.stack([losses[k] for k in t.names])
        ls = torch.stack([losses[k] for k in s.names])
        c_cdan, c_fd, c_tsl, c_ssl = remainder_coefficients(cur_epoch)
        remainder = (c_cdan * losses["cdan_loss"] + c_fd * losses["feature_discriminator_s_loss"]
                     + c_tsl * losses["t_sl_loss"] + c_ssl * losses["s_sl_loss"])
        balanced = torch.sum(t.weights.detach() * lt) + torch.sum(s.weights.detach() * ls)
        for opt in self.optimizers:
            opt.zero_grad()
        # per-loss gradient norms over the shared blocks (retain the graph), then everything the host needs in one copy
        sums_t = shared_grad_norm_sums(list(lt.unbind(0)), list(self.shared_t.parameters()))
        sums_s = shared_grad_norm_sums(list(ls.unbind(0)), list(self.shared_s.parameters()))
        host = torch.cat([lt.detach(), ls.detach(), sums_t, sums_s, t.weights.detach(), s.weights.detach()]).cpu().numpy()
        nt, ns_ = len(t.names), len(s.names)
        lv_t, lv_s = host[:nt], host[nt:nt + ns_]
        sm_t, sm_s = host[nt + ns_:2 * nt + ns_], host[2 * nt + ns_:2 * (nt + ns_)]
        w_t, w_s = host[2 * (nt + ns_):3 * nt + 2 * ns_], host[3 * nt + 2 * ns_:]
        if t.initial is None:                                      # :663-666
            t.initial = 1 / (1 + np.exp(-lv_t))
            s.initial = 1 / (1 + np.exp(-lv_s))
        norms_t, target_t, grad_t = _weight_gradient(w_t, sm_t, lv_t, t.initial, self.alpha)
        norms_s, target_s, grad_s = _weight_gradient(w_s, sm_s, lv_s, s.initial, self.alpha)
        # the reference's two backward passes add up to grad(balanced + remainder) + grad(remainder)
        (balanced + 2.0 * remainder).backward()
        dev = t.weights.device
        t.weights.grad = torch.from_numpy(grad_t).to(dev)          # :744-745
        s.weights.grad = torch.from_numpy(grad_s).to(dev)
        t.optimizer.step()
        s.optimizer.step()
        for opt in self.optimizers:
            opt.step()
        t.renormalize_()
        s.renormalize_()
        with torch.no_grad():
            for mod, c in self.clamps:
                for p in mod.parameters():
                    p.clamp_(-c, c)
        return dict(norms_t=norms_t, norms_s=norms_s, target_t=target_t, target_s=target_s, grad_w_t=grad_t, grad_w_s=grad_s,
                    loss_t=lv_t, loss_s=lv_s)


class GraphedJointStage:
    """A joint-stage step as ONE CUDA graph: the named losses (forward of every module), the per-loss gradient-norm passes over
    the shared blocks, the GradNorm weight gradient, the main backward, Adam on the balanced weights, the module optimizers,
    the renormalisation and the WGAN clamps (train_and_test.py:547-611 + 646-766 for one batch).  Eagerly that step is
    ~700 launches and 15 ms at cfg2 size (``tools/bench_eval.py``), almost all of it host launch latency.

    loss_fn(*inputs) -> dict of the named losses; it is called on STATIC copies of the inputs.  The driver must have been built
    with ``capturable=True`` and capturable optimizers (``torch.optim.RMSprop(..., capturable=True)`` or a fused
    ``train_step.FlatParameters`` step).  The first two calls run eagerly (the first fixes the GradNorm reference losses,
    both warm the allocator and the libraries up), the third captures, later calls replay; the remainder multipliers of the
    current epoch travel in a device tensor, so one graph serves every epoch."""

    EAGER_STEPS = 2

    def __init__(self, driver: JointStageDriver, loss_fn: Callable[..., Dict[str, torch.Tensor]]):
        self.driver, self.loss_fn = driver, loss_fn
        self.calls = 0
        self._graph = None
        self._static_in: Optional[List[torch.Tensor]] = None
        self._coef = None
        self._out: Optional[Dict[str, torch.Tensor]] = None

    def _body(self) -> Dict[str, torch.Tensor]:
        return self.driver.step_device(self.loss_fn(*self._static_in), self._coef)

    def step(self, *inputs: torch.Tensor, cur_epoch: int = 0) -> Dict[str, torch.Tensor]:
        if self._static_in is None:
            self._static_in = [t.clone() for t in inputs]
            self._coef = torch.zeros(4, device=inputs[0].device, dtype=torch.float32)
        for dst, src in zip(self._static_in, inputs):
            dst.copy_(src, non_blocking=True)
        self._coef.copy_(torch.tensor(remainder_coefficients(cur_epoch), dtype=torch.float32), non_blocking=True)
        self.calls += 1
        if self.calls <= self.EAGER_STEPS:
            # on a side stream, as the warm-up of a graph capture has to be
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                out = self._body()
            torch.cuda.current_stream().wait_stream(side)
            return out
        if self._graph is None:
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                self._out = self._body()
            # capturing does not execute: the step itself is the first replay
        self._graph.replay()
        return self._out
