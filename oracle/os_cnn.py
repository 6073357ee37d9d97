"""Oracle restatement of the reference's OS-CNN path (plain torch on CPU, fp32 or fp64).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Every function cites the reference
file:line it restates (paths relative to /root/reference).  The restatement is functional:
a model is an ``OrderedDict`` of tensors keyed exactly like the reference's ``state_dict``
plus the integer layer-parameter list, so the same dict can be loaded into the reference's
modules, into the product's modules, or evaluated here.

Pinned against the reference itself by ``oracle/make_golden.py`` -> ``tests/golden``.
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

LayerParams = List[Tuple[int, int, int]]  # [(in_ch, out_ch, kernel), ...] one OS layer

BN_EPS = 1e-5        # torch.nn.BatchNorm1d default, OS_CNN/OS_CNN.py:65
BN_MOMENTUM = 0.1    # torch.nn.BatchNorm1d default


# --------------------------------------------------------------------------------------
# kernel-bank sizing  (OS_CNN/OS_CNN_Structure_build.py)
# --------------------------------------------------------------------------------------
def primes_in_range(start: int, end: int) -> List[int]:
    """OS_CNN_Structure_build.py:3-13.  Trial division with an empty divisor loop for 1,
    so 1 is reported as "prime" (and so is 2)."""
    out = []
    for v in range(start, end + 1):
        if all(v % d for d in range(2, v)):
            out.append(v)
    return out


def out_channels_for_budget(budget: int, in_channel: int, primes: Sequence[int]) -> int:
    """OS_CNN_Structure_build.py:16-18: int(budget / (in_channel * sum(primes)))."""
    return int(budget / (in_channel * sum(primes)))


def generate_layer_parameter_list(start: int, end: int, budgets: Sequence[int],
                                  in_channel: int = 1) -> List[LayerParams]:
    """OS_CNN_Structure_build.py:20-42.  One layer per budget (every prime gets the same
    out-channel count), then a closing layer of two kernels (start, start+1) whose
    out-channel count is len(primes) * out_channels(first budget, original in_channel)."""
    primes = primes_in_range(start, end)
    first_in = in_channel
    layers: List[LayerParams] = []
    for budget in budgets:
        oc = out_channels_for_budget(budget, in_channel, primes)
        layers.append([(in_channel, oc, p) for p in primes])
        in_channel = len(primes) * oc
    last_oc = len(primes) * out_channels_for_budget(budgets[0], first_in, primes)
    layers.append([(in_channel, last_oc, start), (in_channel, last_oc, start + 1)])
    return layers


def layer_parameter_list_input_change(lpl: List[LayerParams], input_channel: int) -> List[LayerParams]:
    """OS_CNN/OS_CNN.py:142-152: rewrite layer 0's in_ch (classifier consumes features)."""
    out = []
    for i, layer in enumerate(lpl):
        out.append([(input_channel, oc, k) for (_, oc, k) in layer] if i == 0 else layer)
    return out


def feature_channels(lpl: List[LayerParams]) -> int:
    """Sum of the closing layer's out-channels (OS_CNN/OS_CNN.py:95-97,190-192)."""
    return sum(oc for (_, oc, _) in lpl[-1])


def trainer_layer_lists(C: int, L: int, max_kernel: int = 89):
    """train_and_test.py:38-53: extractor list and the classifier list derived from it."""
    budgets = [8 * 128 * C, 5 * 128 * 256 + 2 * 256 * 128]
    rf = min(int(L / 4), max_kernel)
    ext = generate_layer_parameter_list(1, rf, budgets, C)
    cls = layer_parameter_list_input_change(ext, feature_channels(ext))
    return ext, cls


# --------------------------------------------------------------------------------------
# mask geometry  (OS_CNN/OS_CNN.py:9-43, 59)
# --------------------------------------------------------------------------------------
def mask_interval(k: int, kmax: int) -> Tuple[int, int]:
    """OS_CNN/OS_CNN.py:9-12: live taps [left, left+k) of a size-k kernel in a kmax window."""
    right_zero = math.ceil((kmax - 1) / 2) - math.ceil((k - 1) / 2)
    left = kmax - k - right_zero
    return left, left + k


def bank_geometry(layer: LayerParams) -> Dict:
    """Geometry of one OS layer: channel counts, padding (OS_CNN.py:59), per-out-channel
    live interval, and s(t) = first out channel live at tap t."""
    kmax = layer[-1][-1]                      # OS_CNN.py:24
    cin = layer[0][0]
    lo, hi = [], []
    for (_, oc, k) in layer:
        l, r = mask_interval(k, kmax)
        lo += [l] * oc
        hi += [r] * oc
    cout = len(lo)
    s_of_tap = []
    for t in range(kmax):
        live = [c for c in range(cout) if lo[c] <= t < hi[c]]
        first = live[0] if live else cout
        # nestedness: everything from the first live channel upwards is live
        assert live == list(range(first, cout)), "kernel bank is not nested"
        s_of_tap.append(first)
    return dict(cin=cin, cout=cout, kmax=kmax, pad_l=int((kmax - 1) / 2), pad_r=int(kmax / 2),
                lo=lo, hi=hi, s_of_tap=s_of_tap)


def build_mask(layer: LayerParams) -> np.ndarray:
    """OS_CNN/OS_CNN.py:15-20,37-41: float32 mask [Cout, Cin, Kmax]."""
    g = bank_geometry(layer)
    m = np.zeros((g["cout"], g["cin"], g["kmax"]), np.float32)
    for c in range(g["cout"]):
        m[c, :, g["lo"][c]:g["hi"][c]] = 1.0
    return m


def live_macs_per_position(layer: LayerParams) -> int:
    """Cin * sum_g(out_g * k_g): effective MACs per output position."""
    return layer[0][0] * sum(oc * k for (_, oc, k) in layer)


# --------------------------------------------------------------------------------------
# parameter init replay  (OS_CNN/OS_CNN.py:23-43, 61-65)
# --------------------------------------------------------------------------------------
def _replay_layer_init(layer: LayerParams):
    """Consumes the torch global RNG exactly like build_layer_with_layer_parameter.__init__:
    one nn.Conv1d per prime (OS_CNN.py:29), then one discarded big Conv1d (OS_CNN.py:61)."""
    g = bank_geometry(layer)
    w = torch.zeros(g["cout"], g["cin"], g["kmax"])
    b = torch.zeros(g["cout"])
    c0 = 0
    for (ic, oc, k) in layer:
        conv = torch.nn.Conv1d(ic, oc, k)
        l, r = mask_interval(k, g["kmax"])
        w[c0:c0 + oc, :, l:r] = conv.weight.detach()
        b[c0:c0 + oc] = conv.bias.detach()
        c0 += oc
    torch.nn.Conv1d(g["cin"], g["cout"], g["kmax"])   # discarded, RNG only
    return w, b


def _bn_state(prefix: str, n: int, sd: "OrderedDict[str, torch.Tensor]"):
    sd[prefix + "weight"] = torch.ones(n)
    sd[prefix + "bias"] = torch.zeros(n)
    sd[prefix + "running_mean"] = torch.zeros(n)
    sd[prefix + "running_var"] = torch.ones(n)
    sd[prefix + "num_batches_tracked"] = torch.tensor(0, dtype=torch.long)


def init_extractor(lpl: List[LayerParams]) -> "OrderedDict[str, torch.Tensor]":
    """State of OS_CNN_res(lpl) (OS_CNN.py:183-205; Res_OS_layer builds the OS_block first,
    then the 1x1 shortcut conv, OS_CNN.py:173-174).  Keys follow the reference state_dict."""
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for i, layer in enumerate(lpl):
        w, b = _replay_layer_init(layer)
        sd[f"net_1.net.net.{i}.conv1d.weight"] = w
        sd[f"net_1.net.net.{i}.conv1d.bias"] = b
        _bn_state(f"net_1.net.net.{i}.bn.", w.shape[0], sd)
    cf = feature_channels(lpl)
    res = torch.nn.Conv1d(lpl[0][0][0], cf, 1)
    sd["net_1.res.conv1d.weight"] = res.weight.detach().clone()
    sd["net_1.res.conv1d.bias"] = res.bias.detach().clone()
    _bn_state("net_1.res.bn.", cf, sd)
    return sd


def init_classifier(lpl: List[LayerParams], n_class: int) -> "OrderedDict[str, torch.Tensor]":
    """State of OS_CNN(lpl, n_class) (OS_CNN.py:80-99): 3 OS layers, then nn.Linear."""
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for i, layer in enumerate(lpl):
        w, b = _replay_layer_init(layer)
        sd[f"net.{i}.conv1d.weight"] = w
        sd[f"net.{i}.conv1d.bias"] = b
        _bn_state(f"net.{i}.bn.", w.shape[0], sd)
    lin = torch.nn.Linear(feature_channels(lpl), n_class)
    sd["hidden.weight"] = lin.weight.detach().clone()
    sd["hidden.bias"] = lin.bias.detach().clone()
    return sd


# --------------------------------------------------------------------------------------
# forward  (OS_CNN/OS_CNN.py:67-77, 101-110, 155-180, 207-217)
# --------------------------------------------------------------------------------------
DENSE_WGRAD = False     # True: d/dW as the reference's autograd returns it, see masked_conv


OPERAND_ROUND = None    # e.g. torch.bfloat16: emulate an engine that rounds the convolution OPERANDS (see _RoundedConv)


def _q(t: torch.Tensor) -> torch.Tensor:
    return t.to(OPERAND_ROUND).to(t.dtype)


class _RoundedConv(torch.autograd.Function):
    """The masked convolution with its operands rounded to ``OPERAND_ROUND`` at exactly the points where a low-precision
    tensor-core engine rounds them, everything else (accumulation, BatchNorm, reductions) in the working precision:

      forward : y  = conv(q(x), q(W * mask)) + b
      backward: dX = dgrad(q(dY), q(W * mask));  dW = wgrad(q(dY), q(x)) * mask;  db = sum dY     (straight-through in q)

    With float64 working precision this isolates the effect of the operand rounding alone: a bf16 engine must agree with
    it far more closely than with the exact oracle (tests/test_gpu_fullsize.py)."""

    @staticmethod
    def forward(ctx, x, w, b, mask, pad_l, pad_r):
        xq = F.pad(_q(x), (pad_l, pad_r))
        wq = _q(w * mask)
        ctx.save_for_backward(xq, wq, mask)
        ctx.pads = (pad_l, pad_r, x.shape[-1])
        return F.conv1d(xq, wq, b)

    @staticmethod
    def backward(ctx, dy):
        xq, wq, mask = ctx.saved_tensors
        pad_l, pad_r, L = ctx.pads
        dyq = _q(dy)
        dx = torch.nn.grad.conv1d_input(xq.shape, wq, dyq)[..., pad_l:pad_l + L]
        dw = torch.nn.grad.conv1d_weight(xq, wq.shape, dyq) * mask
        return dx, dw, dy.sum(dim=(0, 2)), None, None, None


_MASKS = {}


def _mask_on(layer: LayerParams, like: torch.Tensor) -> torch.Tensor:
    """The layer's 0/1 mask on ``like``'s device and dtype -- kept, as the reference keeps ``weight_mask`` as a module
    attribute (OS_CNN.py:55-58), so that the timed baselines do not rebuild it every call."""
    key = (tuple(map(tuple, layer)), str(like.device), like.dtype)
    m = _MASKS.get(key)
    if m is None:
        m = _MASKS[key] = torch.from_numpy(build_mask(layer)).to(device=like.device, dtype=like.dtype)
    return m


def masked_conv(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, layer: LayerParams) -> torch.Tensor:
    """OS_CNN.py:68-71: W*mask, ConstantPad1d((Kmax-1)//2, Kmax//2), Conv1d.

    The reference masks ``weight.data`` (outside autograd) and convolves with the parameter itself, so its ``W.grad``
    is the unmasked dense gradient -- non-zero on masked taps.  By default this restatement differentiates
    through the mask (``grad * mask``, what the parity tests compare); with ``DENSE_WGRAD`` the value is still
    ``W * mask`` but the gradient passes straight to ``W``, which is what GradNorm's norms see
    (train_and_test.py:683-690)."""
    g = bank_geometry(layer)
    mask = _mask_on(layer, w)
    if OPERAND_ROUND is not None:
        return _RoundedConv.apply(x, w, b, mask, g["pad_l"], g["pad_r"])
    xp = F.pad(x, (g["pad_l"], g["pad_r"]))
    wm = w + (w * mask - w).detach() if DENSE_WGRAD else w * mask
    return F.conv1d(xp, wm, b)


def batch_norm(y: torch.Tensor, sd: Dict[str, torch.Tensor], prefix: str, training: bool,
               update_running: bool = True) -> torch.Tensor:
    """BatchNorm1d semantics (OS_CNN.py:65,72): train = batch mean and
    biased variance over (B, L), running stats updated with the unbiased variance; eval =
    running stats.  Written out explicitly (not F.batch_norm) so that it is a restatement."""
    gamma, beta = sd[prefix + "weight"], sd[prefix + "bias"]
    if training:
        n = y.shape[0] * y.shape[2]
        mean = y.mean(dim=(0, 2))
        var = ((y - mean[None, :, None]) ** 2).mean(dim=(0, 2))
        if update_running:
            with torch.no_grad():
                rm, rv = sd[prefix + "running_mean"], sd[prefix + "running_var"]
                rm.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * mean.detach().to(rm.dtype))
                rv.mul_(1 - BN_MOMENTUM).add_(BN_MOMENTUM * (var.detach() * n / (n - 1)).to(rv.dtype))
                sd[prefix + "num_batches_tracked"] += 1
    else:
        mean = sd[prefix + "running_mean"].to(y.dtype)
        var = sd[prefix + "running_var"].to(y.dtype)
    yhat = (y - mean[None, :, None]) / torch.sqrt(var[None, :, None] + BN_EPS)
    return yhat * gamma[None, :, None] + beta[None, :, None]


def os_layer(x, sd, prefix: str, layer: LayerParams, relu: bool, training: bool,
             update_running: bool = True):
    """build_layer_with_layer_parameter.forward, OS_CNN.py:67-77."""
    y = masked_conv(x, sd[prefix + "conv1d.weight"], sd[prefix + "conv1d.bias"], layer)
    z = batch_norm(y, sd, prefix + "bn.", training, update_running)
    return F.relu(z) if relu else z


def extractor_forward(sd, lpl: List[LayerParams], x: torch.Tensor, training: bool = True,
                      update_running: bool = True) -> torch.Tensor:
    """OS_CNN_res.forward with n_layers=1 (OS_CNN.py:207-217) = Res_OS_layer.forward
    (OS_CNN.py:176-180): relu(BN(conv1x1(x)) + OS_block(x)), last block layer without ReLU."""
    h = x
    for i, layer in enumerate(lpl):
        h = os_layer(h, sd, f"net_1.net.net.{i}.", layer, relu=(i != len(lpl) - 1),
                     training=training, update_running=update_running)
    if OPERAND_ROUND is not None:
        wr = sd["net_1.res.conv1d.weight"]
        r = _RoundedConv.apply(x, wr, sd["net_1.res.conv1d.bias"], torch.ones_like(wr), 0, 0)
    else:
        r = F.conv1d(x, sd["net_1.res.conv1d.weight"], sd["net_1.res.conv1d.bias"])   # kernel 1, pad (0,0)
    r = batch_norm(r, sd, "net_1.res.bn.", training, update_running)
    return F.relu(r + h)


def classifier_forward(sd, lpl: List[LayerParams], x: torch.Tensor, training: bool = True,
                       update_running: bool = True, few_shot: bool = False):
    """OS_CNN.forward (OS_CNN.py:101-110): 3 OS layers (all ReLU) -> mean over L -> Linear.
    Returns (logits, pooled)."""
    h = x
    for i, layer in enumerate(lpl):
        h = os_layer(h, sd, f"net.{i}.", layer, relu=True, training=training,
                     update_running=update_running)
    pooled = h.mean(dim=-1)                    # AdaptiveAvgPool1d(1) + squeeze(-1)
    if few_shot:
        return pooled, pooled
    return F.linear(pooled, sd["hidden.weight"], sd["hidden.bias"]), pooled


def host_argmax(logits: torch.Tensor) -> np.ndarray:
    """utils.py:34-36: prediction = np.argmax on the host over fp32 logits (first max wins)."""
    return np.argmax(logits.detach().cpu().numpy(), axis=1)


# --------------------------------------------------------------------------------------
# explicit backward formulas -- used by op-level kernel tests
# --------------------------------------------------------------------------------------
def conv_dgrad(dy: torch.Tensor, w: torch.Tensor, layer: LayerParams) -> torch.Tensor:
    """dX[b,ci,j] = sum_t sum_co W[co,ci,t]*mask * dY[b,co,j-t+pL]   (A1)."""
    g = bank_geometry(layer)
    mask = torch.from_numpy(build_mask(layer)).to(w.dtype)
    wt = (w * mask).flip(-1).transpose(0, 1)           # [Cin, Cout, Kmax], taps reversed
    dyp = F.pad(dy, (g["pad_r"], g["pad_l"]))
    return F.conv1d(dyp, wt)


def conv_wgrad(dy: torch.Tensor, x: torch.Tensor, layer: LayerParams) -> torch.Tensor:
    """dW[co,ci,t] = sum_b sum_l dY[b,co,l] X[b,ci,l+t-pL] on live taps, 0 elsewhere (A1, F4)."""
    g = bank_geometry(layer)
    L = x.shape[-1]
    xp = F.pad(x, (g["pad_l"], g["pad_r"]))
    dw = torch.stack([torch.einsum("bol,bil->oi", dy, xp[:, :, t:t + L]) for t in range(g["kmax"])], dim=-1)
    return dw * torch.from_numpy(build_mask(layer)).to(dw.dtype)


def bn_relu_backward(dz, y, gamma, mean, var, relu_mask, training: bool):
    """A2: d = dZ*[Z>0]; train: dY = g/sqrt(v+eps) (d - S1/N - yhat S2/N); eval: dY = g/sqrt(v+eps) d."""
    n = y.shape[0] * y.shape[2]
    d = dz * relu_mask
    inv = 1.0 / torch.sqrt(var + BN_EPS)
    yhat = (y - mean[None, :, None]) * inv[None, :, None]
    s1 = d.sum(dim=(0, 2))
    s2 = (d * yhat).sum(dim=(0, 2))
    if training:
        dy = (gamma * inv)[None, :, None] * (d - s1[None, :, None] / n - yhat * s2[None, :, None] / n)
    else:
        dy = (gamma * inv)[None, :, None] * d
    return dy, s2, s1      # dY, dgamma, dbeta


# --------------------------------------------------------------------------------------
# helpers for tests / bench
# --------------------------------------------------------------------------------------
def clone_state(sd, dtype=None, requires_grad: bool = False):
    out = OrderedDict()
    for k, v in sd.items():
        t = v.detach().clone()
        if t.is_floating_point():
            if dtype is not None:
                t = t.to(dtype)
            if requires_grad and not any(s in k for s in ("running_", "num_batches")):
                t.requires_grad_(True)
        out[k] = t
    return out


def synthetic_batch(B: int, C: int, L: int, n_class: int, domain_id: int = 0):
    """SURVEY 8d synthetic inputs: seeded randn, per-series z-normalisation over L, randint labels."""
    g = torch.Generator().manual_seed(1234 + domain_id)
    x = torch.randn(B, C, L, generator=g)
    x = (x - x.mean(-1, keepdim=True)) / x.std(-1, keepdim=True)
    y = torch.randint(0, n_class, (B,), generator=g)
    return x.float(), y.long()
