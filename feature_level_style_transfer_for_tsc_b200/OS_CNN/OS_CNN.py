"""OS-CNN modules -- drop-in mirror of the reference's ``OS_CNN/OS_CNN.py`` on the B200 kernels.

Same class names, constructor / forward signatures, attributes (``.net``, ``.net_1``, ``.res``, ``.hidden``,
``.conv1d``, ``.bn``, ``.weight_mask``, ``.length_before_classification``) and ``state_dict`` keys as the
reference, and the same consumption of the torch RNG at construction (so one seed gives bit-identical
initial parameters -- SURVEY.md appendix A5).  The arithmetic does not go through torch: every forward is one
``os_stack`` autograd node that runs the hand-written sm_100a kernels (masked multi-size Conv1d as an
implicit GEMM, BatchNorm statistics / apply, ReLU, shortcut add) on the c8 device layout.

There is no CPU path: inputs must be CUDA fp32 tensors and ``libtsc_b200.so`` must be built.
"""
import math
import os

import torch
import torch.nn as nn

from .. import _lib
from .. import ops
from .. import functional as _F
from ..functional import LayerSpec, StackSpec, direct_grads, os_stack


def calculate_mask_index(kernel_length_now, largest_kernel_lenght):
    """(left, right): the taps [left, right) a size-``kernel_length_now`` kernel occupies inside the common
    ``largest_kernel_lenght`` window (reference lines 9-12; even kernels lean right)."""
    right_zero = math.ceil((largest_kernel_lenght - 1) / 2) - math.ceil((kernel_length_now - 1) / 2)
    left_zero = largest_kernel_lenght - kernel_length_now - right_zero
    return left_zero, left_zero + kernel_length_now


def creat_mask(number_of_input_channel, number_of_output_channel, kernel_length_now, largest_kernel_lenght):
    """ones inside the live window, zeros outside; shape (arg0, arg1, largest) as in the reference (lines 15-20)."""
    left, right = calculate_mask_index(kernel_length_now, largest_kernel_lenght)
    mask = torch.zeros(number_of_input_channel, number_of_output_channel, largest_kernel_lenght)
    mask[:, :, left:right] = 1.0
    return mask.numpy()


def creak_layer_mask(layer_parameter_list):
    """mask, initial weight and bias of one OS layer (reference lines 23-43).  Every prime kernel is drawn
    from its own ``nn.Conv1d`` (its own fan-in bound) in list order, which is what fixes the RNG stream."""
    largest = layer_parameter_list[-1][-1]
    masks, weights, biases = [], [], []
    for (ic, oc, k) in layer_parameter_list:
        conv = nn.Conv1d(in_channels=ic, out_channels=oc, kernel_size=k)
        left, right = calculate_mask_index(k, largest)
        big = torch.zeros(oc, ic, largest)
        big[:, :, left:right] = conv.weight.detach()
        weights.append(big)
        biases.append(conv.bias.detach().clone())
        m = torch.zeros(oc, ic, largest)
        m[:, :, left:right] = 1.0
        masks.append(m)
    return (torch.cat(masks, 0).numpy().astype("float32"), torch.cat(weights, 0).numpy().astype("float32"),
            torch.cat(biases, 0).numpy().astype("float32"))


_DEFERRED_COUNTERS = None      # a list while a trainer batches the ``num_batches_tracked += 1`` of a whole step


def defer_batch_counters(on: bool):
    """on: collect the ``num_batches_tracked`` tensors of the BatchNorm layers that run in training mode instead of
    incrementing each with its own kernel; off: return the collected list (the caller does one
    ``torch._foreach_add_``).  Only valid for BatchNorm layers with a fixed momentum."""
    global _DEFERRED_COUNTERS
    got = _DEFERRED_COUNTERS
    _DEFERRED_COUNTERS = [] if on else None
    return got


def _bn_layer_spec(geom, bn, relu, zero_masked=True, dense_wgrad=False):
    """LayerSpec for one conv+BN pair; advances ``num_batches_tracked`` like nn.BatchNorm1d.forward."""
    use_batch_stats = bn.training or bn.running_mean is None
    momentum = 0.0
    if bn.training and bn.track_running_stats and bn.running_mean is not None:
        if _DEFERRED_COUNTERS is not None and bn.momentum is not None:
            _DEFERRED_COUNTERS.append(bn.num_batches_tracked)
        else:
            bn.num_batches_tracked.add_(1)
        momentum = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked.item())
    return LayerSpec(geom=geom, relu=relu, training=use_batch_stats, momentum=float(momentum), eps=float(bn.eps),
                     running_mean=bn.running_mean, running_var=bn.running_var, zero_masked=zero_masked,
                     dense_wgrad=bool(dense_wgrad))


def _bn_params(conv, bn):
    if bn.weight is None:
        raise RuntimeError("BatchNorm1d(affine=False) is not supported by the OS-CNN kernels")
    return [conv.weight, conv.bias, bn.weight, bn.bias]


def _run_stack(layers, x, shortcut=None, final_relu=False, pooled=False):
    specs, params = [], []
    for layer in layers:
        specs.append(_bn_layer_spec(layer.geometry, layer.bn, layer.relu_or_not_at_last_layer,
                                    dense_wgrad=getattr(layer, "dense_wgrad", False)))
        params += _bn_params(layer.conv1d, layer.bn)
    sc_spec = None
    if shortcut is not None:
        sc_spec = _bn_layer_spec(shortcut.geometry, shortcut.bn, False, zero_masked=False)
        params += _bn_params(shortcut.conv1d, shortcut.bn)
    # a layer wider than one TMEM accumulator tile (> 256 padded channels: the reference's layer recipe gives 336 at L = 64,
    # 560 at L = 32, train_and_test.py:38-53) takes its whole stack to the fp32 CUDA-core engine -- same C-ABI, same layouts
    wide = any(sp.geom.wide for sp in specs) or (sc_spec is not None and sc_spec.geom.wide)
    conv_eng = _lib.ENGINE_SIMT if wide else ops.get_engine("conv")
    wgrad_eng = _lib.ENGINE_SIMT if wide else ops.get_engine("wgrad")
    spec = StackSpec(layers=specs, shortcut=sc_spec, final_relu=final_relu, engine=conv_eng,
                     wgrad_engine=wgrad_eng, op_dtype=(_lib.TSC_F32 if wide else ops.op_dtype()), direct_grads=direct_grads(),
                     pooled=(pooled and shortcut is None and not wide and ops.engine_name() == "tcgen05" and _F.FUSED_PATH
                             and os.environ.get("TSC_NO_POOLED") != "1"))
    if x.dtype != torch.float32:
        raise RuntimeError(f"OS-CNN input must be float32 (the reference casts with .float()), got {x.dtype}")
    out = os_stack(spec, x.contiguous(), params)
    if pooled and not spec.pooled:
        out = out.mean(dim=-1)                    # engines without the fused pooled epilogue
    return out


def _own_head(hidden: nn.Linear, pooled: torch.Tensor) -> bool:
    """The fused head kernel covers the classifier shapes of the reference (<= 64 classes, <= 256 pooled channels, fp32)."""
    return (pooled.is_cuda and hidden.bias is not None and hidden.out_features <= 64 and hidden.in_features <= 1024
            and os.environ.get("TSC_TORCH_HEAD") != "1")


class build_layer_with_layer_parameter(nn.Module):
    """One OS layer: mask*W -> zero pad -> Conv1d(Cin, sum Cout, Kmax) -> BatchNorm1d -> optional ReLU
    (reference lines 46-77).

    ``dense_wgrad`` (beyond the reference surface; default from ``TSC_DENSE_WGRAD``, off): when true this layer's
    ``conv1d.weight.grad`` is the reference's unmasked gradient (non-zero on masked taps) instead of
    ``grad * mask``; ``functional.dense_wgrad()`` switches it per backward call."""

    dense_wgrad = os.environ.get("TSC_DENSE_WGRAD", "0") == "1"

    def __init__(self, layer_parameters, relu_or_not_at_last_layer=True, with_nvidia=True):
        super(build_layer_with_layer_parameter, self).__init__()
        self.relu_or_not_at_last_layer = relu_or_not_at_last_layer
        os_mask, init_weight, init_bias = creak_layer_mask(layer_parameters)
        out_channels, in_channels, max_kernel_size = os_mask.shape
        # not in the state_dict (the reference keeps it as a plain attribute); follows .cuda()/.to()
        self.register_buffer("weight_mask", torch.from_numpy(os_mask).float(), persistent=False)
        self.padding = nn.ConstantPad1d((int((max_kernel_size - 1) / 2), int(max_kernel_size / 2)), 0)
        # the big Conv1d draws (and discards) its own init: part of the reference's RNG stream
        self.conv1d = nn.Conv1d(in_channels=in_channels, out_channels=out_channels, kernel_size=max_kernel_size)
        self.conv1d.weight = nn.Parameter(torch.from_numpy(init_weight), requires_grad=True)
        self.conv1d.bias = nn.Parameter(torch.from_numpy(init_bias), requires_grad=True)
        self.bn = nn.BatchNorm1d(num_features=out_channels)
        self.geometry = ops.bank_geometry(layer_parameters)

    def forward(self, X):
        return _run_stack([self], X)


class OS_CNN(nn.Module):
    """Classifier head over extracted features: OS layers (all ReLU) -> global average pool -> Linear
    (reference lines 80-110).  Returns ``(logits, pooled)``; with ``few_shot`` both are the pooled features."""

    def __init__(self, layer_parameter_list, n_class, few_shot=False):
        super(OS_CNN, self).__init__()
        self.few_shot = few_shot
        self.layer_parameter_list = layer_parameter_list
        self.layer_list = [build_layer_with_layer_parameter(p) for p in layer_parameter_list]
        self.net = nn.Sequential(*self.layer_list)
        self.averagepool = nn.AdaptiveAvgPool1d(1)
        out_put_channel_numebr = sum(p[1] for p in layer_parameter_list[-1])
        self.hidden = nn.Linear(out_put_channel_numebr, n_class)
        self.length_before_classification = out_put_channel_numebr

    def forward(self, X):
        X_f = _run_stack(list(self.net), X, pooled=True)        # AdaptiveAvgPool1d(1) + squeeze(-1), fused into the stack
        if not self.few_shot:
            if _own_head(self.hidden, X_f):
                return _F.head_cross_entropy(X_f, self.hidden.weight, self.hidden.bias, None)[0], X_f
            return self.hidden(X_f), X_f
        return X_f.unsqueeze(-1), X_f

    def forward_loss(self, X, labels):
        """(logits, pooled, mean cross-entropy against ``labels``): ``forward`` followed by ``nn.CrossEntropyLoss()`` as the
        trainer applies it (train_and_test.py:593-603), with the Linear head and the loss in ONE kernel each way.  Beyond
        the reference surface; used by train_step."""
        if self.few_shot:
            raise RuntimeError("few_shot classifiers have no Linear head to train with a cross-entropy loss")
        X_f = _run_stack(list(self.net), X, pooled=True)
        if _own_head(self.hidden, X_f):
            logits, ce = _F.head_cross_entropy(X_f, self.hidden.weight, self.hidden.bias, labels.contiguous())
            return logits, X_f, ce
        logits = self.hidden(X_f)
        return logits, X_f, torch.nn.functional.cross_entropy(logits, labels)


class OS_block(nn.Module):
    """A chain of OS layers; the last one may skip its ReLU (reference lines 117-139)."""

    def __init__(self, layer_parameter_list, relu_or_not_at_last_layer=True):
        super(OS_block, self).__init__()
        self.layer_parameter_list = layer_parameter_list
        self.relu_or_not_at_last_layer = relu_or_not_at_last_layer
        n = len(layer_parameter_list)
        self.layer_list = [
            build_layer_with_layer_parameter(p, True if i != n - 1 else relu_or_not_at_last_layer)
            for i, p in enumerate(layer_parameter_list)]
        self.net = nn.Sequential(*self.layer_list)

    def forward(self, X):
        return _run_stack(list(self.net), X)


def layer_parameter_list_input_change(layer_parameter_list, input_channel):
    """Same list with layer 0 re-targeted to ``input_channel`` inputs (reference lines 142-152)."""
    return [[(input_channel, oc, k) for (_, oc, k) in layer] if i == 0 else layer
            for i, layer in enumerate(layer_parameter_list)]


class SampaddingConv1D_BN(nn.Module):
    """'same'-padded Conv1d + BatchNorm1d, used as the 1x1 shortcut (reference lines 155-166)."""

    def __init__(self, in_channels, out_channels, kernel_size):
        super(SampaddingConv1D_BN, self).__init__()
        self.padding = nn.ConstantPad1d((int((kernel_size - 1) / 2), int(kernel_size / 2)), 0)
        self.conv1d = nn.Conv1d(in_channels=in_channels, out_channels=out_channels, kernel_size=kernel_size)
        self.bn = nn.BatchNorm1d(num_features=out_channels)
        self.geometry = ops.dense_geometry(in_channels, out_channels, kernel_size)
        self.relu_or_not_at_last_layer = False

    def forward(self, X):
        return _run_stack([self], X)


class Res_OS_layer(nn.Module):
    """relu(BN(conv1x1(X)) + OS_block(X)) with the block's last layer un-activated (reference lines 169-180).
    One fused stack: the add and the ReLU happen in the BatchNorm-apply kernel of the two branches."""

    def __init__(self, layer_parameter_list, out_put_channel_numebr):
        super(Res_OS_layer, self).__init__()
        self.layer_parameter_list = layer_parameter_list
        self.net = OS_block(layer_parameter_list, False)
        self.res = SampaddingConv1D_BN(layer_parameter_list[0][0][0], out_put_channel_numebr, 1)

    def forward(self, X):
        return _run_stack(list(self.net.net), X, shortcut=self.res, final_relu=True)


class OS_CNN_res(nn.Module):
    """Feature extractor: ``n_layers`` residual OS layers, no pooling / classifier (reference lines 183-220)."""

    def __init__(self, layer_parameter_list, n_layers=1):
        super(OS_CNN_res, self).__init__()
        self.layer_parameter_list = layer_parameter_list
        self.n_layers = n_layers
        out_put_channel_numebr = sum(p[1] for p in layer_parameter_list[-1])
        new_layer_parameter_list = layer_parameter_list_input_change(layer_parameter_list, out_put_channel_numebr)
        self.net_1 = Res_OS_layer(layer_parameter_list, out_put_channel_numebr)
        self.net_list = [Res_OS_layer(new_layer_parameter_list, out_put_channel_numebr) for _ in range(n_layers - 1)]
        if self.n_layers > 1:
            self.net = nn.Sequential(*self.net_list)

    def forward(self, X):
        temp = self.net_1(X)
        if self.n_layers > 1:
            temp = self.net(temp)
        return temp

    def return_last_layer(self):
        """The OS_block of the (single) residual layer -- what GradNorm differentiates against
        (train_and_test.py:681-690)."""
        return self.net_1.net
