"""Host-side mirror of the one ``widgets.py`` module the hot path needs: ``DimensionUnification`` (reference
widgets.py:66-78), which aligns source features to the target's (channels, length) before AdaIN.
It is a SURVEY 8(f) "next" row: a Linear over L and a 1x1 Conv1d, both dense GEMMs, left to torch/cuBLAS here
(same class name, constructor, attribute names and state_dict keys as the reference)."""
import numpy as np
import torch
import torch.nn as nn


class DimensionUnification(nn.Module):
    def __init__(self, source_channel, target_channel, source_length, target_length):
        super(DimensionUnification, self).__init__()
        self.length_unification = nn.Linear(in_features=source_length, out_features=target_length)
        self.relu1 = nn.ReLU()
        self.channel_unification = nn.Conv1d(in_channels=source_channel, out_channels=target_channel, kernel_size=1)
        self.relu2 = nn.ReLU()

    def forward(self, source_feature):
        h = self.relu1(self.length_unification(source_feature))
        conv = self.channel_unification
        if torch.backends.cuda.matmul.allow_tf32 and h.is_cuda:
            # beside the bf16 tensor-core engine (the trainer allows TF32 there): the convolution proper -- cuDNN runs it
            # in TF32 on the NCL layout with the bias fused; the 2-D x 3-D matmul below costs two transposing copies of
            # the [B, C, L] activations forward and one backward (3 x 13 us per cfg2 step in the ncu launch list)
            return self.relu2(torch.nn.functional.conv1d(h, conv.weight, conv.bias))
        # fp32 engine: the 1x1 convolution as a plain fp32 matmul -- cuDNN would run it in TF32 (torch's conv default),
        # which the CPU oracle does not, and the hot-path parity is judged in fp32 (SURVEY 8d "precision of the reference")
        return self.relu2(torch.matmul(conv.weight.squeeze(-1), h) + conv.bias[:, None])


def calc_coeff(iter_num, high=1.0, low=0.0, alpha=2.0, max_iter=50.0):
    """Warm-up of the gradient-reversal strength: a sigmoid ramp from ``low`` (iteration 0) towards ``high``
    (reference widgets.py:12-13, there with the removed ``np.float``)."""
    ramp = 2.0 / (1.0 + np.exp(-alpha * iter_num / max_iter)) - 1.0
    return float(low + (high - low) * ramp)


def _reverse_gradient(x, coeff):
    """Identity forward, ``-coeff * grad`` backward (the reference registers a hook, widgets.py:8-10,121-122)."""
    if x.requires_grad:
        x.register_hook(lambda grad: grad * (-coeff))
    return x


def init_weights(m):
    """Initialisation of the critic (reference widgets.py:82-92; of its three branches only the ``Linear`` one can
    fire for ``AdversarialNetworkforCDAN``): Xavier-normal weight, zero bias.  Consumes the RNG like the reference."""
    if isinstance(m, nn.Linear):
        nn.init.xavier_normal_(m.weight)
        nn.init.zeros_(m.bias)


class AdversarialNetworkforCDAN(nn.Module):
    """The C-DAN critic (reference widgets.py:95-131): gradient-reversed input, Linear-ReLU-Dropout x2, Linear(.,1),
    and a warm-up schedule of the reversal strength that advances once per training-mode call.

    Beyond the reference surface: ``critic(x)`` is the MLP without the reversal hook and ``reversal_coefficients``
    advances the schedule for a fused multi-call evaluation -- used by ``C_DAN.CDAN``, whose backward kernel applies the
    reversals from ``coeff_buffer`` (device) so that a CUDA graph can be replayed while the schedule moves."""

    def __init__(self, in_feature, hidden_size):
        super(AdversarialNetworkforCDAN, self).__init__()
        self.ad_layer1 = nn.Linear(in_feature, hidden_size)
        self.ad_layer2 = nn.Linear(hidden_size, hidden_size)
        self.ad_layer3 = nn.Linear(hidden_size, 1)
        self.relu1 = nn.ReLU()
        self.relu2 = nn.ReLU()
        self.dropout1 = nn.Dropout(0.2)
        self.dropout2 = nn.Dropout(0.2)
        self.apply(init_weights)
        self.iter_num = -1
        self.alpha = 100.0
        self.low = 0.0
        self.high = 1.0
        self.max_iter = 20.0
        self.coeff = float(0.001)
        self.coeff_buffer = None          # device [3]: critic-input reversal of call 1, of call 2, entropy reversal
        self._coeff_host = None

    def _advance(self):
        if self.training:
            self.iter_num += 1
        if self.iter_num >= self.max_iter:
            self.iter_num = self.max_iter
        self.coeff = calc_coeff(self.iter_num, self.high, self.low, self.alpha, self.max_iter)
        return self.coeff

    def advance_schedule(self, calls=2):
        """Host side of ``calls`` consecutive forward calls: returns the coefficient of each call and, last, the value
        ``self.coeff`` holds afterwards (what C_DAN.py:67 reads for the entropy hooks)."""
        vals = [self._advance() for _ in range(calls)]
        return vals + [self.coeff]

    def reversal_coefficients(self, calls=2):
        dev = self.ad_layer1.weight.device
        if self.coeff_buffer is None or self.coeff_buffer.device != dev:
            self.coeff_buffer = torch.zeros(calls + 1, device=dev, dtype=torch.float32)
            self._coeff_host = None
        if dev.type == "cuda" and torch.cuda.is_current_stream_capturing():
            return self.coeff_buffer          # the graph's owner advances the schedule before every replay
        vals = self.advance_schedule(calls)
        if vals != self._coeff_host:
            self.coeff_buffer.copy_(torch.tensor(vals, dtype=torch.float32))
            self._coeff_host = vals
        return self.coeff_buffer

    def critic(self, x):
        x = self.dropout1(self.relu1(self.ad_layer1(x)))
        x = self.dropout2(self.relu2(self.ad_layer2(x)))
        return self.ad_layer3(x)

    def forward(self, x):
        coeff = self._advance()
        return self.critic(_reverse_gradient(x * 1.0, coeff))
