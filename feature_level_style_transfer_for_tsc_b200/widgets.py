"""Host-side mirror of the one ``widgets.py`` module the hot path needs: ``DimensionUnification`` (reference
widgets.py:66-78), which aligns source features to the target's (channels, length) before AdaIN.
It is a SURVEY 8(f) "next" row: a Linear over L and a 1x1 Conv1d, both dense GEMMs, left to torch/cuBLAS here
(same class name, constructor, attribute names and state_dict keys as the reference)."""
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
        # the 1x1 convolution as a plain fp32 matmul: cuDNN would run it in TF32 (torch's conv default), which the
        # CPU oracle does not, and the hot-path parity is judged in fp32 (SURVEY 8d "precision of the reference")
        conv = self.channel_unification
        return self.relu2(torch.matmul(conv.weight.squeeze(-1), h) + conv.bias[:, None])
