"""Multi-source ensemble vote -- the arithmetic of the reference's ``multi_source_voting.py`` script (lines 281-424) as
functions over the CUDA modules.

The reference is a script over three hard-coded checkpoints; its steps are:
  1. per model, logits of the target *training* split -> argmax -> per-class precision of the predictions (:281-357),
  2. precision / mean-over-models precision, NaN -> 0 (:358-367),
  3. per model, logits of the *test* split -> softmax -> entropy -> p * (1 + 120 exp(-H)) * 9 ** weight (:368-400),
  4. sum over the models, argmax, accuracy (:401-407).
Steps 1 and 3-4 are one kernel each (``tsc_class_precision``, ``tsc_entropy_vote``); the forward passes are the hot-path
kernels in eval mode.  Logits stay on the device; the host sees the final predictions only.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch

from . import ops
from .utils import _predictor_for

ENTROPY_GAIN = 120.0       # multi_source_voting.py:387
WEIGHT_BASE = 9.0          # :387 np.power(9, weight)


def collect_logits(modules, dataloader) -> Tuple[torch.Tensor, torch.Tensor]:
    """(logits [N, K] fp32, labels [N] int64), both on the device, for a loader of (x, y) batches (:281-293)."""
    outs, labels = [], []
    predict = _predictor_for(list(modules))
    with torch.no_grad():
        for _, (x, y) in enumerate(dataloader):
            x = x.float().cuda()
            outs.append(predict(x).clone())                          # the predictor's output buffer is reused
            labels.append(y.to(device=x.device, dtype=torch.int64))
    return torch.cat(outs).contiguous(), torch.cat(labels).contiguous()


def class_precision(train_logits: torch.Tensor, train_labels: torch.Tensor) -> torch.Tensor:
    """[K] fp64: of the series predicted as class k, the fraction labelled k; 0 if k is never predicted (:294-311)."""
    return ops.class_precision(train_logits.contiguous(), train_labels.contiguous())[2]


def entropy_vote(test_logits: Sequence[torch.Tensor], precisions: Sequence[torch.Tensor],
                 entropy_gain: float = ENTROPY_GAIN, weight_base: float = WEIGHT_BASE):
    """(score [N, K] fp32, pred [N] int32) of the weighted vote over M models (:358-423)."""
    lg = torch.stack([t.contiguous() for t in test_logits]).contiguous()
    pr = torch.stack([p.contiguous() for p in precisions]).contiguous()
    return ops.entropy_vote(lg, pr, entropy_gain, weight_base)


def vote(model_chains: Sequence[List], train_loader, test_loader):
    """The whole script for M (extractor, ..., classifier) chains already in eval mode: returns
    (pred [N] int32 device, accuracy, score [N, K])."""
    precisions, tests, labels = [], [], None
    for chain in model_chains:
        tr_logits, tr_labels = collect_logits(chain, train_loader)
        precisions.append(class_precision(tr_logits, tr_labels))
        te_logits, labels = collect_logits(chain, test_loader)
        tests.append(te_logits)
    score, pred = entropy_vote(tests, precisions)
    _, counts, _ = ops.class_precision(score, labels)              # argmax of the summed scores against the labels
    acc = float(counts[1].sum().item()) / int(labels.numel())
    return pred, acc, score
