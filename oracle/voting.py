"""Oracle of the evaluation / multi-source voting path, numpy on CPU.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference has no function for this: ``multi_source_voting.py`` is a script.  The restatement below follows its
arithmetic line by line and is **pinned** against the script itself: ``oracle/make_golden.py`` executes the unmodified
source lines 294-307 (per-class precision), 358-367 (weights) and 406-423 (entropy vote) of
``/root/reference/multi_source_voting.py`` on seeded logits and stores inputs and results in
``tests/golden/voting_small.npz``; ``tests/test_oracle_voting.py`` compares.  ``accuracy`` restates what
``utils.py:27-183`` does with ``sklearn.metrics.accuracy_score`` after a host ``np.argmax``.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np


def host_argmax(logits: np.ndarray) -> np.ndarray:
    """utils.py:36-37 / multi_source_voting.py:294: ``np.argmax(y_predict, axis=1)`` (first maximum)."""
    return np.argmax(np.asarray(logits), axis=1)


def accuracy(logits: np.ndarray, labels: np.ndarray) -> float:
    """utils.py:38-45: accuracy_score(predictions, labels) = mean(pred == label)."""
    return float(np.mean(host_argmax(logits) == np.asarray(labels)))


def class_precision(logits: np.ndarray, labels: np.ndarray, n_class: int) -> np.ndarray:
    """multi_source_voting.py:294-307: for each class i, of the training series *predicted* as i the fraction whose
    label is i; 0 when nothing was predicted as i.  Python-int division, i.e. float64."""
    pred = host_argmax(logits)
    labels = np.asarray(labels)
    out = np.zeros(n_class, dtype=np.float64)
    for i in range(n_class):
        n_pred = int(np.sum(pred == i))
        n_ok = int(np.sum((pred == i) & (labels == i)))
        out[i] = n_ok / n_pred if n_pred != 0 else 0
    return out


def normalized_weights(precisions: Sequence[np.ndarray]) -> np.ndarray:
    """multi_source_voting.py:358-367: every model's precision vector divided by the mean over the models, NaN (0/0: a
    class no model ever predicts) replaced by 0.  Returns [M, K] float64."""
    w = np.stack([np.asarray(p, dtype=np.float64) for p in precisions])
    avg = w[0].copy()
    for m in range(1, len(w)):
        avg = avg + w[m]
    avg = avg / len(w)
    with np.errstate(invalid="ignore", divide="ignore"):
        return np.nan_to_num(w / avg)


def entropy_vote(logits: Sequence[np.ndarray], weights: np.ndarray, entropy_gain: float = 120.0, weight_base: float = 9.0):
    """multi_source_voting.py:406-423.  logits: M arrays [N, K] float32; weights [M, K] float64 (normalized_weights).
    Per row: softmax in float32 without max subtraction, H = scipy.stats.entropy(p) (natural log of p / sum(p)),
    p * (1 + gain * exp(-H)) * base ** w_m stored back as float32; the M float32 arrays are added in order; argmax.
    Returns (score [N, K] float32, pred [N])."""
    total = None
    for m, lg in enumerate(logits):
        r = np.array(lg, dtype=np.float32, copy=True)
        for i in range(len(r)):
            r[i] = np.exp(r[i]) / np.sum(np.exp(r[i]))
            pk = r[i] / np.sum(r[i])
            with np.errstate(divide="ignore", invalid="ignore"):
                h = np.sum(np.where(pk > 0, -pk * np.log(pk), np.float32(0)), dtype=np.float32)
            r[i] = r[i] * (1 + np.float32(entropy_gain) * np.exp(-h)) * np.power(weight_base, weights[m])
        total = r if total is None else total + r
    return total, np.argmax(total, axis=1)
