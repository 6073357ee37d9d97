"""Evaluation helpers -- drop-in mirror of the reference's ``utils.py`` on the CUDA modules.

Same function names and argument lists as utils.py:9-182.  What differs is where the work happens: the reference
copies every batch's logits to the host, takes ``np.argmax`` there and grows two numpy arrays with ``np.concatenate``;
here every batch stays on the device, one ``tsc_class_precision`` launch per batch does the argmax and the
(predicted, correct) class counts, and ONE device-to-host copy at the end of the loader yields the accuracy
(= ``sklearn.metrics.accuracy_score``: correct / total).  The forward passes run under ``torch.no_grad()`` with whatever
train/eval mode the caller set (the reference calls ``.eval()`` first, train_and_test.py:785-789).

``with_nvidia=False`` raises: there is no CPU path.
"""
from __future__ import annotations

import os

import torch

from . import ops


def _state(obj):
    return obj.state_dict()


def save_target_classification_modules(target_feature_extraction_module, target_classification_module, cur_epoch):
    """utils.py:9-15 (same file name and keys: checkpoints are interchangeable with the reference's)."""
    os.makedirs("train_log", exist_ok=True)
    torch.save({
        'epoch': cur_epoch,
        'feature_extraction_state_dict': _state(target_feature_extraction_module),
        'classification_state_dict': _state(target_classification_module),
    }, "train_log/epoch_" + str(cur_epoch) + ".tar")


def save_source_classification_modules(source_feature_extraction_module, source_to_target_feature_trans,
                                       source_classification_module, cur_epoch):
    """utils.py:18-25"""
    os.makedirs("train_log", exist_ok=True)
    torch.save({
        'epoch': cur_epoch,
        'feature_extraction_state_dict': _state(source_feature_extraction_module),
        'source_to_target_feature_trans': _state(source_to_target_feature_trans),
        'classification_state_dict': _state(source_classification_module),
    }, "train_log/epoch_" + str(cur_epoch) + "_source.tar")


def predict_logits(modules, x):
    """logits of ``classifier(...(extractor(x)))`` for a chain of modules whose last one is an ``OS_CNN``."""
    h = x
    for m in modules[:-1]:
        h = m(h)
    return modules[-1](h)[0]


USE_GRAPHS = True       # replay one CUDA graph per batch shape in the evaluation loops (a forward pass is ~11 short launches)


class GraphedPredictor:
    """``predict_logits`` behind one CUDA graph per batch shape, for module chains that are entirely in eval mode: an
    evaluation pass at the reference's loader batch of 20 series is launch-bound when run eagerly.  A chain with any module
    in training mode runs eagerly (a captured training-mode pass would replay its warm-up's running-statistics updates,
    and ``momentum=None`` BatchNorm reads its counter on the host).  The graph contains the pack kernel and the convolutions
    derive the BatchNorm coefficients in their prologue, so it reads the live parameters and running statistics: a
    checkpoint loaded or a training step taken between two calls is seen by the next replay.  The returned logits are a
    static buffer, valid until the next call."""

    def __init__(self, modules, enabled=None):
        self.modules = list(modules)
        self.enabled = USE_GRAPHS if enabled is None else enabled
        self._graphs = {}

    def __call__(self, x):
        if not self.enabled or not x.is_cuda or any(m.training for mod in self.modules for m in mod.modules()):
            with torch.no_grad():
                return predict_logits(self.modules, x)
        key = tuple(x.shape)
        ent = self._graphs.get(key)
        if ent is None:
            static_x = x.clone()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side), torch.no_grad():          # warm-up: issue plans, cuBLAS workspaces, allocator
                for _ in range(2):
                    predict_logits(self.modules, static_x)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.no_grad(), torch.cuda.graph(graph):
                static_out = predict_logits(self.modules, static_x)
            ent = self._graphs[key] = (graph, static_x, static_out)
        graph, static_x, static_out = ent
        static_x.copy_(x)
        graph.replay()
        return static_out


_PREDICTORS = {}


def _predictor_for(modules) -> GraphedPredictor:
    """One predictor (and so one set of captured graphs) per module chain, kept across calls: the evaluation helpers
    run every few epochs on the same modules (train_and_test.py:783-789)."""
    import weakref
    key = tuple(id(m) for m in modules)
    ent = _PREDICTORS.get(key)
    if ent is None or any(r() is not m for r, m in zip(ent[0], modules)):
        ent = _PREDICTORS[key] = ([weakref.ref(m) for m in modules], GraphedPredictor(modules))
    return ent[1]


def loader_accuracy(modules, dataloader, with_nvidia=True):
    """accuracy over a loader of (x, y) batches; returns (accuracy, n_series)."""
    if not with_nvidia:
        raise RuntimeError("the tsc_b200 modules have no CPU path (with_nvidia=False)")
    total = None
    n = 0
    predict = _predictor_for(list(modules))
    with torch.no_grad():
        for _, (x, y) in enumerate(dataloader):
            x = x.float().cuda()
            y = y.to(device=x.device, dtype=torch.int64).contiguous()
            logits = predict(x).contiguous()
            _, counts, _ = ops.class_precision(logits, y)
            total = counts[1].sum() if total is None else total + counts[1].sum()
            n += int(x.shape[0])
    if n == 0:
        raise RuntimeError("empty dataloader")
    return float(total.item()) / n, n


def _report(str_out, log=True):
    if log:
        os.makedirs("numpy_saved_with_accuracy", exist_ok=True)
        with open("numpy_saved_with_accuracy/the_log.txt", "a", encoding='utf-8') as f:
            f.write(str_out + "\n")
    print(str_out)


def eval_model_testdata(target_feature_extraction_module, target_classification_module, test_dataloader, cur_epoch,
                        with_nvidia=True):
    """utils.py:27-51"""
    acc, _ = loader_accuracy([target_feature_extraction_module, target_classification_module], test_dataloader, with_nvidia)
    _report("epoch_num:" + str(cur_epoch) + " accuracy_for_test:" + str(acc))
    return acc


def eval_model_traindata(target_feature_extraction_module, target_classification_module, train_dataloader, cur_epoch,
                         with_nvidia=True):
    """utils.py:53-77"""
    acc, _ = loader_accuracy([target_feature_extraction_module, target_classification_module], train_dataloader, with_nvidia)
    _report("epoch_num:" + str(cur_epoch) + " accuracy_for_train:" + str(acc))
    return acc


def eval_source_model_traindata(source_feature_extraction_module, source_to_target_feature_trans, source_classification_module,
                                train_dataloader, cur_epoch, with_nvidia=True):
    """utils.py:79-103"""
    acc, _ = loader_accuracy([source_feature_extraction_module, source_to_target_feature_trans, source_classification_module],
                             train_dataloader, with_nvidia)
    _report("epoch_num:" + str(cur_epoch) + " accuracy_for_source_train:" + str(acc))
    return acc


def eval_source_model_testdata(source_feature_extraction_module, source_to_target_feature_trans, source_classification_module,
                               test_dataloader, cur_epoch, with_nvidia=True):
    """utils.py:105-129"""
    acc, _ = loader_accuracy([source_feature_extraction_module, source_to_target_feature_trans, source_classification_module],
                             test_dataloader, with_nvidia)
    _report("epoch_num:" + str(cur_epoch) + " accuracy_for_source_test:" + str(acc))
    return acc


def eval_target_model_being_pretrained(target_feature_extraction_module, target_classification_module, target_dataloader,
                                       cur_epoch, whether_test=False, with_nvidia=True):
    """utils.py:131-155 (prints only)"""
    acc, _ = loader_accuracy([target_feature_extraction_module, target_classification_module], target_dataloader, with_nvidia)
    _report("epoch_num:" + str(cur_epoch) + (" accuracy_for_test:" if whether_test else " accuracy_for_train:") + str(acc), log=False)
    return acc


def eval_source_model_being_pretrained(source_feature_extraction_module, source_to_target_feature_trans,
                                       source_classification_module, source_dataloader, cur_epoch, whether_test=False,
                                       with_nvidia=True):
    """utils.py:157-182 (prints only)"""
    acc, _ = loader_accuracy([source_feature_extraction_module, source_to_target_feature_trans, source_classification_module],
                             source_dataloader, with_nvidia)
    _report("epoch_num:" + str(cur_epoch) + (" accuracy_for_source_test:" if whether_test else " accuracy_for_source_train:")
            + str(acc), log=False)
    return acc
