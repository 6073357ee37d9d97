"""Pins the oracle restatements of the 8f rows (voting, GradNorm) against vectors produced by executing the reference's own
source lines (oracle/make_golden.py --drivers -> tests/golden/{voting_small,gradnorm_small}.npz).  CPU only."""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import GOLDEN, as_lpl, rel_err
from oracle import grad_norm as GN
from oracle import os_cnn as O
from oracle import voting as V


@pytest.fixture(scope="module")
def voting():
    return np.load(os.path.join(GOLDEN, "voting_small.npz"))


@pytest.fixture(scope="module")
def gradnorm():
    return np.load(os.path.join(GOLDEN, "gradnorm_small.npz"))


def test_voting_oracle_matches_the_reference_script(voting):
    K = voting["train_logits1"].shape[1]
    precs = []
    for m in (1, 2, 3):
        p = V.class_precision(voting[f"train_logits{m}"], voting["label_list_train"], K)
        assert np.array_equal(p, voting[f"precision{m}"])           # integer counts, one float64 division: bit-exact
        precs.append(p)
    assert precs[0][K - 1] == 0.0                                   # a class no model predicts ...
    w = V.normalized_weights(precs)
    for m in (1, 2, 3):
        assert np.array_equal(w[m - 1], voting[f"weight{m}"])
    assert np.all(w[:, K - 1] == 0.0)                               # ... goes through 0/0 -> nan_to_num -> 0
    score, pred = V.entropy_vote([voting[f"test_logits{m}"] for m in (1, 2, 3)], w)
    assert score.dtype == np.float32
    assert rel_err(score, voting["score"]) < 1e-6
    assert np.array_equal(pred, voting["predict"])
    assert V.accuracy(score, voting["label_list"]) == float(voting["acc"])
    assert np.array_equal(V.host_argmax(np.array([[1.0, 3.0, 3.0], [2.0, 2.0, 1.0]])), [1, 0])     # first maximum wins


class OracleModule:
    """The reference's module interface over the functional oracle (state dict + layer list)."""

    def __init__(self, sd, lpl, kind):
        self.sd, self.lpl, self.kind, self.training = sd, lpl, kind, True

    def train(self):
        self.training = True

    def eval(self):
        self.training = False

    def __call__(self, x):
        # OS_CNN.py:68 masks the parameter in place at every forward (what an optimizer wrote on masked taps is discarded)
        prefix = "net_1.net.net." if self.kind == "fe" else "net."
        with torch.no_grad():
            for i, layer in enumerate(self.lpl):
                self.sd[f"{prefix}{i}.conv1d.weight"].mul_(torch.from_numpy(O.build_mask(layer)))
        fn = O.extractor_forward if self.kind == "fe" else O.classifier_forward
        return fn(self.sd, self.lpl, x, training=self.training)

    def hidden(self, x):
        return F.linear(x, self.sd["hidden.weight"], self.sd["hidden.bias"])

    def params(self):
        return [(k, v) for k, v in self.sd.items() if v.requires_grad]

    def last_block_params(self):
        """return_last_layer().parameters() (OS_CNN.py:218-220): the OS_block of the residual layer."""
        return [v for k, v in self.sd.items() if k.startswith("net_1.net.") and v.requires_grad]


def sync_state(mods, names, gradnorm, b):
    with torch.no_grad():
        for nm, m in zip(names, mods):
            for k, v in m.sd.items():
                if "num_batches" not in k:
                    v.copy_(torch.from_numpy(gradnorm[f"b{b}/after/{nm}/{k}"]))


def load_gradnorm_modules(gradnorm, meta):
    lpl, lpl_c = as_lpl(meta["lpl"]), as_lpl(meta["lpl_cls"])
    mods = []
    for nm in ("fe_t", "cl_t", "fe_s", "cl_s"):
        sd = {k.split("/", 2)[2]: torch.from_numpy(gradnorm[k].copy()) for k in gradnorm.files if k.startswith(f"init/{nm}/")}
        sd = O.clone_state(sd, requires_grad=True)
        mods.append(OracleModule(sd, lpl if nm.startswith("fe") else lpl_c, nm[:2]))
    return mods


def test_gradnorm_oracle_matches_the_reference_lines(gradnorm):
    meta = json.loads(str(gradnorm["meta"]))
    mods = load_gradnorm_modules(gradnorm, meta)
    names = ("fe_t", "cl_t", "fe_s", "cl_s")
    lrs = (0.001, 0.003, 0.001, 0.003)
    opts = [torch.optim.RMSprop([v for _, v in m.params()], lr=lr) for m, lr in zip(mods, lrs)]
    w_t = torch.nn.Parameter(torch.tensor(GN.INIT_T))
    w_s = torch.nn.Parameter(torch.tensor(GN.INIT_S))
    opt_t, opt_s = torch.optim.Adam([w_t], lr=GN.LR_T), torch.optim.Adam([w_s], lr=GN.LR_S)
    initial_t = initial_s = None
    O.DENSE_WGRAD = True
    try:
        for b in range(2):
            xt, yt = torch.from_numpy(gradnorm[f"b{b}/xt"]), torch.from_numpy(gradnorm[f"b{b}/yt"])
            xs, ys = torch.from_numpy(gradnorm[f"b{b}/xs"]), torch.from_numpy(gradnorm[f"b{b}/ys"])
            assert np.array_equal(w_t.detach().numpy(), gradnorm[f"b{b}/weights_t_before"])
            if b > 0:
                # RMSprop's first update is sign-like (lr * g / (sqrt(0.01 g^2) + eps)): rounding-level gradients (conv biases
                # in front of a BatchNorm) move by O(lr) with the sign of the noise, so batch b starts from the reference's
                # own post-step state; the GradNorm state (initial losses, balanced weights, their Adam moments) carries over
                sync_state(mods, names, gradnorm, b - 1)
            losses = GN.named_losses(mods, xt, yt, xs, ys, meta["style_weight"])
            k_tol = 1.0
            for k, v in losses.items():
                assert abs(float(v.detach()) - float(gradnorm[f"b{b}/loss/{k}"])) < k_tol * 2e-6 * max(1.0, abs(float(v.detach()))), k
            lt = [losses[k] for k in GN.LOSSES_T]
            ls = [losses[k] for k in GN.LOSSES_S]
            lv_t = np.array([float(v.detach()) for v in lt], dtype=np.float32)
            lv_s = np.array([float(v.detach()) for v in ls], dtype=np.float32)
            if initial_t is None:
                initial_t, initial_s = GN.sigmoid_np(lv_t), GN.sigmoid_np(lv_s)
            assert rel_err(initial_t, gradnorm[f"b{b}/initial_t"]) < 1e-6
            c = GN.remainder_coefficients(meta["cur_epoch"])
            remainder = sum(ci * losses[k] for ci, k in zip(c, GN.REMAINDER))
            total = torch.sum(w_t.detach() * torch.stack(lt)) + torch.sum(w_s.detach() * torch.stack(ls)) + remainder
            for o in opts:
                o.zero_grad()
            norms_t = GN.side_norms(lt, w_t, mods[0].last_block_params())
            norms_s = GN.side_norms(ls, w_s, mods[2].last_block_params())
            assert rel_err(norms_t.detach(), gradnorm[f"b{b}/norms_t"]) < k_tol * 1e-4
            assert rel_err(norms_s.detach(), gradnorm[f"b{b}/norms_s"]) < k_tol * 1e-4
            g_t, target_t = GN.weight_gradient(norms_t, w_t, lv_t, initial_t)
            g_s, target_s = GN.weight_gradient(norms_s, w_s, lv_s, initial_s)
            assert rel_err(target_t, gradnorm[f"b{b}/target_t"]) < k_tol * 1e-4
            assert rel_err(target_s, gradnorm[f"b{b}/target_s"]) < k_tol * 1e-4
            assert rel_err(g_t, gradnorm[f"b{b}/grad_w_t"]) < k_tol * 1e-4 and rel_err(g_s, gradnorm[f"b{b}/grad_w_s"]) < k_tol * 1e-4
            # the two backward passes of the reference = grad(total) + grad(remainder)
            (total + remainder).backward()
            for nm, m in zip(names, mods):
                for k, p in m.params():
                    ref = gradnorm[f"b{b}/grad/{nm}/{k}"]
                    assert rel_err(p.grad, ref) < k_tol * 2e-4 or np.abs(ref).max() < 1e-6, (nm, k)
            # masked taps of the reference gradient are NOT zero (F4): the dense restatement reproduces them
            k1 = "net_1.net.net.1.conv1d.weight"
            mask = O.build_mask(as_lpl(meta["lpl"])[1])
            assert np.abs(gradnorm[f"b{b}/grad/fe_t/{k1}"] * (1 - mask)).max() > 1e-4
            w_t.grad, w_s.grad = g_t, g_s
            opt_t.step(); opt_s.step()
            for o in opts:
                o.step()
            GN.renormalize_(w_t, GN.TOTAL_T)
            GN.renormalize_(w_s, GN.TOTAL_S)
            assert rel_err(w_t.detach(), gradnorm[f"b{b}/weights_t_after"]) < 1e-6
            assert rel_err(w_s.detach(), gradnorm[f"b{b}/weights_s_after"]) < 1e-6
            for nm, m in zip(names, mods):               # the RMSprop step itself (lr 0.001 / 0.003, train_and_test.py:97-101)
                for k, v in m.sd.items():
                    if "num_batches" in k or "conv1d.bias" in k:
                        continue
                    ref = gradnorm[f"b{b}/after/{nm}/{k}"]            # (mean error: single near-zero gradients may flip sign)
                    assert np.abs(v.detach().numpy() - ref).mean() < 1e-3 * np.abs(ref).mean() + 1e-7, (nm, k)
    finally:
        O.DENSE_WGRAD = False


def test_gradnorm_host_arithmetic_of_the_cuda_driver(gradnorm):
    """The host half of the CUDA driver (numpy, no GPU needed) from the reference's own norms: sum_p ||g_p|| = norm / |w|."""
    from feature_level_style_transfer_for_tsc_b200 import grad_norm as D
    assert D.remainder_coefficients(0) == (3, 3, 2, 2) and D.remainder_coefficients(30) == (1.5, 2, 1.8, 1.8)
    for side, names in (("t", GN.LOSSES_T), ("s", GN.LOSSES_S)):
        for b in range(2):
            w = gradnorm[f"b{b}/weights_{side}_before"]
            sums = gradnorm[f"b{b}/norms_{side}"] / np.abs(w)
            lv = np.array([gradnorm[f"b{b}/loss/{k}"] for k in names], dtype=np.float32)
            norms, target, grad = D._weight_gradient(w, sums, lv, gradnorm[f"b{b}/initial_{side}"], D.ALPHA)
            assert rel_err(norms, gradnorm[f"b{b}/norms_{side}"]) < 1e-6
            assert rel_err(target, gradnorm[f"b{b}/target_{side}"]) < 1e-5
            assert rel_err(grad, gradnorm[f"b{b}/grad_w_{side}"]) < 1e-5
