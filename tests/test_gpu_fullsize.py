"""Full-size parity of the product engine (tcgen05, bf16 operands) against the fp64 oracle -- BASELINE.json's own sizes.

* cfg2: the whole training step at B = 128 per domain, C = 9, L = 128, 6 classes: features, generated features, logits,
  loss, argmax (with the number of excluded near-ties asserted small) and the L2 error of EVERY parameter gradient.
* cfg4: extractor + classifier forward and backward at B = 256, C = 3, L = 1024 (primes up to 89).

The measured errors are written to ``gpurun_out/r2_parity_fullsize.json`` (committed copy: ``profiles/``); the asserted
bounds are about 3x the values measured on B200 (stated next to each assert).  North-star tolerance for a bf16 tensor-core
path: 1e-2 on forward quantities.  Gradients go through train-mode BatchNorm backward, which subtracts two projections
from the incoming gradient: at B*L = 16384 positions that cancellation is benign (unlike at the toy batches of
test_gpu_modules.py), so the gradient bounds here are the meaningful ones.
"""
import json
import os

import numpy as np
import pytest
import torch

from conftest import ROOT, rel_err
from oracle import os_cnn as O
from oracle import step as OS

pytestmark = pytest.mark.gpu

OUT = os.path.join(ROOT, "gpurun_out", "r2_parity_fullsize.json")


@pytest.fixture(scope="module")
def T():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import feature_level_style_transfer_for_tsc_b200 as pkg
    pkg._lib.load()
    pkg.set_engine("tcgen05")
    return pkg


def l2_rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def _record(section, payload):
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    data = {}
    if os.path.exists(OUT):
        with open(OUT) as f:
            data = json.load(f)
    data[section] = payload
    with open(OUT, "w") as f:
        json.dump(data, f, indent=1, sort_keys=True)


def _argmax_agreement(got_logits, ref_logits, tol):
    """Rows whose fp64 top-2 gap is below 4*tol*max|logit| are 'undecided' (SURVEY section 4); returns
    (agree on decided rows, number of undecided rows, rows)."""
    ref = ref_logits.detach().cpu().numpy().astype(np.float64)
    got = got_logits.detach().float().cpu().numpy()
    top2 = np.sort(ref, axis=1)
    decided = (top2[:, -1] - top2[:, -2]) > 4 * tol * np.abs(ref).max()
    agree = bool(np.array_equal(np.argmax(got, 1)[decided], np.argmax(ref, 1)[decided]))
    return agree, int((~decided).sum()), int(ref.shape[0]), int((np.argmax(got, 1) != np.argmax(ref, 1)).sum())


def _to_fp64(ms):
    for name in ("fe_t", "cl_t", "fe_s", "du", "cl_s"):
        setattr(ms, name, O.clone_state(getattr(ms, name), torch.float64))


@pytest.mark.parametrize("style_weight", [1.0, 1.0e4])
def test_cfg2_step_at_full_batch_against_the_fp64_oracle(T, style_weight):
    from feature_level_style_transfer_for_tsc_b200.train_step import StyleTransferModelSet
    C, Ln, K, B = 9, 128, 6, 128
    torch.manual_seed(0)
    model = StyleTransferModelSet(C, Ln, K, C, Ln, K).cuda()
    model.train()
    oms = OS.ModelSet(C, Ln, K, C, Ln, K, seed=0)
    _to_fp64(oms)
    oms.set_requires_grad()
    xt, yt = O.synthetic_batch(B, C, Ln, K, 0)
    xs, ys = O.synthetic_batch(B, C, Ln, K, 1)
    ref = OS.step_forward(oms, xt.double(), yt, xs.double(), ys, style_weight, training=True)
    ref["loss"].backward()

    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True          # as the trainer runs the step beside the bf16 engine
    try:
        out = model(xt.cuda(), yt.cuda(), xs.cuda(), ys.cuda(), style_weight)
        out["loss"].backward()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    torch.cuda.synchronize()
    assert T.ops.read_watchdog() == 0

    fwd = {k: rel_err(out[k].detach().cpu(), ref[k].detach()) for k in ("tf", "ssf", "s2t", "logits_t", "logits_s")}
    fwd["loss"] = abs(float(out["loss"]) - float(ref["loss"])) / abs(float(ref["loss"]))
    fwd["l_style"] = abs(float(out["l_style"]) - float(ref["l_style"])) / abs(float(ref["l_style"]))
    arg = {}
    for k in ("logits_t", "logits_s"):
        agree, undecided, rows, flips = _argmax_agreement(out[k], ref[k], 1e-2)
        arg[k] = dict(agree_on_decided=agree, undecided_rows=undecided, rows=rows, raw_flips=flips)
    grads = {}
    for gname, sd in oms.groups().items():
        mod = getattr(model, gname)
        named = dict(mod.named_parameters())
        for k, v in sd.items():
            if v.grad is None:
                continue
            g = named[k].grad
            assert g is not None, (gname, k)
            if k.endswith("conv1d.bias"):
                continue        # d(conv bias) behind a train-mode BatchNorm: zero here, rounding noise (1e-10) in the oracle
            grads[f"{gname}.{k}"] = l2_rel(g.detach().cpu(), v.grad)
    worst = max(grads.values())
    _record(f"cfg2_B{B}_style{style_weight:g}", dict(forward_rel_err=fwd, argmax=arg, grad_l2_rel=grads,
                                                      grad_l2_rel_max=worst, engine="tcgen05 (bf16 operands, fp32 accumulate)",
                                                      oracle="oracle/step.py in float64"))
    # ---- bounds: forward <= 1e-2 (north_star); measured on B200: see profiles/r2_parity_fullsize.json ----
    assert fwd["tf"] < 1e-2 and fwd["ssf"] < 1e-2 and fwd["logits_t"] < 1e-2 and fwd["logits_s"] < 1e-2, fwd
    assert fwd["loss"] < 1e-2, fwd
    # AdaIN divides by the content row's sigma: rows DimensionUnification's ReLU left almost constant amplify the bf16
    # error of the features (conditioning of the operator, not of the kernel: its op-level test is at 1e-5)
    assert fwd["s2t"] < 8e-2, fwd
    for k, a in arg.items():
        assert a["agree_on_decided"], (k, a)
        assert a["undecided_rows"] <= a["rows"] // 4, (k, a)       # near-ties are counted, not hidden
    assert worst < GRAD_BOUND[style_weight], (worst, sorted(grads.items(), key=lambda kv: -kv[1])[:5])


# parameter-gradient L2 bounds of the cfg2 step (3x the worst tensor measured on B200, profiles/r2_parity_fullsize.json)
GRAD_BOUND = {1.0: 6e-2, 1.0e4: 1.5e-1}


def test_cfg4_long_series_forward_backward_against_the_fp64_oracle(T):
    from feature_level_style_transfer_for_tsc_b200.train_step import SingleDomainModelSet
    C, Ln, K, B = 3, 1024, 4, 256
    torch.manual_seed(0)
    model = SingleDomainModelSet(C, Ln, K).cuda()
    model.train()
    torch.manual_seed(0)
    lpl, lpl_c = O.trainer_layer_lists(C, Ln)
    fe = O.clone_state(O.init_extractor(lpl), torch.float64, requires_grad=True)
    cl = O.clone_state(O.init_classifier(lpl_c, K), torch.float64, requires_grad=True)
    x, y = O.synthetic_batch(B, C, Ln, K, 0)
    feat = O.extractor_forward(fe, lpl, x.double(), True)
    logits, _ = O.classifier_forward(cl, lpl_c, feat, True)
    loss = torch.nn.functional.cross_entropy(logits, y)
    loss.backward()

    xg = x.cuda()
    f_gpu = model.fe(xg)
    lg_gpu, _ = model.cl(f_gpu)
    l_gpu = torch.nn.functional.cross_entropy(lg_gpu, y.cuda())
    l_gpu.backward()
    torch.cuda.synchronize()
    assert T.ops.read_watchdog() == 0
    fwd = dict(features=rel_err(f_gpu.detach().cpu(), feat.detach()), logits=rel_err(lg_gpu.detach().cpu(), logits.detach()),
               loss=abs(float(l_gpu) - float(loss)) / abs(float(loss)))
    agree, undecided, rows, flips = _argmax_agreement(lg_gpu, logits, 1e-2)
    grads = {}
    for gname, sd, mod in (("fe", fe, model.fe), ("cl", cl, model.cl)):
        named = dict(mod.named_parameters())
        for k, v in sd.items():
            if not v.is_floating_point() or v.grad is None or k.endswith("conv1d.bias"):
                continue
            grads[f"{gname}.{k}"] = l2_rel(named[k].grad.detach().cpu(), v.grad)
    worst = max(grads.values())
    _record(f"cfg4_B{B}_L{Ln}", dict(forward_rel_err=fwd, argmax=dict(agree_on_decided=agree, undecided_rows=undecided, rows=rows,
                                                                       raw_flips=flips),
                                     grad_l2_rel=grads, grad_l2_rel_max=worst))
    assert fwd["features"] < 1e-2 and fwd["logits"] < 1e-2 and fwd["loss"] < 1e-2, fwd
    # random-init logits of this configuration are nearly tied (min top-2 gap 2.2e-4): most rows are
    # 'undecided' at a 1e-2 tolerance; the decided ones must agree and the count is recorded, not hidden
    assert agree
    assert worst < 1.5e-1, (worst, sorted(grads.items(), key=lambda kv: -kv[1])[:5])


def test_few_shot_classifier_returns_the_pooled_features(T):
    """OS_CNN(..., few_shot=True) (OS_CNN.py:106-107): no Linear, both outputs are the pooled features."""
    from feature_level_style_transfer_for_tsc_b200.OS_CNN.OS_CNN import OS_CNN
    lpl_e = O.generate_layer_parameter_list(1, 7, [216, 2160], 3)
    cf = O.feature_channels(lpl_e)
    lpl_c = O.layer_parameter_list_input_change(lpl_e, cf)
    torch.manual_seed(0)
    cl = OS_CNN(lpl_c, 4, few_shot=True).cuda()
    torch.manual_seed(0)
    ocl = O.clone_state(O.init_classifier(lpl_c, 4), torch.float64, requires_grad=True)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(32, cf, 64, generator=g)
    a, b = cl(x.cuda())
    oa, ob = O.classifier_forward(ocl, lpl_c, x.double(), True, few_shot=True)
    # the reference returns the un-squeezed pool output first ([B, C, 1]) and the squeezed one second
    assert tuple(a.shape) == (32, cf, 1) and tuple(b.shape) == (32, cf)
    assert rel_err(b.detach().cpu(), ob.detach()) < 1e-2 and rel_err(a.detach().cpu()[..., 0], oa.detach()) < 1e-2
    b.sum().backward()
    ob.sum().backward()
    gw = dict(cl.named_parameters())["net.2.conv1d.weight"].grad
    assert l2_rel(gw.cpu(), ocl["net.2.conv1d.weight"].grad) < 1e-1
