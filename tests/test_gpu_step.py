"""The benchmarked training step (SURVEY 8d, cfg2 data flow) on the GPU against the oracle step on the CPU."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import os_cnn as O
from oracle import step as OS

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def T():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import feature_level_style_transfer_for_tsc_b200 as pkg
    pkg._lib.load()
    return pkg


def l2_rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


@pytest.mark.parametrize("use_graph", [False, True])
def test_two_training_steps_match_the_oracle_fp32_engine(T, use_graph):
    """fp32 engine: loss of step 1 and 2, and every parameter after 2 RMSprop steps.  (RMSprop's first steps
    are sign-like, lr*g/(|g|*sqrt(1-a)+eps): parameters with a near-zero gradient are ill conditioned, so the
    parameter check is an L2 one.)"""
    from feature_level_style_transfer_for_tsc_b200.train_step import StyleTransferModelSet, Trainer
    T.set_engine("simt")
    Ct, Lt, Kt, Cs, Ls, Ks, B = 3, 96, 3, 2, 80, 4, 6        # every layer <= 256 channels (TSC_MAX_CHANNELS)
    torch.manual_seed(0)
    model = StyleTransferModelSet(Ct, Lt, Kt, Cs, Ls, Ks).cuda()
    tr = Trainer(model, style_weight=50.0, use_graph=use_graph)
    oms = OS.ModelSet(Ct, Lt, Kt, Cs, Ls, Ks, seed=0)
    xt, yt = O.synthetic_batch(B, Ct, Lt, Kt, 0)
    xs, ys = O.synthetic_batch(B, Cs, Ls, Ks, 1)
    # RMSprop's first steps are sign-like (lr*g / (0.1|g| + eps)): an element whose gradient is rounding noise moves by
    # +-10 lr whatever the noise says.  The parameter check below therefore looks only at the well-conditioned elements,
    # |g| > 5 % of the tensor's largest |g| in the oracle's step-0 gradient.
    oms.set_requires_grad()
    OS.step_forward(oms, xt, yt, xs, ys, 50.0, training=True)["loss"].backward()
    solid = {}
    for gname, k, p in oms.trainable():
        if p.grad is not None:
            solid[(gname, k)] = p.grad.abs() > 0.05 * p.grad.abs().max()
            p.grad = None
    oms2 = OS.ModelSet(Ct, Lt, Kt, Cs, Ls, Ks, seed=0)        # the probe pass above advanced the BN running statistics
    oms = oms2
    for step in range(2):
        loss = float(tr.step(xt.cuda(), yt.cuda(), xs.cuda(), ys.cuda()))
        oloss = OS.train_step(oms, xt, yt, xs, ys, 50.0)
        # step 0 is pure forward parity; step 1 also sees the (sign-like, ill-conditioned) first RMSprop update
        assert abs(loss - oloss) < (2e-4 if step == 0 else 5e-3) * abs(oloss), (step, loss, oloss)
    torch.cuda.synchronize()
    for name, sd in oms.groups().items():
        got = getattr(model, name).state_dict()
        for k, v in sd.items():
            if "num_batches" in k:
                assert int(got[k]) == int(v)
            elif "running" in k:
                assert rel_err(got[k].cpu(), v) < 5e-3, (name, k)
            elif k.endswith("conv1d.bias") or (name, k) not in solid:
                continue            # zero gradient up to rounding: sign-like RMSprop makes this pure noise
            else:
                lr = OS.ModelSet.LRS[name]
                diff = (got[k].cpu() - v.detach()).abs()
                sel = solid[(name, k)]
                if k.endswith("conv1d.weight"):
                    # masked taps: exact zeros here; in the reference (and the oracle) they carry the garbage update of
                    # a garbage gradient until the next forward re-masks them -- compare live taps only
                    mask = getattr(getattr(model, name).get_submodule(k[:-len(".conv1d.weight")]), "weight_mask", None)
                    if mask is not None:
                        assert float((got[k].cpu() * (1 - mask.cpu())).abs().max()) == 0.0, (name, k)
                        sel = sel & (mask.cpu() > 0)
                if int(sel.sum()) == 0:
                    continue
                # two sign-like steps of about 10 lr and 7 lr: agreement to one lr on >= 90 % of the solid elements
                assert float((diff[sel] <= lr).float().mean()) >= 0.9, (name, k)
    T.set_engine("tcgen05")


def test_training_step_tensor_core_engine_tracks_the_oracle(T):
    from feature_level_style_transfer_for_tsc_b200.train_step import StyleTransferModelSet, Trainer
    T.set_engine("tcgen05")
    C, Ln, K, B = 9, 128, 6, 16
    torch.manual_seed(0)
    model = StyleTransferModelSet(C, Ln, K, C, Ln, K).cuda()
    tr = Trainer(model, use_graph=True)
    oms = OS.ModelSet(C, Ln, K, C, Ln, K, seed=0)
    xt, yt = O.synthetic_batch(B, C, Ln, K, 0)
    xs, ys = O.synthetic_batch(B, C, Ln, K, 1)
    oms.set_requires_grad()
    ref = OS.step_forward(oms, xt, yt, xs, ys, 1.0)
    model.train()
    with torch.no_grad():       # (an autograd graph kept alive from the default stream would break the capture below)
        out = model(xt.cuda(), yt.cuda(), xs.cuda(), ys.cuda(), 1.0)
    assert rel_err(out["tf"].detach().cpu(), ref["tf"].detach()) < 1e-2
    # AdaIN divides by the content row's sigma; rows that DimensionUnification's ReLU left almost constant have
    # sigma ~ sqrt(eps) and amplify the bf16 error of the features (conditioning, not a kernel error: the op-level
    # AdaIN test is at 1e-5) -- hence the wider bound on this one tensor
    assert rel_err(out["s2t"].detach().cpu(), ref["s2t"].detach()) < 8e-2
    assert rel_err(out["logits_t"].detach().cpu(), ref["logits_t"].detach()) < 1e-2
    assert abs(float(out["loss"]) - float(ref["loss"])) < 1e-2 * abs(float(ref["loss"]))
    ref_arg = O.host_argmax(ref["logits_t"])
    top2 = np.sort(ref["logits_t"].detach().numpy(), axis=1)
    decided = (top2[:, -1] - top2[:, -2]) > 4e-2 * np.abs(top2).max()
    assert np.array_equal(O.host_argmax(out["logits_t"])[decided], ref_arg[decided])
    losses = [float(tr.step(xt.cuda(), yt.cuda(), xs.cuda(), ys.cuda())) for _ in range(6)]
    torch.cuda.synchronize()
    assert T.ops.read_watchdog() == 0
    assert all(np.isfinite(losses)) and losses[-1] < losses[0]          # it trains
