"""GPU parity of the 8f rows: the dense weight gradient, the GradNorm joint-stage driver, the evaluation helpers
and the multi-source vote -- against the vectors the reference's own lines produced (tests/golden/gradnorm_small.npz,
voting_small.npz) and against the oracle."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, as_lpl, rel_err
from oracle import grad_norm as GN
from oracle import os_cnn as O
from oracle import voting as V

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def T():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import feature_level_style_transfer_for_tsc_b200 as pkg
    pkg._lib.load()
    return pkg


@pytest.fixture(scope="module")
def voting():
    return np.load(os.path.join(GOLDEN, "voting_small.npz"))


@pytest.fixture(scope="module")
def gradnorm():
    return np.load(os.path.join(GOLDEN, "gradnorm_small.npz"))


def test_multi_l2norm_matches_torch(T):
    g = torch.Generator(device="cuda").manual_seed(3)
    big = torch.randn(1 << 20, device="cuda", generator=g)
    tensors = [torch.randn(228, 72, 31, device="cuda", generator=g), torch.randn(7, device="cuda", generator=g),
               big[1:1001],                                         # 4-byte aligned only: the scalar path
               torch.zeros(5, device="cuda"), torch.randn(1, device="cuda", generator=g),
               torch.randn(144, 3, device="cuda", generator=g).t()]      # non-contiguous: copied by the wrapper
    out = T.ops.multi_l2norm(tensors).cpu().numpy()
    ref = np.array([float(torch.linalg.vector_norm(t.double())) for t in tensors])
    assert out.shape == (len(tensors) + 1,)
    assert rel_err(out[:-1], ref) < 2e-6
    assert abs(out[-1] - ref.sum()) < 2e-6 * ref.sum()
    again = T.ops.multi_l2norm(tensors).cpu().numpy()
    assert np.array_equal(out, again)                                # fixed summation order
    with pytest.raises(RuntimeError):
        T.ops.multi_l2norm([])
    with pytest.raises(RuntimeError):
        T.ops.multi_l2norm([torch.zeros(3)])                         # CPU tensor: no fallback


def test_class_precision_and_vote_match_the_reference_script(T, voting):
    K = voting["train_logits1"].shape[1]
    labels_tr = torch.from_numpy(voting["label_list_train"].astype(np.int64)).cuda()
    precs = []
    for m in (1, 2, 3):
        lg = torch.from_numpy(voting[f"train_logits{m}"]).cuda()
        pred, counts, prec = T.ops.class_precision(lg, labels_tr)
        assert np.array_equal(pred.cpu().numpy(), np.argmax(voting[f"train_logits{m}"], axis=1))      # bit-exact argmax
        assert np.array_equal(prec.cpu().numpy(), voting[f"precision{m}"])                              # bit-exact (ints, one division)
        assert int(counts[0].sum()) == lg.shape[0]
        precs.append(prec)
    tests = [torch.from_numpy(voting[f"test_logits{m}"]).cuda() for m in (1, 2, 3)]
    from feature_level_style_transfer_for_tsc_b200 import multi_source_voting as MV
    score, pred = MV.entropy_vote(tests, precs)
    assert rel_err(score.cpu(), voting["score"]) < 1e-5             # fp32 softmax / entropy: expf, logf vs numpy's
    assert np.array_equal(pred.cpu().numpy(), voting["predict"])
    labels_te = torch.from_numpy(voting["label_list"].astype(np.int64)).cuda()
    _, counts, _ = T.ops.class_precision(score, labels_te)
    assert int(counts[1].sum()) / labels_te.numel() == float(voting["acc"])
    # ties: the first maximum wins, as numpy.argmax
    tie = torch.tensor([[1.0, 3.0, 3.0], [2.0, 2.0, 1.0]], device="cuda")
    assert T.ops.class_precision(tie)[0].tolist() == [1, 0]
    with pytest.raises(RuntimeError):
        T.ops.class_precision(torch.zeros(4, 65, device="cuda"))     # more classes than the kernel supports


def build_gradnorm_modules(T, gradnorm, meta, state="init", b=None):
    from feature_level_style_transfer_for_tsc_b200.OS_CNN.OS_CNN import OS_CNN, OS_CNN_res
    lpl, lpl_c = as_lpl(meta["lpl"]), as_lpl(meta["lpl_cls"])
    torch.manual_seed(meta["seed"])
    fe_t = OS_CNN_res(lpl); cl_t = OS_CNN(lpl_c, meta["n_class"])
    fe_s = OS_CNN_res(lpl); cl_s = OS_CNN(lpl_c, meta["n_class"])
    mods = (fe_t, cl_t, fe_s, cl_s)
    for nm, m in zip(("fe_t", "cl_t", "fe_s", "cl_s"), mods):
        for k, v in m.state_dict().items():
            if "num_batches" not in k:
                assert np.array_equal(v.numpy(), gradnorm[f"init/{nm}/{k}"]), (nm, k)      # same seed, same RNG stream (A5)
        m.cuda().train()
    return mods


def sync_state(mods, gradnorm, b):
    with torch.no_grad():
        for nm, m in zip(("fe_t", "cl_t", "fe_s", "cl_s"), mods):
            for k, v in m.state_dict().items():
                if "num_batches" not in k:
                    v.copy_(torch.from_numpy(gradnorm[f"b{b}/after/{nm}/{k}"]))


# fp32 engine: the arithmetic is pinned tightly.  bf16 tensor-core engine: the balanced losses include a Gram style loss (a
# difference of nearly equal Gram matrices, x 1e3) and gradients that pass a BatchNorm backward behind an average pool
# (one bf16 ulp is amplified ~100x), so only aggregate bounds are meaningful there; every kernel involved is
# checked op-level on identical inputs in tests/test_gpu_kernels.py.
@pytest.mark.parametrize("engine,tol_norm,tol_grad,tol_w", [("simt", 2e-4, 1e-3, 1e-5), ("tcgen05", 0.15, 0.3, 1e-4)])
def test_gradnorm_joint_stage_driver(T, gradnorm, engine, tol_norm, tol_grad, tol_w):
    """Two consecutive batches of train_and_test.py:646-766 on the CUDA modules: per-loss norms over the last block (dense
    weight gradients), GradNorm targets, weight gradients, balanced weights after Adam + renormalisation, and the
    parameter gradients the optimizers see (= grad(total) + grad(remainder))."""
    from feature_level_style_transfer_for_tsc_b200 import grad_norm as D
    from feature_level_style_transfer_for_tsc_b200 import functional as TF
    T.set_engine(engine)
    meta = json.loads(str(gradnorm["meta"]))
    mods = build_gradnorm_modules(T, gradnorm, meta)
    names = ("fe_t", "cl_t", "fe_s", "cl_s")
    opts = [torch.optim.RMSprop(m.parameters(), lr=lr) for m, lr in zip(mods, (0.001, 0.003, 0.001, 0.003))]
    critic = torch.nn.Linear(4, 3).cuda()
    drv = D.JointStageDriver(mods[0].return_last_layer(), mods[2].return_last_layer(), opts, clamps=[(critic, 0.0005)])
    masks = {i: O.build_mask(as_lpl(meta["lpl"])[i]) for i in range(3)}
    for b in range(2):
        if b > 0:
            sync_state(mods, gradnorm, b - 1)         # see tests/test_oracle_drivers.py: RMSprop's first step is sign-like
        xt, yt = torch.from_numpy(gradnorm[f"b{b}/xt"]).cuda(), torch.from_numpy(gradnorm[f"b{b}/yt"]).cuda()
        xs, ys = torch.from_numpy(gradnorm[f"b{b}/xs"]).cuda(), torch.from_numpy(gradnorm[f"b{b}/ys"]).cuda()
        assert rel_err(drv.t.weights.detach().cpu(), gradnorm[f"b{b}/weights_t_before"]) < 1e-5
        losses = GN.named_losses(mods, xt, yt, xs, ys, meta["style_weight"], adain=TF.adain, gram_style_loss=TF.gram_style_loss)
        for k, v in losses.items():
            ref = float(gradnorm[f"b{b}/loss/{k}"])
            assert abs(float(v.detach()) - ref) < tol_norm * max(0.05, abs(ref)), k
        info = drv.step(losses, meta["cur_epoch"])
        torch.cuda.synchronize()
        assert T.ops.read_watchdog() == 0
        for side in ("t", "s"):
            assert rel_err(info[f"norms_{side}"], gradnorm[f"b{b}/norms_{side}"]) < tol_norm, (b, side)
            assert rel_err(info[f"target_{side}"], gradnorm[f"b{b}/target_{side}"]) < tol_norm, (b, side)
            assert rel_err(info[f"grad_w_{side}"], gradnorm[f"b{b}/grad_w_{side}"]) < tol_norm, (b, side)
            w = getattr(drv, side).weights.detach().cpu().numpy()
            assert rel_err(w, gradnorm[f"b{b}/weights_{side}_after"]) < tol_w, (b, side)
            assert abs(w.sum() - (7.0 if side == "t" else 8.0)) < 1e-5
        assert rel_err(drv.t.initial, gradnorm[f"b{b}/initial_t"]) < tol_norm
        # what the optimizers saw: the main backward runs with masked gradients (grad * mask, exact zeros elsewhere)
        for nm, m in zip(names, mods):
            for k, p in m.named_parameters():
                ref = gradnorm[f"b{b}/grad/{nm}/{k}"]
                g = p.grad.cpu().numpy()
                if k.endswith("conv1d.weight") and "res" not in k:
                    i = int(k.split(".conv1d")[0].split(".")[-1])
                    mask = masks[i] if nm.startswith("fe") else O.build_mask(as_lpl(meta["lpl_cls"])[i])
                    assert np.abs(g * (1 - mask)).max() == 0.0
                    ref = ref * mask
                if k.endswith("conv1d.bias"):
                    continue                                          # ~0 behind a BatchNorm (rounding noise in the reference)
                err = np.linalg.norm((g - ref).ravel()) / max(np.linalg.norm(ref.ravel()), 1e-12)
                assert err < tol_grad or np.abs(ref).max() < 1e-6, (b, nm, k, err)
        assert float(critic.weight.abs().max()) <= 0.0005 + 1e-9      # WGAN clamp (:763-766)


def _joint_stage_setup(T, gradnorm, capturable):
    from feature_level_style_transfer_for_tsc_b200 import grad_norm as D
    meta = json.loads(str(gradnorm["meta"]))
    mods = build_gradnorm_modules(T, gradnorm, meta)
    opts = [torch.optim.RMSprop(m.parameters(), lr=lr, capturable=capturable) for m, lr in zip(mods, (0.001, 0.003, 0.001, 0.003))]
    drv = D.JointStageDriver(mods[0].return_last_layer(), mods[2].return_last_layer(), opts, capturable=capturable)
    return meta, mods, drv


def test_joint_stage_device_path_matches_the_host_path(T, gradnorm):
    """JointStageDriver.step_device (no host round trip: device-side GradNorm weight gradient, remainder multipliers in a device
    tensor) against JointStageDriver.step -- the path pinned by the reference-executed golden vectors -- for the first batch:
    same norms / targets / weight gradients, the same parameter gradients, the same balanced weights afterwards."""
    from feature_level_style_transfer_for_tsc_b200 import grad_norm as D
    from feature_level_style_transfer_for_tsc_b200 import functional as TF
    T.set_engine("simt")
    xt, yt = torch.from_numpy(gradnorm["b0/xt"]).cuda(), torch.from_numpy(gradnorm["b0/yt"]).cuda()
    xs, ys = torch.from_numpy(gradnorm["b0/xs"]).cuda(), torch.from_numpy(gradnorm["b0/ys"]).cuda()
    meta, mods_a, drv_a = _joint_stage_setup(T, gradnorm, False)
    _, mods_b, drv_b = _joint_stage_setup(T, gradnorm, True)
    la = GN.named_losses(mods_a, xt, yt, xs, ys, meta["style_weight"], adain=TF.adain, gram_style_loss=TF.gram_style_loss)
    info_a = drv_a.step(la, meta["cur_epoch"])
    lb = GN.named_losses(mods_b, xt, yt, xs, ys, meta["style_weight"], adain=TF.adain, gram_style_loss=TF.gram_style_loss)
    coef = torch.tensor(D.remainder_coefficients(meta["cur_epoch"]), dtype=torch.float32, device="cuda")
    info_b = drv_b.step_device(lb, coef)
    torch.cuda.synchronize()
    for k in ("norms_t", "norms_s", "target_t", "target_s", "grad_w_t", "grad_w_s", "loss_t", "loss_s"):
        assert rel_err(info_b[k].cpu().numpy(), info_a[k]) < 1e-5, k
    for side in ("t", "s"):
        assert rel_err(getattr(drv_b, side).weights.detach().cpu(), getattr(drv_a, side).weights.detach().cpu()) < 1e-5
        assert rel_err(getattr(drv_b, side).initial, getattr(drv_a, side).initial) < 1e-6
    for ma, mb in zip(mods_a, mods_b):
        for (k, pa), (_, pb) in zip(ma.named_parameters(), mb.named_parameters()):
            ga, gb = pa.grad.float(), pb.grad.float()             # the same kernels on the same values
            assert float((ga - gb).abs().max()) <= 1e-6 * max(1e-6, float(ga.abs().max())), k
    T.set_engine("tcgen05")


def test_graphed_joint_stage_follows_the_eager_steps(T, gradnorm):
    """GraphedJointStage on the product engine: five joint-stage steps on five different batches -- two eager, the capture, two
    more replays -- against the same five steps run eagerly through step_device on an identical module set.  A replay that read
    stale inputs, skipped an optimizer or lost the epoch's multipliers would be off by O(1); what is allowed is the rounding
    difference of library kernels chosen under capture (RMSprop's first steps are sign-like and amplify it)."""
    from feature_level_style_transfer_for_tsc_b200 import grad_norm as D
    from feature_level_style_transfer_for_tsc_b200 import functional as TF
    T.set_engine("tcgen05")
    meta, mods_a, drv_a = _joint_stage_setup(T, gradnorm, True)
    _, mods_b, drv_b = _joint_stage_setup(T, gradnorm, True)
    sw = meta["style_weight"]

    def loss_fn_b(xt, yt, xs, ys):
        return GN.named_losses(mods_b, xt, yt, xs, ys, sw, adain=TF.adain, gram_style_loss=TF.gram_style_loss)

    stage = D.GraphedJointStage(drv_b, loss_fn_b)
    B, C, Ln = gradnorm["b0/xt"].shape
    K = meta["n_class"]
    epochs = [0, 0, 0, 13, 30]                                        # the multipliers change while ONE graph is replayed
    hist_a, hist_b = [], []
    for i, ep in enumerate(epochs):
        xt, yt = O.synthetic_batch(B, C, Ln, K, 40 + 2 * i)
        xs, ys = O.synthetic_batch(B, C, Ln, K, 41 + 2 * i)
        xt, yt, xs, ys = xt.cuda(), yt.cuda(), xs.cuda(), ys.cuda()
        la = GN.named_losses(mods_a, xt, yt, xs, ys, sw, adain=TF.adain, gram_style_loss=TF.gram_style_loss)
        coef = torch.tensor(D.remainder_coefficients(ep), dtype=torch.float32, device="cuda")
        ia = drv_a.step_device(la, coef)
        ib = stage.step(xt, yt, xs, ys, cur_epoch=ep)
        torch.cuda.synchronize()
        hist_a.append({k: v.detach().cpu().numpy().copy() for k, v in ia.items()})
        hist_b.append({k: v.detach().cpu().numpy().copy() for k, v in ib.items()})
    assert stage._graph is not None and stage.calls == 5
    assert T.ops.read_watchdog() == 0
    for i, (a, b) in enumerate(zip(hist_a, hist_b)):
        for k in ("loss_t", "loss_s", "norms_t", "norms_s"):
            assert rel_err(b[k], a[k]) < 5e-2, (i, k, a[k], b[k])
    for side in ("t", "s"):
        wa, wb = getattr(drv_a, side).weights.detach().cpu().numpy(), getattr(drv_b, side).weights.detach().cpu().numpy()
        assert rel_err(wb, wa) < 1e-2 and abs(wb.sum() - (7.0 if side == "t" else 8.0)) < 1e-4, (side, wa, wb)
    moved = 0.0
    for nm, ma, mb in zip(("fe_t", "cl_t", "fe_s", "cl_s"), mods_a, mods_b):
        for (k, pa), (_, pb) in zip(ma.named_parameters(), mb.named_parameters()):
            if k.endswith("conv1d.bias"):
                continue                                              # gradient ~0 behind a BatchNorm: pure sign noise under RMSprop
            init = torch.from_numpy(gradnorm[f"init/{nm}/{k}"]).cuda()
            step_size = float((pa.detach() - init).norm())
            moved += step_size
            assert float((pa.detach() - pb.detach()).norm()) <= 0.25 * step_size + 1e-6, (nm, k)
    assert moved > 0.0                                                # the optimizers really stepped inside the graph


@pytest.mark.parametrize("engine,tol", [("simt", 3e-4), ("tcgen05", 0.1)])
def test_dense_wgrad_reproduces_the_reference_masked_tap_gradients(T, gradnorm, engine, tol):
    """F4: with dense_wgrad the kernel-bank gradient equals the reference's autograd result on EVERY tap."""
    from feature_level_style_transfer_for_tsc_b200 import functional as TF
    T.set_engine(engine)
    meta = json.loads(str(gradnorm["meta"]))
    mods = build_gradnorm_modules(T, gradnorm, meta)
    xt, yt = torch.from_numpy(gradnorm["b0/xt"]).cuda(), torch.from_numpy(gradnorm["b0/yt"]).cuda()
    xs, ys = torch.from_numpy(gradnorm["b0/xs"]).cuda(), torch.from_numpy(gradnorm["b0/ys"]).cuda()
    losses = GN.named_losses(mods, xt, yt, xs, ys, meta["style_weight"], adain=TF.adain, gram_style_loss=TF.gram_style_loss)
    c = GN.remainder_coefficients(0)
    remainder = sum(ci * losses[k] for ci, k in zip(c, GN.REMAINDER))
    total = (2 * losses["target_nf_loss"] + 5 * losses["target_classification_loss"] + 2 * losses["source_nf_loss"]
             + 2 * losses["source_classification_loss"] + 4 * losses["s2t2s_classification_loss"] + 2 * remainder)
    block = mods[0].return_last_layer()
    with TF.dense_wgrad():
        grads = torch.autograd.grad(total, list(block.parameters()), retain_graph=True)
    for (k, _), g in zip(block.named_parameters(), grads):
        ref = gradnorm[f"b0/grad/fe_t/net_1.net.{k}"]
        if k.endswith("conv1d.bias"):
            continue
        err = np.linalg.norm((g.cpu().numpy() - ref).ravel()) / np.linalg.norm(ref.ravel())
        assert err < tol * 3, (k, err)
        if k.endswith("conv1d.weight"):
            i = int(k.split(".")[1])
            mask = O.build_mask(as_lpl(meta["lpl"])[i])
            off = (1 - mask).astype(bool)
            if off.any():
                ref_off = ref[off]
                assert np.abs(ref_off).max() > 1e-5                    # the reference's "garbage" is really there
                e2 = np.linalg.norm(g.cpu().numpy()[off] - ref_off) / np.linalg.norm(ref_off)
                assert e2 < tol * 3, (k, e2)
    # without the switch the same graph gives exact zeros on the masked taps
    grads0 = torch.autograd.grad(total, list(block.parameters()))
    for (k, _), g in zip(block.named_parameters(), grads0):
        if k.endswith("conv1d.weight"):
            mask = O.build_mask(as_lpl(meta["lpl"])[int(k.split(".")[1])])
            assert np.abs(g.cpu().numpy() * (1 - mask)).max() == 0.0


def test_eval_helpers_and_vote_end_to_end(T, tables, small_pair, tmp_path, monkeypatch):
    """utils.eval_* and multi_source_voting.vote over the CUDA modules in eval mode: the logits are the golden eval logits of
    the reference modules, the vote equals the oracle's vote over the same logits."""
    from feature_level_style_transfer_for_tsc_b200 import multi_source_voting as MV
    from feature_level_style_transfer_for_tsc_b200 import utils as U
    from feature_level_style_transfer_for_tsc_b200.OS_CNN.OS_CNN import OS_CNN, OS_CNN_res
    T.set_engine("simt")
    monkeypatch.chdir(tmp_path)
    meta = tables["small"]
    lpl_e, lpl_c = as_lpl(meta["lpl_ext"]), as_lpl(meta["lpl_cls"])
    out = small_pair["out"]
    chains = []
    for j in range(3):
        torch.manual_seed(meta["seed"] + j)
        fe, cl = OS_CNN_res(lpl_e).cuda().eval(), OS_CNN(lpl_c, meta["n_class"]).cuda().eval()
        chains.append([fe, cl])
    x, y = torch.from_numpy(out["x"]), torch.from_numpy(out["y"])
    loader = [(x[:4], y[:4]), (x[4:], y[4:])]                          # ragged last batch
    # model 0 = the golden model in its *initial* state: eval BN with fresh running statistics
    lg0, lab = MV.collect_logits(chains[0], loader)
    sd_fe = {k: torch.from_numpy(v.copy()) for k, v in small_pair["init_fe"].items()}
    sd_cl = {k: torch.from_numpy(v.copy()) for k, v in small_pair["init_cl"].items()}
    ref_lg = O.classifier_forward(sd_cl, lpl_c, O.extractor_forward(sd_fe, lpl_e, x, training=False), training=False)[0]
    assert rel_err(lg0.cpu(), ref_lg.detach()) < 2e-4
    assert np.array_equal(lab.cpu().numpy(), out["y"])
    acc = U.eval_model_testdata(chains[0][0], chains[0][1], loader, 3)
    assert acc == V.accuracy(lg0.cpu().numpy(), out["y"])
    assert "epoch_num:3 accuracy_for_test:" in open("numpy_saved_with_accuracy/the_log.txt").read()
    assert U.eval_target_model_being_pretrained(chains[0][0], chains[0][1], loader, 0, whether_test=True) == acc
    with pytest.raises(RuntimeError):
        U.loader_accuracy(chains[0], loader, with_nvidia=False)
    pred, acc_v, score = MV.vote(chains, loader, loader)
    logits = [MV.collect_logits(c, loader)[0].cpu().numpy() for c in chains]
    precs = [V.class_precision(l, out["y"], meta["n_class"]) for l in logits]
    ref_score, ref_pred = V.entropy_vote(logits, V.normalized_weights(precs))
    assert rel_err(score.cpu(), ref_score) < 1e-5
    assert np.array_equal(pred.cpu().numpy(), ref_pred)
    assert acc_v == float(np.mean(ref_pred == out["y"]))
    U.save_target_classification_modules(chains[0][0], chains[0][1], 2)
    ck = torch.load("train_log/epoch_2.tar")
    assert set(ck) == {"epoch", "feature_extraction_state_dict", "classification_state_dict"}
    assert list(ck["feature_extraction_state_dict"].keys()) == list(small_pair["init_fe"].keys())


@pytest.mark.parametrize("C,Ln,B,K", [(3, 32, 6, 4), (9, 128, 16, 6), (1, 160, 4, 3), (3, 400, 3, 5)])
def test_inference_path_equals_the_training_kernels_in_eval_mode(T, C, Ln, B, K):
    """Forward-only eval calls fold BatchNorm / ReLU / shortcut / pooling into the convolution epilogue: same numbers as the
    conv -> bn_apply kernels of the training path (bf16 operands in both; a rounding flip of an intermediate is the only
    difference), and within the tensor-core tolerance of the fp32 oracle.  L = 160 / 400 take the two-kernel pooled tail
    (more than one CTA per sample) and a ragged last tile."""
    from feature_level_style_transfer_for_tsc_b200 import functional as TF
    from feature_level_style_transfer_for_tsc_b200.OS_CNN.OS_CNN import OS_CNN, OS_CNN_res
    T.set_engine("tcgen05")
    lpl_e, lpl_c = O.trainer_layer_lists(C, Ln) if Ln >= 100 else (as_lpl([[(C, 4, 1), (C, 4, 2), (C, 4, 3)],
                                                                           [(12, 6, 1), (12, 6, 2), (12, 6, 3)],
                                                                           [(18, 10, 1), (18, 10, 2)]]), None)
    if lpl_c is None:
        lpl_c = O.layer_parameter_list_input_change(lpl_e, O.feature_channels(lpl_e))
    torch.manual_seed(2)
    fe, cl = OS_CNN_res(lpl_e).cuda(), OS_CNN(lpl_c, K).cuda()
    x, _ = O.synthetic_batch(B, C, Ln, K, 3)
    xd = x.cuda()
    fe.train(); cl.train()
    with torch.no_grad():                      # a few training-mode passes give the running statistics real values
        for _ in range(3):
            cl(fe(xd))
    sd_fe = {k: v.detach().cpu().clone() for k, v in fe.state_dict().items()}
    sd_cl = {k: v.detach().cpu().clone() for k, v in cl.state_dict().items()}
    fe.eval(); cl.eval()
    with torch.no_grad():
        feat = fe(xd)
        logits, pooled = cl(feat)
        TF.INFERENCE_PATH = False
        try:
            feat0 = fe(xd)
            logits0, pooled0 = cl(feat0)
        finally:
            TF.INFERENCE_PATH = True
    torch.cuda.synchronize()
    assert T.ops.read_watchdog() == 0
    assert feat.shape == feat0.shape and logits.shape == (B, K) and pooled.shape == pooled0.shape
    assert rel_err(feat.cpu(), feat0.cpu()) < 5e-3
    assert rel_err(pooled.cpu(), pooled0.cpu()) < 5e-3
    assert rel_err(logits.cpu(), logits0.cpu()) < 5e-3
    ref_feat = O.extractor_forward(sd_fe, lpl_e, x, training=False)
    ref_logits, ref_pooled = O.classifier_forward(sd_cl, lpl_c, ref_feat, training=False)
    assert rel_err(feat.cpu(), ref_feat) < 2e-2                                # tcgen05 engine: bf16 operands, 1e-2 class
    assert rel_err(pooled.cpu(), ref_pooled) < 2e-2
    assert rel_err(logits.cpu(), ref_logits) < 2e-2
    # with autograd on and parameters that require grad the same call must build a graph (the joint stage's eval-BN
    # classifier call, train_and_test.py:584-586): not the inference path
    out = cl(feat.requires_grad_(True))[0]
    assert out.requires_grad


def test_driver_kernels_edge_cases(T):
    """Smallest and largest supported shapes, degenerate rows, and the numpy corner cases the reference inherits
    (softmax without max subtraction overflows to NaN for logits > 88; numpy.argmax returns the first NaN)."""
    from feature_level_style_transfer_for_tsc_b200 import multi_source_voting as MV
    g = torch.Generator().manual_seed(9)
    # one series, one class, one model
    lg = torch.zeros(1, 1)
    pred, counts, prec = T.ops.class_precision(lg.cuda(), torch.zeros(1, dtype=torch.int64).cuda())
    assert pred.tolist() == [0] and counts.tolist() == [[1], [1]] and prec.tolist() == [1.0]
    score, p = MV.entropy_vote([lg.cuda()], [prec])
    ref_s, ref_p = V.entropy_vote([lg.numpy()], V.normalized_weights([prec.cpu().numpy()]))
    assert rel_err(score.cpu(), ref_s) < 1e-6 and p.tolist() == ref_p.tolist()
    # the largest shape: 64 classes, 8 models, a ragged number of rows; uniform rows (maximum entropy) and a one-hot row
    N, K, M = 1000 + 37, 64, 8
    logits = [torch.randn(N, K, generator=g) * 3 for _ in range(M)]
    logits[0][5] = 0.0
    logits[1][6] = -40.0
    logits[1][6, 17] = 40.0
    labels = torch.randint(0, K, (N,), generator=g)
    precs = [T.ops.class_precision(l.cuda(), labels.cuda())[2] for l in logits]
    for l, pr in zip(logits, precs):
        assert np.array_equal(pr.cpu().numpy(), V.class_precision(l.numpy(), labels.numpy(), K))
    score, p = MV.entropy_vote([l.cuda() for l in logits], precs)
    ref_s, ref_p = V.entropy_vote([l.numpy() for l in logits], V.normalized_weights([pr.cpu().numpy() for pr in precs]))
    assert rel_err(score.cpu(), ref_s) < 1e-5
    assert np.array_equal(p.cpu().numpy(), ref_p)
    # exp overflow: inf / inf = NaN in the reference's softmax; numpy.argmax picks the first NaN -- so does the kernel
    over = torch.tensor([[1.0, 100.0, 2.0], [0.5, 0.1, 0.2]])
    one = torch.ones(3, dtype=torch.float64).cuda()
    score, p = MV.entropy_vote([over.cuda()], [one])
    with np.errstate(all="ignore"):
        ref_s, ref_p = V.entropy_vote([over.numpy()], np.ones((1, 3)))
    assert np.isnan(ref_s[0]).any() and p.tolist() == ref_p.tolist()
    assert np.array_equal(np.isnan(score.cpu().numpy()), np.isnan(ref_s))
    assert T.ops.class_precision(torch.tensor([[0.0, float("nan"), 5.0]]).cuda())[0].tolist() == [1]
    with pytest.raises(RuntimeError):
        MV.entropy_vote([lg.cuda()] * 9, [prec] * 9)                 # more models than TSC_MAX_VOTERS
    # 32 tensors in one norm call, one of them empty
    ts = [torch.randn(17 * (i + 1), device="cuda") for i in range(31)] + [torch.empty(0, device="cuda")]
    out = T.ops.multi_l2norm(ts).cpu().numpy()
    ref = np.array([float(torch.linalg.vector_norm(t.double())) for t in ts])
    assert rel_err(out[:-1], ref) < 2e-6 and out[31] == 0.0
    with pytest.raises(RuntimeError):
        T.ops.multi_l2norm(ts + ts[:1])
    # a single series through the inference path
    from feature_level_style_transfer_for_tsc_b200.OS_CNN.OS_CNN import OS_CNN, OS_CNN_res
    T.set_engine("tcgen05")
    lpl_e, lpl_c = O.trainer_layer_lists(9, 128)
    torch.manual_seed(4)
    fe, cl = OS_CNN_res(lpl_e).cuda().eval(), OS_CNN(lpl_c, 6).cuda().eval()
    x, _ = O.synthetic_batch(3, 9, 128, 6, 0)
    with torch.no_grad():
        all3 = cl(fe(x.cuda()))[0]
        one_ = cl(fe(x[1:2].cuda()))[0]
    assert torch.equal(all3[1:2], one_)                              # eval mode: no coupling across the batch, bit for bit
