"""C-DAN consumer (BASELINE configuration 3) on the GPU: the fused kernels against the reference's own values
(tests/golden/cdan_small.npz) and the oracle; the cfg3 pair / multi-source step against the oracle step."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import cdan as OC
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


def _mirror_from_golden(z, hidden):
    from feature_level_style_transfer_for_tsc_b200.C_DAN import RandomLayer
    from feature_level_style_transfer_for_tsc_b200.widgets import AdversarialNetworkforCDAN
    rl = RandomLayer([z["R0"].shape[0], z["R1"].shape[0]], with_nvidia=False)
    rl.random_matrix = [torch.from_numpy(z["R0"]).cuda(), torch.from_numpy(z["R1"]).cuda()]
    ad = AdversarialNetworkforCDAN(1024, hidden)
    ad.load_state_dict({k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("ad/")})
    ad.dropout1.p = 0.0
    ad.dropout2.p = 0.0
    return rl, ad.cuda()


def test_cdan_matches_the_reference_vectors(T, tables, cdan_small):
    """Three consecutive calls (two training-mode, one eval-mode: the reversal schedule moves 0 -> 0.9866 -> ~1) on the
    reference's seeded inputs: loss, input gradients, critic gradients."""
    from feature_level_style_transfer_for_tsc_b200 import C_DAN
    z = cdan_small
    t = tables["cdan_small"]
    rl, ad = _mirror_from_golden(z, t["hidden"])
    _run_calls(C_DAN, z, t, rl, ad)


def test_cdan_rejects_every_other_call_shape(T):
    """No alternate route: the outer-product form (random_layer=None) and other view counts raise."""
    from feature_level_style_transfer_for_tsc_b200 import C_DAN
    from feature_level_style_transfer_for_tsc_b200.widgets import AdversarialNetworkforCDAN
    ad = AdversarialNetworkforCDAN(1024, 8).cuda()
    f, l = torch.zeros(2, 3, 4, device="cuda"), torch.zeros(2, 3, device="cuda")
    with pytest.raises(RuntimeError):
        C_DAN.CDAN(f, f, l, l, ad, None)
    with pytest.raises(RuntimeError):
        C_DAN.CDAN(f, f, l, l, ad, C_DAN.RandomLayer([12], with_nvidia=False).cuda())


def _run_calls(C_DAN, z, t, rl, ad):
    for call, training in enumerate((True, True, False)):
        ad.train(training)
        ins = [torch.from_numpy(z[f"c{call}/{n}"]).cuda().requires_grad_(True) for n in ("ft", "fs", "lt", "ls")]
        ad.zero_grad()
        loss = C_DAN.CDAN(*ins, ad, rl)
        loss.backward()
        torch.cuda.synchronize()
        assert abs(float(loss) - float(z[f"c{call}/loss"])) <= 1e-5 * max(1.0, abs(float(z[f"c{call}/loss"]))), call
        assert abs(ad.coeff - t["coeff_after_call"][call]) < 1e-12
        for x, n in zip(ins, ("dft", "dfs", "dlt", "dls")):
            assert rel_err(x.grad.cpu(), z[f"c{call}/{n}"]) < 5e-5, (call, n)
        for k, p in ad.named_parameters():
            if k == "ad_layer3.bias":
                # d loss / d(last bias) = B * (sum w_target - sum w_generated), both sums 1 up to rounding: noise
                assert float(p.grad.abs().max()) < 5e-6
                continue
            assert rel_err(p.grad.cpu(), z[f"c{call}/dad/{k}"]) < 5e-5, (call, k)
    assert ad.iter_num == t["iter_num_after"]


def test_cdan_kernels_against_the_oracle_fp64(T):
    """Op level at the cfg2 class count and the full random width, B not a multiple of anything: fused forward /
    backward kernels against oracle/cdan.py in fp64 on the same inputs."""
    ops = T.ops
    B, K, D = 37, 6, 1024
    g = torch.Generator().manual_seed(2)
    y0 = torch.randn(2 * B, D, generator=g) * 30.0
    logits = torch.randn(2 * B, K, generator=g) * 2.0
    r1 = torch.randn(K, D, generator=g)
    crit_w = torch.randn(D, 1, generator=g) * 0.05
    coeff = torch.tensor([0.25, 0.75, 0.5])
    # oracle composition in fp64
    y0d, lgd, r1d = (t.double().requires_grad_(r) for t, r in ((y0, True), (logits, True), (r1, False)))
    p = torch.softmax(lgd, 1)
    fusion_d = (y0d / 32.0) * (p @ r1d)
    fin = OC._Reverse.apply(fusion_d[:B], 0.25), OC._Reverse.apply(fusion_d[B:], 0.75)
    out_d = torch.cat(fin) @ crit_w.double()
    h = OC._Reverse.apply(OC.entropy(p), 0.5)
    u_d = 1.0 + torch.exp(-h)
    wt, ws = u_d[:B] / u_d[:B].sum().detach(), u_d[B:] / u_d[B:].sum().detach()
    loss_d = torch.sum(wt * out_d[:B]) - torch.sum(ws * out_d[B:])
    loss_d.backward()
    # kernels
    fusion, prob, u = ops.cdan_fuse_fwd(y0.cuda(), logits.cuda(), r1.cuda(), 32.0)
    assert rel_err(fusion.cpu(), fusion_d.detach()) < 1e-5
    assert rel_err(prob.cpu(), p.detach()) < 1e-6 and rel_err(u.cpu(), u_d.detach()) < 1e-6
    out = fusion @ crit_w.cuda()
    loss, saved = ops.cdan_distance_fwd(u, out)
    assert abs(float(loss) - float(loss_d)) < 1e-5 * max(1.0, abs(float(loss_d)))
    du, dcrit = ops.cdan_distance_bwd(torch.ones((), device="cuda"), saved, B)
    dfusion = dcrit @ crit_w.cuda().t()
    dy0, dlogits = ops.cdan_fuse_bwd(dfusion.contiguous(), y0.cuda(), prob, r1.cuda(), u, du, coeff.cuda(), 32.0)
    torch.cuda.synchronize()
    assert rel_err(dy0.cpu(), y0d.grad) < 1e-5
    assert rel_err(dlogits.cpu(), lgd.grad) < 2e-5


@pytest.mark.parametrize("use_graph", [False, True])
def test_pair_steps_match_the_oracle_fp32_engine(T, use_graph):
    """cfg3 pair (eval-BatchNorm classifier on the generated features + C-DAN + WGAN clipping of the critic): losses of
    three consecutive steps against the oracle (fp32 engine; dropout off, its mask is generator-specific)."""
    from feature_level_style_transfer_for_tsc_b200.train_step import TransferPairModelSet, Trainer
    T.set_engine("simt")
    Ct, Lt, Kt, Cs, Ls, Ks, B = 3, 96, 3, 2, 80, 4, 6
    torch.manual_seed(0)
    model = TransferPairModelSet(Ct, Lt, Kt, Cs, Ls, Ks, critic_hidden=64).cuda()
    model.ad_net.dropout1.p = model.ad_net.dropout2.p = 0.0
    tr = Trainer(model, style_weight=50.0, use_graph=use_graph)
    oms = OS.PairModelSet(Ct, Lt, Kt, Cs, Ls, Ks, seed=0, critic_hidden=64)
    assert torch.equal(model.random_layer.random_matrix[0].cpu(), oms.mats[0])
    assert torch.equal(model.ad_net.ad_layer2.weight.detach().cpu(), oms.ad_net["ad_layer2.weight"])
    xt, yt = O.synthetic_batch(B, Ct, Lt, Kt, 0)
    xs, ys = O.synthetic_batch(B, Cs, Ls, Ks, 1)
    for step in range(3):
        loss = float(tr.step(xt.cuda(), yt.cuda(), xs.cuda(), ys.cuda()))
        oloss = OS.pair_train_step(oms, xt, yt, xs, ys, 50.0)
        # step 0 is pure forward parity; later steps see RMSprop's sign-like, ill-conditioned first updates
        assert abs(loss - oloss) < (2e-4 if step == 0 else 3e-2) * max(1.0, abs(oloss)), (step, loss, oloss)
        assert model.ad_net.iter_num == oms.ad_state.iter_num == 2 * step + 1
    torch.cuda.synchronize()
    got = model.ad_net.state_dict()
    for k, v in oms.ad_net.items():
        assert float(got[k].abs().max()) <= 0.0005 + 1e-9                 # clipped like the reference's critic
    assert rel_err(model.cl_t.state_dict()["net.0.bn.running_mean"].cpu(), oms.cl_t["net.0.bn.running_mean"]) < 5e-3
    assert int(model.cl_t.state_dict()["net.0.bn.num_batches_tracked"]) == 3      # the eval-mode call does not count
    T.set_engine("tcgen05")


def test_multi_source_step_tensor_core_engine(T):
    """Three sources of different (C, L, classes) on one target batch, tcgen05 engine, one CUDA graph: the first loss
    tracks the oracle within the bf16 tolerance, later losses stay finite and fall."""
    from feature_level_style_transfer_for_tsc_b200.train_step import MultiSourceModelSet, Trainer
    T.set_engine("tcgen05")
    target, sources, B = (9, 128, 6), [(1, 128, 5), (3, 256, 4), (9, 128, 6)], 16
    torch.manual_seed(0)
    model = MultiSourceModelSet(target, sources, critic_hidden=128).cuda()
    for pair in model.pairs:
        pair.ad_net.dropout1.p = pair.ad_net.dropout2.p = 0.0
    tr = Trainer(model, use_graph=True)
    osets = OS.multi_source_models(target, sources, seed=0, critic_hidden=128)
    for pair, oms in zip(model.pairs, osets):
        assert torch.equal(pair.fe_s.state_dict()["net_1.res.conv1d.weight"].cpu(), oms.fe_s["net_1.res.conv1d.weight"])
        assert torch.equal(pair.random_layer.random_matrix[1].cpu(), oms.mats[1])
    xt, yt = O.synthetic_batch(B, *target[:2], target[2], 0)
    batches = [O.synthetic_batch(B, C, L, K, 1 + i) for i, (C, L, K) in enumerate(sources)]
    flat = [t.cuda() for xb in batches for t in xb]
    # the pair losses have both signs (C-DAN is a difference of two critic means, weighted 3): the bf16 tolerance is
    # relative to the magnitude of the terms, not to their partly cancelling sum
    probe = OS.multi_source_models(target, sources, seed=0, critic_hidden=128)
    scale = 0.0
    for ms, (xs, ys) in zip(probe, batches):
        ms.set_requires_grad()
        o = OS.pair_step_forward(ms, xt, yt, xs, ys, 1.0)
        scale += float(o["ce_t"].detach().abs() + o["ce_s"].detach().abs() + ms.CDAN_WEIGHT * o["cdan"].detach().abs())
    oloss = OS.multi_source_train_step(osets, xt, yt, batches, 1.0)
    losses = [float(tr.step(xt.cuda(), yt.cuda(), *flat)) for _ in range(6)]
    torch.cuda.synchronize()
    assert T.ops.read_watchdog() == 0
    assert abs(losses[0] - oloss) < 1e-2 * scale, (losses[0], oloss, scale)
    assert all(np.isfinite(losses))
    assert [p.ad_net.iter_num for p in model.pairs] == [11, 11, 11]
