"""Op-level parity of every C-ABI kernel against the oracle, on identical inputs (test kernels
op by op; tolerances: fp32 SIMT engine <= 1e-5 scale-relative, tcgen05/bf16 engine <= 1e-2).
All tests call through the C-ABI (ctypes -> libtsc_b200.so)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err
from oracle import os_cnn as O
from oracle import style as S

pytestmark = pytest.mark.gpu

TOL_F32 = 1e-5
TOL_BF16 = 1e-2


@pytest.fixture(scope="module")
def T():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import feature_level_style_transfer_for_tsc_b200 as pkg
    pkg._lib.load()          # fails loudly when the extension is missing
    return pkg


def c8_ref(x, dtype=torch.float32):
    """independent torch statement of the c8 layout: [B,C,L] -> [B,Cp/8,L,8]"""
    B, C, L = x.shape
    Cp = (C + 15) // 16 * 16
    xp = F.pad(x, (0, 0, 0, Cp - C))
    return xp.reshape(B, Cp // 8, 8, L).permute(0, 1, 3, 2).contiguous().to(dtype)


def c8_to_ncl_ref(x8, C):
    B, cpc, L, _ = x8.shape
    return x8.float().permute(0, 1, 3, 2).reshape(B, cpc * 8, L)[:, :C].contiguous()


BANKS = {
    "small7": [(3, 4, 1), (3, 4, 2), (3, 4, 3), (3, 4, 5), (3, 4, 7)],
    "mid13": [(20, 5, k) for k in (1, 2, 3, 5, 7, 11, 13)],
    "even2": [(30, 20, 1), (30, 20, 2)],
    "cfg2_l1": O.trainer_layer_lists(9, 128)[0][1],       # 72 -> 228, Kmax 31
    "cfg2_l0": O.trainer_layer_lists(9, 128)[0][0],       # 9 -> 72
    "cfg2_l2": O.trainer_layer_lists(9, 128)[0][2],       # 228 -> 144, Kmax 2
    "cfg4_l1": O.trainer_layer_lists(3, 1024)[0][1],      # 25 -> 225, Kmax 89
    "one": [(5, 7, 1)],
}
SHAPES = {"small7": (3, 50), "mid13": (2, 160), "even2": (2, 96), "cfg2_l1": (3, 128), "cfg2_l0": (2, 128),
          "cfg2_l2": (2, 128), "cfg4_l1": (1, 300), "one": (2, 33)}


def make_case(name, seed=0):
    layer = BANKS[name]
    g = O.bank_geometry(layer)
    B, L = SHAPES[name]
    gen = torch.Generator().manual_seed(seed)
    x = torch.randn(B, g["cin"], L, generator=gen)
    w = torch.randn(g["cout"], g["cin"], g["kmax"], generator=gen) / np.sqrt(g["cin"] * 3.0)
    b = torch.randn(g["cout"], generator=gen)
    dy = torch.randn(B, g["cout"], L, generator=gen)
    return layer, g, x, w, b, dy


def test_layout_roundtrip(T):
    ops, L = T.ops, T._lib
    x = torch.randn(3, 21, 77)
    xd = x.cuda()
    for dt, td in ((L.TSC_F32, torch.float32), (L.TSC_BF16, torch.bfloat16)):
        x8 = ops.ncl_to_c8(xd, dt)
        assert x8.shape == (3, 4, 77, 8) and x8.dtype == td
        assert torch.equal(x8.cpu(), c8_ref(x, td))
    back = ops.c8_to_ncl(ops.ncl_to_c8(xd, L.TSC_F32), 21)
    assert torch.equal(back.cpu(), x)


@pytest.mark.parametrize("name", list(BANKS))
def test_conv_fwd_dgrad_wgrad_simt_fp32(T, name):
    ops, L = T.ops, T._lib
    layer, g, x, w, b, dy = make_case(name)
    geom = ops.bank_geometry(layer)
    assert geom.s_of_tap == g["s_of_tap"]
    wd = w.clone().cuda()
    wf = ops.pack_weights(geom, wd, L.DIR_FWD, L.TSC_F32, True)
    mask = torch.from_numpy(O.build_mask(layer))
    assert torch.equal(wd.cpu(), w * mask)            # weight.data = weight * mask (OS_CNN.py:68)
    y8 = ops.osconv(L.ENGINE_SIMT, L.DIR_FWD, geom, ops.ncl_to_c8(x.cuda(), L.TSC_F32), wf, b.cuda())
    y = c8_to_ncl_ref(y8.cpu(), g["cout"])
    y_ref = O.masked_conv(x.double(), w.double(), b.double(), layer)
    assert rel_err(y, y_ref) < TOL_F32
    full = y8.cpu().permute(0, 1, 3, 2).reshape(x.shape[0], -1, x.shape[2])
    if geom.cout_p > g["cout"]:
        assert float(full[:, g["cout"]:].abs().max()) == 0.0          # pad channels stay zero
    # dgrad
    wdg = ops.pack_weights(geom, wd, L.DIR_DGRAD, L.TSC_F32, False)
    dx8 = ops.osconv(L.ENGINE_SIMT, L.DIR_DGRAD, geom, ops.ncl_to_c8(dy.cuda(), L.TSC_F32), wdg, None)
    dx = ops.c8_to_ncl(dx8, g["cin"]).cpu()
    assert rel_err(dx, O.conv_dgrad(dy.double(), w.double(), layer)) < TOL_F32
    # wgrad
    dw = ops.oswgrad(L.ENGINE_SIMT, geom, ops.ncl_to_c8(dy.cuda(), L.TSC_F32), ops.ncl_to_c8(x.cuda(), L.TSC_F32)).cpu()
    dw_ref = O.conv_wgrad(dy.double(), x.double(), layer)
    assert rel_err(dw, dw_ref) < TOL_F32
    assert float((dw * (1 - mask)).abs().max()) == 0.0      # masked taps: exact zeros


@pytest.mark.parametrize("name", list(BANKS))
def test_conv_tcgen05_matches_simt_on_bf16_operands(T, name):
    """The tensor-core engine against (a) the SIMT engine fed the same bf16 operands (same products, fp32
    accumulation: only the summation order differs) and (b) the fp64 oracle at the bf16 tolerance."""
    ops, L = T.ops, T._lib
    if not L.load().tsc_device_supports_tcgen05():
        pytest.skip("device has no tcgen05")
    layer, g, x, w, b, dy = make_case(name, seed=1)
    geom = ops.bank_geometry(layer)
    wd = w.clone().cuda()
    for direction, inp, bias in ((L.DIR_FWD, x, b.cuda()), (L.DIR_DGRAD, dy, None)):
        wp = ops.pack_weights(geom, wd, direction, L.TSC_BF16, direction == L.DIR_FWD)
        in8 = ops.ncl_to_c8(inp.cuda(), L.TSC_BF16)
        y_tc = ops.osconv(L.ENGINE_TCGEN05, direction, geom, in8, wp, bias)
        torch.cuda.synchronize()
        assert ops.read_watchdog() == 0, "tcgen05 pipeline timed out"
        y_simt = ops.osconv(L.ENGINE_SIMT, direction, geom, in8, wp, bias)
        assert rel_err(y_tc.cpu(), y_simt.cpu()) < 2e-5, f"direction {direction}"
        cout = g["cout"] if direction == L.DIR_FWD else g["cin"]
        ref = (O.masked_conv(x.double(), w.double(), b.double(), layer) if direction == L.DIR_FWD
               else O.conv_dgrad(dy.double(), w.double(), layer))
        assert rel_err(c8_to_ncl_ref(y_tc.cpu(), cout), ref) < TOL_BF16


@pytest.mark.parametrize("name", ["small7", "mid13", "cfg2_l1", "cfg2_l2", "cfg4_l1"])
def test_conv_tcgen05_fused_epilogues(T, name):
    """Forward: per-CTA (mean, M2) partials of y merge to the batch statistics.  Dgrad: the written gradient is
    already masked by the ReLU of the layer below and the per-CTA (S1, S2) partials sum to the BatchNorm-backward
    reductions, both against fp64 torch on the kernel's own fp32 output."""
    ops, L = T.ops, T._lib
    if not L.load().tsc_device_supports_tcgen05():
        pytest.skip("device has no tcgen05")
    layer, g, x, w, b, dy = make_case(name, seed=3)
    geom = ops.bank_geometry(layer)
    B, Ln = x.shape[0], x.shape[2]
    ncta, ltiles = ops.n_conv_ctas(B, Ln), (Ln + 127) // 128
    wd = w.clone().cuda()
    # forward statistics
    wp = ops.pack_weights(geom, wd, L.DIR_FWD, L.TSC_BF16, True)
    part = torch.full((ncta, geom.cout_p, 2), float("nan"), device="cuda")
    y8 = ops.osconv(L.ENGINE_TCGEN05, L.DIR_FWD, geom, ops.ncl_to_c8(x.cuda(), L.TSC_BF16), wp, b.cuda(), stat_partial=part)
    torch.cuda.synchronize()
    assert ops.read_watchdog() == 0
    y = c8_to_ncl_ref(y8.cpu(), geom.cout).double()                       # [B, C, L]
    part = part.cpu().double()
    n_rows = torch.tensor([min(128, Ln - (i % ltiles) * 128) for i in range(ncta)], dtype=torch.float64)
    mean_i, m2_i = part[:, :geom.cout, 0], part[:, :geom.cout, 1]
    N = float(B * Ln)
    mean = (mean_i * n_rows[:, None]).sum(0) / N
    m2 = m2_i.sum(0) + (n_rows[:, None] * (mean_i - mean) ** 2).sum(0)
    assert rel_err(mean, y.mean(dim=(0, 2))) < 1e-5
    assert rel_err(m2 / N, y.var(dim=(0, 2), unbiased=False)) < 1e-4
    # dgrad with the mask / reductions of a layer below that has geom.cin channels
    gen = torch.Generator().manual_seed(11)
    cin, cp = geom.cin, geom.cin_p
    y_below = torch.randn(B, cin, Ln, generator=gen)
    scale, shift = torch.randn(cin, generator=gen), torch.randn(cin, generator=gen) * 0.3
    mean_b, invstd_b = torch.randn(cin, generator=gen) * 0.1, torch.rand(cin, generator=gen) + 0.5
    pad = lambda v: F.pad(v, (0, cp - cin)).cuda()
    wpd = ops.pack_weights(geom, wd, L.DIR_DGRAD, L.TSC_BF16, False)
    dy8 = ops.ncl_to_c8(dy.cuda(), L.TSC_BF16)
    plain = ops.osconv(L.ENGINE_TCGEN05, L.DIR_DGRAD, geom, dy8, wpd, None)
    for relu in (True, False):
        red = torch.full((ncta, cp, 2), float("nan"), device="cuda")
        mask = (ops.ncl_to_c8(y_below.cuda(), L.TSC_F32), pad(scale) if relu else None, pad(shift) if relu else None,
                pad(mean_b), pad(invstd_b))
        d8 = ops.osconv(L.ENGINE_TCGEN05, L.DIR_DGRAD, geom, dy8, wpd, None, mask=mask, red_partial=red)
        torch.cuda.synchronize()
        assert ops.read_watchdog() == 0
        dz = c8_to_ncl_ref(plain.cpu(), cin).double()
        z = scale[None, :, None].double() * y_below.double() + shift[None, :, None].double()
        d_ref = dz * (z > 0) if relu else dz
        assert rel_err(c8_to_ncl_ref(d8.cpu(), cin), d_ref) < 1e-6
        yhat = (y_below.double() - mean_b[None, :, None].double()) * invstd_b[None, :, None].double()
        red = red.cpu().double().sum(0)
        assert rel_err(red[:cin, 0], d_ref.sum(dim=(0, 2))) < 1e-4
        assert rel_err(red[:cin, 1], (d_ref * yhat).sum(dim=(0, 2))) < 1e-4


@pytest.mark.parametrize("name", list(BANKS))
def test_wgrad_tcgen05(T, name):
    ops, L = T.ops, T._lib
    if not L.load().tsc_device_supports_tcgen05():
        pytest.skip("device has no tcgen05")
    layer, g, x, w, b, dy = make_case(name, seed=2)
    geom = ops.bank_geometry(layer)
    dy8, x8 = ops.ncl_to_c8(dy.cuda(), L.TSC_BF16), ops.ncl_to_c8(x.cuda(), L.TSC_BF16)
    dw_tc = ops.oswgrad(L.ENGINE_TCGEN05, geom, dy8, x8)
    torch.cuda.synchronize()
    assert ops.read_watchdog() == 0, "tcgen05 pipeline timed out"
    dw_simt = ops.oswgrad(L.ENGINE_SIMT, geom, dy8, x8)
    assert rel_err(dw_tc.cpu(), dw_simt.cpu()) < 5e-5
    mask = torch.from_numpy(O.build_mask(layer))
    assert float((dw_tc.cpu() * (1 - mask)).abs().max()) == 0.0
    assert rel_err(dw_tc.cpu(), O.conv_wgrad(dy.double(), x.double(), layer)) < TOL_BF16


@pytest.mark.parametrize("B,C,L", [(4, 20, 50), (16, 72, 128), (3, 228, 128), (2, 7, 1000)])
def test_batchnorm_fwd_bwd(T, B, C, L):
    ops, Lb = T.ops, T._lib
    gen = torch.Generator().manual_seed(3)
    y = torch.randn(B, C, L, generator=gen) * 2 + 0.5
    y2 = torch.randn(B, C, L, generator=gen)
    gamma = torch.rand(C, generator=gen) + 0.5
    beta = torch.randn(C, generator=gen)
    dz = torch.randn(B, C, L, generator=gen)
    rm, rv = torch.randn(C, generator=gen) * 0.1, torch.rand(C, generator=gen) + 0.5
    y8 = ops.ncl_to_c8(y.cuda(), Lb.TSC_F32)
    # --- training statistics
    rm_d, rv_d = rm.clone().cuda(), rv.clone().cuda()
    co = ops.bn_stats(y8, C, gamma.cuda(), beta.cuda(), rm_d, rv_d, 0.1, 1e-5)
    yd = y.double()
    mean, var = yd.mean((0, 2)), yd.var((0, 2), unbiased=False)
    assert rel_err(co.mean.cpu()[:C], mean) < 1e-6
    assert rel_err(co.invstd.cpu()[:C], 1 / torch.sqrt(var + 1e-5)) < 1e-6
    n = B * L
    assert rel_err(rm_d.cpu(), 0.9 * rm + 0.1 * mean) < 1e-6
    assert rel_err(rv_d.cpu(), 0.9 * rv + 0.1 * var * n / (n - 1)) < 1e-6
    # --- apply (+relu), all output kinds
    sd = {"weight": gamma.double(), "bias": beta.double(), "running_mean": rm.double(), "running_var": rv.double(),
          "num_batches_tracked": torch.tensor(0)}
    z_ref = torch.relu(O.batch_norm(yd, sd, "", True, update_running=False))
    z_ncl = ops.bn_apply(y8, co, C, True, Lb.OUT_NCL_F32).cpu()
    assert rel_err(z_ncl, z_ref) < TOL_F32
    z8 = ops.bn_apply(y8, co, C, True, Lb.OUT_C8_F32).cpu()
    assert torch.equal(c8_to_ncl_ref(z8, C), z_ncl)
    zb = ops.bn_apply(y8, co, C, True, Lb.OUT_C8_BF16).cpu()
    assert torch.equal(zb, c8_ref(z_ncl, torch.bfloat16))
    # --- backward, train mode with relu
    dz8 = ops.ncl_to_c8(dz.cuda(), Lb.TSC_F32)
    s1, s2 = ops.bn_bwd_reduce(dz8, y8, co, C, (y8, co))
    dy8 = ops.bn_bwd_apply(dz8, y8, co, gamma.cuda(), s1, s2, True, C, Lb.TSC_F32, (y8, co))
    dy_ref, dg_ref, db_ref = O.bn_relu_backward(dz.double(), yd, gamma.double(), mean, var, (z_ref > 0).double(), True)
    assert rel_err(ops.c8_to_ncl(dy8, C).cpu(), dy_ref) < 2e-5
    assert rel_err(s2.cpu()[:C], dg_ref) < 2e-5 and rel_err(s1.cpu()[:C], db_ref) < 2e-5
    # --- eval mode (running stats, gradient still flows: train_and_test.py:583-586)
    coe = ops.bn_eval_coeffs(C, gamma.cuda(), beta.cuda(), rm.cuda(), rv.cuda(), 1e-5)
    ze_ref = torch.relu(O.batch_norm(yd, sd, "", False))
    assert rel_err(ops.bn_apply(y8, coe, C, True, Lb.OUT_NCL_F32).cpu(), ze_ref) < TOL_F32
    s1, s2 = ops.bn_bwd_reduce(dz8, y8, coe, C, (y8, coe))
    dy8 = ops.bn_bwd_apply(dz8, y8, coe, gamma.cuda(), s1, s2, False, C, Lb.TSC_F32, (y8, coe))
    dy_ref, dg_ref, db_ref = O.bn_relu_backward(dz.double(), yd, gamma.double(), rm.double(), rv.double(),
                                                (ze_ref > 0).double(), False)
    assert rel_err(ops.c8_to_ncl(dy8, C).cpu(), dy_ref) < 2e-5
    assert rel_err(s2.cpu()[:C], dg_ref) < 2e-5 and rel_err(s1.cpu()[:C], db_ref) < 2e-5
    # --- two-branch apply (shortcut add + relu) and its shared mask
    y28 = ops.ncl_to_c8(y2.cuda(), Lb.TSC_F32)
    co2 = ops.bn_stats(y28, C, gamma.cuda(), beta.cuda(), None, None, 0.0, 1e-5)
    a = O.batch_norm(yd, sd, "", True, update_running=False)
    bb = O.batch_norm(y2.double(), sd, "", True, update_running=False)
    out = ops.bn_apply(y8, co, C, True, Lb.OUT_NCL_F32, y2=y28, co2=co2).cpu()
    assert rel_err(out, torch.relu(a + bb)) < TOL_F32
    s1, s2 = ops.bn_bwd_reduce(dz8, y8, co, C, (y8, co), (y28, co2))
    dy8 = ops.bn_bwd_apply(dz8, y8, co, gamma.cuda(), s1, s2, True, C, Lb.TSC_F32, (y8, co), (y28, co2))
    dy_ref, _, _ = O.bn_relu_backward(dz.double(), yd, gamma.double(), mean, var, ((a + bb) > 0).double(), True)
    # elements whose pre-activation is within rounding of 0 may flip the mask (F7): compare away from them
    safe = ((a + bb).abs() > 1e-5).double()
    assert rel_err(ops.c8_to_ncl(dy8, C).cpu() * safe, dy_ref * safe) < 1e-4


@pytest.mark.parametrize("B,C,L,relu,two", [(5, 20, 100, True, False), (5, 20, 128, True, True), (4, 7, 100, False, False),
                                            (6, 40, 300, True, True), (16, 228, 128, True, False)])
def test_fused_batchnorm_matches_the_unfused_kernels(T, B, C, L, relu, two):
    """The fused BatchNorm path (statistics merged from 128-row partials in the apply prologue; backward sums from
    partials) against the stand-alone kernels that are themselves checked against the oracle in
    test_batchnorm_fwd_bwd: same formulas, so fp32 round-off only."""
    ops, Lb = T.ops, T._lib
    gen = torch.Generator().manual_seed(13)
    Cp = ops.pad16(C)
    ncta, ltiles = ops.n_conv_ctas(B, L), (L + 127) // 128

    def partials(y):            # what the conv epilogue writes: (mean, M2) per 128-row tile, computed here in fp64
        part = torch.zeros(ncta, Cp, 2, dtype=torch.float64)
        for i in range(ncta):
            b, l0 = i // ltiles, (i % ltiles) * 128
            seg = y[b, :, l0:l0 + 128].double()
            part[i, :C, 0] = seg.mean(-1)
            part[i, :C, 1] = ((seg - seg.mean(-1, keepdim=True)) ** 2).sum(-1)
        return part.float().cuda()

    def make():
        y = torch.randn(B, C, L, generator=gen) * 2 + 0.5
        gamma, beta = torch.rand(C, generator=gen) + 0.5, torch.randn(C, generator=gen) * 0.2
        return y, gamma.cuda(), beta.cuda()

    branches = []
    for _ in range(2 if two else 1):
        y, gamma, beta = make()
        y8 = ops.ncl_to_c8(y.cuda(), Lb.TSC_F32)
        rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
        rm2, rv2 = rm.clone(), rv.clone()
        co = ops.bn_stats(y8, C, gamma, beta, rm, rv, 0.1, 1e-5)
        coef = torch.empty(4, Cp, device="cuda")
        fw = ops.BNLayerFwd(y8, partials(y), gamma, beta, rm2, rv2, 0.1, 1e-5, coef)
        branches.append((y8, gamma, co, fw, rm, rv, rm2, rv2))
    (y8, gamma, co, fw, rm, rv, rm2, rv2) = branches[0]
    second = branches[1] if two else None
    for kind in (Lb.OUT_C8_BF16, Lb.OUT_NCL_F32):
        ref = ops.bn_apply(y8, co, C, relu, kind, y2=second[0] if two else None, co2=second[2] if two else None)
        for br in branches:                       # running statistics advance once per call
            br[6].copy_(torch.zeros_like(br[6])); br[7].copy_(torch.ones_like(br[7]))
        got = ops.bn_apply_fused(fw, second[3] if two else None, C, relu, kind)
        tol = 1e-2 if kind == Lb.OUT_C8_BF16 else 2e-6
        assert rel_err(got.float().cpu(), ref.float().cpu()) < tol
    assert rel_err(fw.coef.cpu()[0, :C], co.mean.cpu()[:C]) < 1e-6 and rel_err(fw.coef.cpu()[1, :C], co.invstd.cpu()[:C]) < 1e-5
    assert rel_err(rm2.cpu(), rm.cpu()) < 1e-6 and rel_err(rv2.cpu(), rv.cpu()) < 1e-5
    # eval mode: coefficients from the running statistics
    ce = ops.bn_eval_coeffs(C, gamma, branches[0][3].beta, rm, rv, 1e-5)
    fe_ = ops.BNLayerFwd(y8, None, gamma, branches[0][3].beta, rm, rv, 0.0, 1e-5, torch.empty(4, Cp, device="cuda"))
    assert rel_err(ops.bn_apply_fused(fe_, None, C, relu, Lb.OUT_NCL_F32).cpu(), ops.bn_apply(y8, ce, C, relu, Lb.OUT_NCL_F32).cpu()) < 2e-6
    # backward
    dout = torch.randn(B, C, L, generator=gen).cuda()
    dz8 = ops.ncl_to_c8(dout, Lb.TSC_F32)
    m1 = (y8, co) if relu else None
    m2 = (second[0], second[2]) if (two and relu) else None
    S = ops.bn_fused_splits(B, C, L)
    bws = []
    for (yy, gg, cc, ff, *_r) in branches:
        coef = torch.stack([cc.mean, cc.invstd, cc.scale, cc.shift]).contiguous()
        bws.append(ops.BNLayerBwd(yy, coef, gg, True, torch.empty(S, Cp, 2, device="cuda")))
    d8 = ops.bn_bwd_top(dout, bws[0], bws[1] if two else None, relu)
    for (yy, gg, cc, ff, *_r), bw in zip(branches, bws):
        s1, s2 = ops.bn_bwd_reduce(dz8, yy, cc, C, m1, m2)
        dy_ref = ops.bn_bwd_apply(dz8, yy, cc, gg, s1, s2, True, C, Lb.TSC_F32, m1, m2)
        bw.dgamma, bw.dbeta = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
        dy = ops.bn_bwd_apply_fused(d8, bw, S, C, Lb.TSC_F32, False)
        assert rel_err(dy.cpu(), dy_ref.cpu()) < 2e-6
        assert rel_err(bw.dgamma.cpu(), s2.cpu()[:C]) < 1e-6 and rel_err(bw.dbeta.cpu(), s1.cpu()[:C]) < 1e-6
        dg0 = bw.dgamma.clone()
        ops.bn_bwd_apply_fused(d8, bw, S, C, Lb.TSC_F32, True)          # accumulate
        assert rel_err(bw.dgamma.cpu(), 2 * dg0.cpu()) < 1e-6


@pytest.mark.parametrize("B,C,L", [(4, 6, 32), (8, 144, 128), (3, 5, 100), (2, 50, 1024), (2, 3, 4096), (2, 3, 777)])
def test_rowstats_and_adain(T, B, C, L):
    ops = T.ops
    gen = torch.Generator().manual_seed(4)
    c = torch.randn(B, C, L, generator=gen) * 3 + 10          # large mean: exercises Welford's stability
    s = torch.randn(B, C, L, generator=gen) * 0.5 - 2
    dy = torch.randn(B, C, L, generator=gen)
    m, v = ops.rowstats(c.cuda())
    m_ref, v_ref = S.row_stats(c.double())
    assert rel_err(m.cpu(), m_ref) < 1e-6 and rel_err(v.cpu(), v_ref) < 1e-5
    out = T.adain(c.cuda().requires_grad_(True), s.cuda().requires_grad_(True))
    assert rel_err(out.detach().cpu(), S.adain(c.double(), s.double())) < TOL_F32
    cd, sd_ = c.cuda().requires_grad_(True), s.cuda().requires_grad_(True)
    T.adain(cd, sd_).backward(dy.cuda())
    dc_ref, ds_ref = S.adain_backward(dy.double(), c.double(), s.double())
    assert rel_err(cd.grad.cpu(), dc_ref) < 5e-5 and rel_err(sd_.grad.cpu(), ds_ref) < 5e-5


@pytest.mark.parametrize("B,C,L", [(2, 6, 11), (4, 144, 128), (2, 50, 300)])
def test_gram_style_loss_simt(T, B, C, L):
    Lb = T._lib
    gen = torch.Generator().manual_seed(5)
    a = torch.randn(B, C, L, generator=gen)
    s = torch.randn(B, C, L, generator=gen) * 1.5
    ad, sd_ = a.cuda().requires_grad_(True), s.cuda().requires_grad_(True)
    loss = T.gram_style_loss(ad, sd_, engine=Lb.ENGINE_SIMT)
    ref = S.gram_style_loss(a.double(), s.double())
    assert abs(float(loss) - float(ref)) <= 1e-5 * abs(float(ref))
    (3.0 * loss).backward()
    da_ref, ds_ref = S.gram_style_loss_backward(a.double(), s.double())
    assert rel_err(ad.grad.cpu(), 3.0 * da_ref) < 2e-5 and rel_err(sd_.grad.cpu(), 3.0 * ds_ref) < 2e-5


@pytest.mark.parametrize("B,C,L", [(2, 16, 64), (4, 144, 128), (2, 50, 300)])
def test_gram_style_loss_tcgen05(T, B, C, L):
    Lb, ops = T._lib, T.ops
    if not Lb.load().tsc_device_supports_tcgen05():
        pytest.skip("device has no tcgen05")
    gen = torch.Generator().manual_seed(6)
    a = torch.randn(B, C, L, generator=gen)
    s = torch.randn(B, C, L, generator=gen) * 1.5
    ad, sd_ = a.cuda().requires_grad_(True), s.cuda().requires_grad_(True)
    loss = T.gram_style_loss(ad, sd_, engine=Lb.ENGINE_TCGEN05)
    torch.cuda.synchronize()
    assert ops.read_watchdog() == 0
    ref = S.gram_style_loss(a.double(), s.double())
    assert abs(float(loss) - float(ref)) <= TOL_BF16 * abs(float(ref))
    loss.backward()
    da_ref, ds_ref = S.gram_style_loss_backward(a.double(), s.double())
    assert rel_err(ad.grad.cpu(), da_ref) < TOL_BF16 and rel_err(sd_.grad.cpu(), ds_ref) < TOL_BF16


def test_pack_weights_multi_matches_per_layer_pack(T):
    """One launch for a list of banks == the per-layer pack (bit-exact), including the in-place masking of W."""
    ops, L = T.ops, T._lib
    names = ["small7", "mid13", "even2", "cfg2_l1", "cfg2_l0", "cfg2_l2", "cfg4_l1", "one", "cfg2_l1"]
    jobs, refs = [], []
    for k, name in enumerate(names):
        layer, g, x, w, b, dy = make_case(name, seed=20 + k)
        geom = ops.bank_geometry(layer)
        w1, w2 = w.clone().cuda(), w.clone().cuda()
        zero = k % 2 == 0
        refs.append((ops.pack_weights_pair(geom, w1, L.TSC_BF16, zero, True), w1))
        jobs.append((geom, w2, zero, True))
    got = ops.pack_weights_multi(jobs, L.TSC_BF16)
    torch.cuda.synchronize()
    for k, (((pf_ref, pd_ref), w_ref), (pf, pd), job) in enumerate(zip(refs, got, jobs)):
        assert torch.equal(pf.view(torch.int16), pf_ref.view(torch.int16)), (names[k], "fwd")
        assert torch.equal(pd.view(torch.int16), pd_ref.view(torch.int16)), (names[k], "dgrad")
        assert torch.equal(job[1], w_ref), (names[k], "masked W")


def test_bad_arguments_raise(T):
    ops, L = T.ops, T._lib
    with pytest.raises(RuntimeError):
        ops.ncl_to_c8(torch.randn(2, 3, 4), L.TSC_F32)            # CPU tensor
    with pytest.raises(ValueError):
        ops.dense_geometry(3, 3000, 1)                             # beyond TSC_MAX_CHANNELS_WIDE
    wide = ops.dense_geometry(3, 300, 1)                           # wider than one TMEM tile: accepted, CUDA-core engine only
    assert wide.wide
    with pytest.raises(RuntimeError):
        x8 = ops.ncl_to_c8(torch.randn(2, 3, 16, device="cuda"), L.TSC_BF16)
        wf = ops.pack_weights(wide, torch.zeros(300, 3, 1, device="cuda"), L.DIR_FWD, L.TSC_BF16, False)
        ops.osconv(L.ENGINE_TCGEN05, L.DIR_FWD, wide, x8, wf, torch.zeros(300, device="cuda"))
    with pytest.raises(ValueError):
        ops.bank_geometry([(1, 2, 3), (1, 2, 1), (1, 2, 3)])       # not nested
    geom = ops.dense_geometry(3, 4, 1)
    with pytest.raises(RuntimeError):
        ops.pack_weights(geom, torch.zeros(4, 3, 2, device="cuda"), L.DIR_FWD, L.TSC_F32, False)


@pytest.mark.parametrize("B,C,K", [(128, 144, 6), (37, 50, 4), (5, 225, 33)])
def test_head_linear_cross_entropy_against_torch(T, B, C, K):
    """Fused classifier head (OS_CNN.py:108-109 + nn.CrossEntropyLoss, train_and_test.py:593-603): logits, loss and every
    gradient against torch autograd in float64, with and without an extra gradient arriving at the logits, and adding into
    existing gradient buffers."""
    from feature_level_style_transfer_for_tsc_b200 import functional as TF
    g = torch.Generator().manual_seed(3)
    pooled = torch.randn(B, C, generator=g)
    W = torch.randn(K, C, generator=g) * 0.2
    bias = torch.randn(K, generator=g)
    y = torch.randint(0, K, (B,), generator=g)
    extra = torch.randn(B, K, generator=g)
    pd, Wd, bd = (t.double().requires_grad_(True) for t in (pooled, W, bias))
    lg_ref = torch.nn.functional.linear(pd, Wd, bd)
    loss_ref = torch.nn.functional.cross_entropy(lg_ref, y)
    (2.5 * loss_ref + (lg_ref * extra.double()).sum()).backward()
    pc, Wc, bc = (t.cuda().requires_grad_(True) for t in (pooled, W, bias))
    lg, loss = TF.head_cross_entropy(pc, Wc, bc, y.cuda())
    (2.5 * loss + (lg * extra.cuda()).sum()).backward()
    torch.cuda.synchronize()
    assert rel_err(lg.detach().cpu(), lg_ref.detach()) < 1e-5
    assert abs(float(loss) - float(loss_ref)) < 1e-5 * abs(float(loss_ref))
    for got, ref in ((pc.grad, pd.grad), (Wc.grad, Wd.grad), (bc.grad, bd.grad)):
        assert rel_err(got.cpu(), ref) < 2e-5
    # logits only (no labels): the eval-mode call of the trainer (train_and_test.py:584-586); gradient through the logits
    pc2 = pooled.cuda().requires_grad_(True)
    lg2, none = TF.head_cross_entropy(pc2, Wc.detach(), bc.detach(), None)
    assert none is None and torch.equal(lg2.detach(), lg.detach())
    (lg2 * extra.cuda()).sum().backward()
    assert rel_err(pc2.grad.cpu(), (extra.double() @ W.double())) < 2e-5
    # direct accumulation into existing .grad buffers (flat data-parallel bucket)
    TF.set_direct_grads(True)
    try:
        W3, b3 = W.cuda().requires_grad_(True), bias.cuda().requires_grad_(True)
        W3.grad, b3.grad = torch.ones_like(W3), torch.ones_like(b3)
        _, loss3 = TF.head_cross_entropy(pooled.cuda(), W3, b3, y.cuda())
        loss3.backward()
    finally:
        TF.set_direct_grads(False)
    Wd.grad = None; bd.grad = None
    torch.nn.functional.cross_entropy(torch.nn.functional.linear(pooled.double(), Wd, bd), y).backward()
    assert rel_err((W3.grad - 1).cpu(), Wd.grad) < 2e-5 and rel_err((b3.grad - 1).cpu(), bd.grad) < 2e-5
    # the total-loss kernel
    a, b = torch.tensor(1.5, device="cuda", requires_grad=True), torch.tensor(-2.0, device="cuda", requires_grad=True)
    tot = TF.weighted_loss_sum([a, b, loss.detach()], [1.0, 3.0, 0.5])
    tot.backward()
    assert abs(float(tot) - (1.5 - 6.0 + 0.5 * float(loss))) < 1e-5 and float(a.grad) == 1.0 and float(b.grad) == 3.0
