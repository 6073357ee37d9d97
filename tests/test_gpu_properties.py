"""Parity at BASELINE.json's FULL sizes through size-independent properties (the oracle needs seconds to minutes there).

* Integer-valued operands (small integers are exact in bf16, their products and every partial sum stay below 2^24 and are
  exact in fp32) make the three tensor-core convolution kernels EXACT: the adjoint identities
      <conv_W(x), dy> = <x, dgrad_W(dy)> = <W, wgrad(dy, x)>
  must then hold bit for bit (evaluated in fp64), whatever the tiling, split or summation order; an impulse input must
  return the masked kernel bank itself, and masked taps of the weight gradient must be exact zeros.
* Row statistics / AdaIN: the output rows carry the style rows' mean and variance; AdaIN is idempotent.
* Gram loss: symmetric in its arguments, zero for identical inputs, invariant under a permutation of the batch.
* Train-mode BatchNorm behind the fused conv epilogue: unit variance / zero mean per channel over B*L.
* Evaluation path: predictions of a batch equal the concatenation of the predictions of its halves, bit for bit.
"""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import os_cnn as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def T():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import feature_level_style_transfer_for_tsc_b200 as pkg
    pkg._lib.load()
    return pkg


FULL = {
    "cfg2_l1": (O.trainer_layer_lists(9, 128)[0][1], 128, 128),        # 72 -> 228, Kmax 31, B=128, L=128
    "cfg2_cl0": (O.trainer_layer_lists(9, 128)[1][0], 128, 128),       # 144 -> 72, Kmax 31
    "cfg2_l2": (O.trainer_layer_lists(9, 128)[0][2], 128, 128),        # 228 -> 144, Kmax 2
    "cfg4_l1": (O.trainer_layer_lists(3, 1024)[0][1], 256, 1024),      # 25 -> 225, Kmax 89, B=256, L=1024
}


def small_ints(shape, gen, lo=-2, hi=2, density=1.0):
    t = torch.randint(lo, hi + 1, shape, generator=gen).float()
    if density < 1.0:
        t = t * (torch.rand(shape, generator=gen) < density).float()
    return t


@pytest.mark.parametrize("name", list(FULL))
def test_integer_operands_make_the_conv_kernels_exact_at_full_size(T, name):
    ops, L = T.ops, T._lib
    layer, B, Ln = FULL[name]
    geom = ops.bank_geometry(layer)
    gen = torch.Generator().manual_seed(5)
    mask = torch.from_numpy(O.build_mask(layer))
    x = small_ints((B, geom.cin, Ln), gen).cuda()
    dy = small_ints((B, geom.cout, Ln), gen, density=0.25).cuda()       # sparse: |dW| <= B*L stays far below 2^24
    W = (small_ints((geom.cout, geom.cin, geom.kmax), gen) * mask).cuda()
    x8, dy8 = ops.ncl_to_c8(x, L.TSC_BF16), ops.ncl_to_c8(dy, L.TSC_BF16)
    wf = ops.pack_weights(geom, W, L.DIR_FWD, L.TSC_BF16, True)
    wd = ops.pack_weights(geom, W, L.DIR_DGRAD, L.TSC_BF16, False)
    y = ops.c8_to_ncl(ops.osconv(L.ENGINE_TCGEN05, L.DIR_FWD, geom, x8, wf, None), geom.cout)
    dx = ops.c8_to_ncl(ops.osconv(L.ENGINE_TCGEN05, L.DIR_DGRAD, geom, dy8, wd, None), geom.cin)
    dW = ops.oswgrad(L.ENGINE_TCGEN05, geom, dy8, x8)
    torch.cuda.synchronize()
    assert ops.read_watchdog() == 0
    for t in (y, dx, dW):
        assert torch.equal(t, t.round()) and float(t.abs().max()) < 2 ** 24      # exact integers
    a = float((y.double() * dy.double()).sum())
    b = float((x.double() * dx.double()).sum())
    c = float((W.double() * dW.double()).sum())
    assert a == b == c, (a, b, c)
    assert float((dW.cpu() * (1 - mask)).abs().max()) == 0.0                       # masked taps: exact zeros
    # the unmasked gradient agrees on the live taps and satisfies its own adjoint identity
    dWd = ops.oswgrad(L.ENGINE_TCGEN05, geom.dense_twin(), dy8, x8)
    assert torch.equal(dWd.cpu() * mask, dW.cpu())
    Wd = small_ints((geom.cout, geom.cin, geom.kmax), gen).cuda()
    gd = geom.dense_twin()
    yd = ops.c8_to_ncl(ops.osconv(L.ENGINE_TCGEN05, L.DIR_FWD, gd, x8, ops.pack_weights(gd, Wd, L.DIR_FWD, L.TSC_BF16, False), None),
                       geom.cout)
    assert float((yd.double() * dy.double()).sum()) == float((Wd.double() * dWd.double()).sum())
    # impulse response: one non-zero input sample returns the (masked) bank, tap-reversed around pad_left
    xi = torch.zeros(B, geom.cin, Ln)
    b0, c0, l0 = B - 1, geom.cin - 1, Ln // 2
    xi[b0, c0, l0] = 1.0
    yi = ops.c8_to_ncl(ops.osconv(L.ENGINE_TCGEN05, L.DIR_FWD, geom, ops.ncl_to_c8(xi.cuda(), L.TSC_BF16), wf, None), geom.cout).cpu()
    want = torch.zeros(geom.cout, Ln)
    for t in range(geom.kmax):
        l = l0 - t + geom.pad_l
        if 0 <= l < Ln:
            want[:, l] = W[:, c0, t].cpu()
    assert torch.equal(yi[b0], want)
    yi[b0] = 0
    assert float(yi.abs().max()) == 0.0                                             # no leakage into other samples


@pytest.mark.parametrize("B,C,Ln", [(128, 144, 128), (256, 50, 1024), (1024, 144, 1024)])
def test_adain_properties_at_full_size(T, B, C, Ln):
    g = torch.Generator(device="cuda").manual_seed(1)
    content = torch.randn(B, C, Ln, device="cuda", generator=g) * 2.0 + 0.5
    style = torch.randn(B, C, Ln, device="cuda", generator=g) * 0.7 - 1.0
    out = T.adain(content, style)
    m_o, v_o = T.ops.rowstats(out)
    m_s, v_s = T.ops.rowstats(style)
    ref_m, ref_v = style.double().mean(-1), style.double().var(-1, unbiased=True)
    assert rel_err(m_s.cpu(), ref_m.cpu()) < 1e-6 and rel_err(v_s.cpu(), ref_v.cpu()) < 1e-5     # Welford kernel vs fp64
    assert float((m_o - m_s).abs().max()) < 2e-5
    # the output variance is sigma_s^2 * var_c / (var_c + eps): equal to the style variance up to eps / var_c
    assert rel_err(v_o.cpu(), (v_s + 1e-5).cpu()) < 1e-4
    again = T.adain(out, style)
    assert rel_err(again.cpu(), out.cpu()) < 5e-5                                    # idempotent (up to eps / var)
    same = T.adain(style, style)
    assert rel_err(same.cpu(), style.cpu()) < 2e-5                                   # identity on its own statistics


@pytest.mark.parametrize("B,C,Ln", [(128, 144, 128), (256, 50, 1024)])
def test_gram_loss_properties_at_full_size(T, B, C, Ln):
    T.set_engine("tcgen05")
    g = torch.Generator(device="cuda").manual_seed(2)
    a = torch.randn(B, C, Ln, device="cuda", generator=g)
    b = torch.randn(B, C, Ln, device="cuda", generator=g)
    lab = float(T.gram_style_loss(a, b))
    lba = float(T.gram_style_loss(b, a))
    assert abs(lab - lba) <= 1e-5 * lab                                              # symmetric
    assert float(T.gram_style_loss(a, a.clone())) <= 1e-10 * lab                      # zero for identical inputs
    perm = torch.randperm(B, device="cuda", generator=g)
    assert abs(float(T.gram_style_loss(a[perm].contiguous(), b[perm].contiguous())) - lab) <= 1e-5 * lab
    # scale law: G is quadratic, the loss quartic
    assert abs(float(T.gram_style_loss(2 * a, 2 * b)) - 16 * lab) <= 1e-5 * 16 * lab
    ref = float(torch.mean((torch.bmm(a.double(), a.double().transpose(1, 2)) - torch.bmm(b.double(), b.double().transpose(1, 2))) ** 2)
                / (C * Ln) ** 2)
    assert abs(lab - ref) <= 1e-4 * ref                                              # 3xTF32 forward against fp64


def test_cfg2_modules_at_full_size(T):
    """Train-mode BatchNorm statistics behind the fused epilogue over B*L = 16384 positions, and the batch-independence of
    the evaluation path, at configuration 2's shape."""
    from feature_level_style_transfer_for_tsc_b200.OS_CNN.OS_CNN import OS_CNN, OS_CNN_res, build_layer_with_layer_parameter
    T.set_engine("tcgen05")
    C, Ln, K, B = 9, 128, 6, 128
    lpl_e, lpl_c = O.trainer_layer_lists(C, Ln)
    torch.manual_seed(0)
    layer = build_layer_with_layer_parameter(lpl_e[1], relu_or_not_at_last_layer=False).cuda().train()
    x = torch.randn(B, 72, Ln, device="cuda") * 3.0 + 1.0
    with torch.no_grad():
        z = layer(x)
    m, v = z.double().mean(dim=(0, 2)), z.double().var(dim=(0, 2), unbiased=False)
    assert float(m.abs().max()) < 1e-4 and float((v - 1).abs().max()) < 2e-3          # gamma = 1, beta = 0, eps = 1e-5
    fe, cl = OS_CNN_res(lpl_e).cuda(), OS_CNN(lpl_c, K).cuda()
    xs, _ = O.synthetic_batch(B, C, Ln, K, 0)
    xs = xs.cuda()
    with torch.no_grad():
        for _ in range(2):
            cl(fe(xs))                                                                # running statistics
        fe.eval(); cl.eval()
        whole_logits, whole = cl(fe(xs))
        parts = [cl(fe(xs[:64].contiguous())), cl(fe(xs[64:].contiguous()))]
    assert torch.equal(whole, torch.cat([p[1] for p in parts]))                      # pooled features: this repo's kernels
    assert rel_err(torch.cat([p[0] for p in parts]).cpu(), whole_logits.cpu()) < 1e-6     # the head GEMM is cuBLAS (may re-tile)
    torch.cuda.synchronize()
    assert T.ops.read_watchdog() == 0
