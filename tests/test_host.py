"""CPU-side checks of the host logic and the C-ABI library (no compute calls: there is no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT, as_lpl
from oracle import os_cnn as O


def test_library_builds_loads_and_exports_every_declared_symbol():
    import feature_level_style_transfer_for_tsc_b200 as T
    T._lib.build()
    lib = T._lib.load()
    header = open(os.path.join(ROOT, "include", "tsc_b200.h")).read()
    declared = set(re.findall(r"\b(tsc_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations found"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/tsc_b200.h but not exported"
    assert declared == set(T._lib.SIGNATURES), declared ^ set(T._lib.SIGNATURES)
    assert lib.tsc_version() == 100
    assert lib.tsc_pad_channels(9) == 16 and lib.tsc_pad_channels(228) == 240


def test_sass_contains_blackwell_native_instructions():
    """tcgen05.mma -> UTC*MMA, TMA -> UTMALDG / UBLKCP, tcgen05.ld -> LDTM (B200_PROFILING.md)."""
    import shutil
    import subprocess
    import feature_level_style_transfer_for_tsc_b200 as T
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    T._lib.build()
    sass = subprocess.run(["cuobjdump", "-sass", T._lib.LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "UTMALDG", "UBLKCP", "LDTM"):
        assert mnemonic in sass, mnemonic


def test_packed_weight_sizes_follow_the_live_taps():
    import feature_level_style_transfer_for_tsc_b200 as T
    ops, L = T.ops, T._lib
    ext, _ = O.trainer_layer_lists(9, 128)
    g = ops.bank_geometry(ext[1])
    og = O.bank_geometry(ext[1])
    assert g.s_of_tap == og["s_of_tap"] and (g.pad_l, g.pad_r) == (og["pad_l"], og["pad_r"])
    assert g.live_macs_per_position() == O.live_macs_per_position(ext[1])
    fwd = g.packed_bytes(L.DIR_FWD, L.TSC_BF16)
    dense = g.kmax * g.cin_p * g.cout_p * 2
    rows = sum((g.cin_p // 8) * (g.cout_p - (0 if s < 16 else (s // 16) * 16)) for s in g.s_of_tap)
    assert fwd == rows * 16 and fwd < 0.6 * dense               # the 43 %-dense bank is stored sparsely
    dg = g.packed_bytes(L.DIR_DGRAD, L.TSC_BF16)
    rows_d = sum((g.cout_p // 8 - (s // 16) * 2) * g.cin_p for s in g.s_of_tap)
    assert dg == rows_d * 16


def test_modules_mirror_reference_interface(tables):
    from feature_level_style_transfer_for_tsc_b200.OS_CNN import OS_CNN as M
    from feature_level_style_transfer_for_tsc_b200.OS_CNN import OS_CNN_Structure_build as SB
    for kmax, row in tables["mask_index"].items():
        for k, (lo, hi) in row.items():
            assert M.calculate_mask_index(int(k), int(kmax)) == (lo, hi)
    for key, raw in tables["layer_lists"].items():
        C, Ln = (int(s[1:]) for s in key.split("_"))
        budgets = [8 * 128 * C, 5 * 128 * 256 + 2 * 256 * 128]
        assert SB.generate_layer_parameter_list(1, min(int(Ln / 4), 89), budgets, C) == as_lpl(raw)
    assert SB.get_Prime_number_in_a_range(1, 31) == tables["primes_1_31"]
    lpl = as_lpl(tables["small"]["lpl_ext"])
    torch.manual_seed(0)
    fe = M.OS_CNN_res(lpl)
    cl = M.OS_CNN(as_lpl(tables["small"]["lpl_cls"]), tables["small"]["n_class"])
    assert {k: list(v.shape) for k, v in fe.state_dict().items()} == tables["state_dict"]["small_fe"]
    assert {k: list(v.shape) for k, v in cl.state_dict().items()} == tables["state_dict"]["small_cl"]
    assert isinstance(fe.return_last_layer(), M.OS_block) and len(list(fe.return_last_layer().parameters())) == 12
    assert cl.length_before_classification == 40 and isinstance(cl.hidden, torch.nn.Linear)
    assert fe.net_1.net.net[0].weight_mask.shape == fe.net_1.net.net[0].conv1d.weight.shape
    mask = torch.from_numpy(O.build_mask(lpl[1]))
    assert torch.equal(fe.net_1.net.net[1].weight_mask, mask)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fe(torch.randn(2, 3, 32))


def test_no_product_import_of_the_oracle():
    """The product must never route through oracle/ (or any CPU fallback)."""
    pkg = os.path.join(ROOT, "feature_level_style_transfer_for_tsc_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(".py"):
                src = open(os.path.join(dirpath, fn)).read()
                assert "import oracle" not in src and "from oracle" not in src, fn


def test_multi_source_model_set_host_logic():
    """Configuration 3 on the host: per-pair parameter groups (learning rates and the critic's WGAN clip of
    train_and_test.py:97-105,763-764), the reversal schedule, and no CPU compute path."""
    from feature_level_style_transfer_for_tsc_b200.train_step import FlatParameters, MultiSourceModelSet
    torch.manual_seed(0)
    model = MultiSourceModelSet((2, 96, 3), [(1, 96, 2), (2, 128, 3)], critic_hidden=8)
    groups = model.parameter_groups()
    assert len(groups) == 12 and [g[1] for g in groups[:6]] == [0.001, 0.003, 0.001, 0.001, 0.003, 0.001]
    assert [g[2] for g in groups[:6]] == [0.0] * 5 + [0.0005]
    flat = FlatParameters(groups)
    assert flat.group_clamp.count(0.0005) == 2 and flat.group_end[-1] == flat.flat_p.numel()
    assert all(p.grad.data_ptr() >= flat.flat_g.data_ptr() for p in flat.params)
    # the critic's schedule: two calls per step, saturating at max_iter, restorable (graph warm-up leaves no trace)
    pair = model.pairs[0]
    st = pair.schedule_state()
    assert st == (-1, 0.001)
    vals = pair.ad_net.advance_schedule(2)
    assert vals[0] == 0.0 and abs(vals[1] - 0.9866142981514305) < 1e-15 and vals[2] == vals[1] and pair.ad_net.iter_num == 1
    for _ in range(20):
        pair.ad_net.advance_schedule(2)
    assert pair.ad_net.iter_num == 20.0 and pair.ad_net.coeff == 1.0
    pair.set_schedule_state(st)
    assert pair.ad_net.iter_num == -1
    pair.ad_net.eval()
    assert pair.ad_net.advance_schedule(2)[0] == pair.ad_net.coeff and pair.ad_net.iter_num == -1     # eval calls do not count
    x = torch.zeros(2, 2, 96)
    y = torch.zeros(2, dtype=torch.long)
    with pytest.raises(RuntimeError):
        model(x, y, torch.zeros(2, 1, 96), y, torch.zeros(2, 2, 128), y)


def test_cdan_mirror_interface():
    """Names and signatures of the reference's C_DAN.py / widgets.py critic."""
    import inspect
    from feature_level_style_transfer_for_tsc_b200 import C_DAN, widgets
    assert list(inspect.signature(C_DAN.CDAN).parameters) == [
        "input_target", "input_g_from_source", "prob_target", "prob_g_from_source", "ad_net", "random_layer"]
    assert list(inspect.signature(C_DAN.RandomLayer.__init__).parameters)[1:] == ["input_dim_list", "output_dim", "with_nvidia"]
    torch.manual_seed(0)
    rl = C_DAN.RandomLayer([12, 3], with_nvidia=False)
    torch.manual_seed(0)
    assert torch.equal(rl.random_matrix[0], torch.randn(12, 1024)) and rl.scale_div == 32.0
    ad = widgets.AdversarialNetworkforCDAN(1024, 16)
    assert list(ad.state_dict()) == ["ad_layer1.weight", "ad_layer1.bias", "ad_layer2.weight", "ad_layer2.bias",
                                     "ad_layer3.weight", "ad_layer3.bias"]
    assert float(ad.ad_layer1.bias.abs().max()) == 0.0 and (ad.iter_num, ad.alpha, ad.max_iter) == (-1, 100.0, 20.0)
    assert abs(widgets.calc_coeff(1, 1.0, 0.0, 100.0, 20.0) - (2.0 / (1.0 + np.exp(-5.0)) - 1.0)) < 1e-15
    with pytest.raises(RuntimeError):        # no CPU path
        C_DAN.CDAN(torch.zeros(2, 3, 4), torch.zeros(2, 3, 4), torch.zeros(2, 3), torch.zeros(2, 3), ad, rl)
    with pytest.raises(RuntimeError):        # no stand-alone torch route of the projection
        rl([torch.zeros(2, 12), torch.zeros(2, 3)])


def test_default_engine_is_the_tensor_core_path_for_every_family():
    """A process that never calls set_engine must still run conv, wgrad AND the Gram loss on tcgen05 (a default that left
    the Gram family on the SIMT checker went unnoticed in round 1 because bench.py always calls set_engine)."""
    import importlib
    import feature_level_style_transfer_for_tsc_b200.ops as ops
    fresh = importlib.reload(ops)
    try:
        L = fresh.L
        assert [fresh.get_engine(f) for f in ("conv", "wgrad", "gram")] == [L.ENGINE_TCGEN05] * 3
        assert fresh.engine_name() == "tcgen05" and fresh.op_dtype() == L.TSC_BF16
    finally:
        importlib.reload(ops)


def test_short_series_layers_are_marked_wide_instead_of_rejected():
    """train_and_test.py:38-53 gives > 256-channel layers for series shorter than ~80 samples; the geometry accepts them (up to
    TSC_MAX_CHANNELS_WIDE) and marks them for the fp32 CUDA-core engine."""
    from feature_level_style_transfer_for_tsc_b200 import _lib as L
    from feature_level_style_transfer_for_tsc_b200.OS_CNN.OS_CNN import OS_CNN, OS_CNN_res
    from feature_level_style_transfer_for_tsc_b200.train_step import trainer_layer_lists
    for Ln, widest in ((64, 336), (32, 560), (16, 1020)):
        ext, cls, cf = trainer_layer_lists(3, Ln)
        fe, cl = OS_CNN_res(ext), OS_CNN(cls, 4)
        flags = [layer.geometry.wide for layer in fe.net_1.net.net]
        assert max(sum(p[1] for p in layer) for layer in ext) == widest and any(flags)
        assert cl.net[0].geometry.wide          # the classifier's first bank reads the widest feature map
    assert L.MAX_CHANNELS == 256 and L.MAX_CHANNELS_WIDE == 2048


def test_remaining_ranges_of_the_gradient_bucket():
    """What the data-parallel trainer still has to all-reduce after the early (overlapped) slices: every element of the
    bucket exactly once, in as few contiguous ranges as possible."""
    from feature_level_style_transfer_for_tsc_b200.train_step import remaining_ranges
    assert remaining_ranges([], 10) == [(0, 10)]
    assert remaining_ranges([(0, 10)], 10) == []
    assert remaining_ranges([(2, 4), (7, 10)], 10) == [(0, 2), (4, 7)]
    assert remaining_ranges([(7, 9), (2, 4)], 10) == [(0, 2), (4, 7), (9, 10)]          # any order
    assert remaining_ranges([(2, 5), (4, 6)], 8) == [(0, 2), (6, 8)]                     # overlapping slices
    # the cfg2 groups (fe_t, cl_t, fe_s, du, cl_s): the classifiers went early
    ends = [100, 160, 260, 300, 372]
    rng = dict(zip(("fe_t", "cl_t", "fe_s", "du", "cl_s"), zip([0] + ends[:-1], ends)))
    rest = remaining_ranges([rng["cl_t"], rng["cl_s"]], ends[-1])
    assert rest == [(0, 100), (160, 300)]
    covered = sorted(rest + [rng["cl_t"], rng["cl_s"]])
    assert covered[0][0] == 0 and covered[-1][1] == ends[-1] and all(a[1] == b[0] for a, b in zip(covered, covered[1:]))


def test_bench_reference_arm_prints_the_contract_line_and_the_product_arm_fails_loudly_without_a_gpu():
    """`bench.py --impl reference` (the CPU arm the driver launches beside the product arm) prints ONE JSON line with the same
    metric / unit / config as the product arm and `impl: reference`; the product arm has no CPU fallback."""
    import json
    import subprocess
    import sys
    from conftest import ROOT
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    import bench
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == bench.UNIT and d["higher_is_better"] is True
    assert d["config"] == bench.step_config(bench.WORKLOAD, 2 * bench.CFG["B"])          # identical in both arms
    assert d["steps"] == 1 and d["warmup"] == 0 and d["n_gpus"] == 1 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"] > 0
    assert d["e2e"] == dict(value=d["value"], unit=bench.UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    if not torch.cuda.is_available():
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], capture_output=True, text=True,
                           timeout=600, cwd=ROOT)
        assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


def test_device_side_gradnorm_weight_gradient_equals_the_host_formula():
    """grad_norm._weight_gradient_device (torch, used inside the graph-captured joint stage) against grad_norm._weight_gradient
    (numpy, the transcription of train_and_test.py:691-715 that the golden vectors pin): norms, targets and the sign-valued
    weight gradient on random balanced weights / norm sums / losses, both side sizes."""
    from feature_level_style_transfer_for_tsc_b200 import grad_norm as D
    rng = np.random.default_rng(5)
    for n in (2, 3):
        for _ in range(50):
            w = rng.uniform(0.2, 6.0, n).astype(np.float32)
            sums = rng.uniform(0.01, 30.0, n).astype(np.float32)
            lv = rng.uniform(-2.0, 4.0, n).astype(np.float32)
            init = (1 / (1 + np.exp(-rng.uniform(-1.0, 3.0, n)))).astype(np.float32)
            norms, target, grad = D._weight_gradient(w, sums, lv, init, D.ALPHA)
            dn, dt, dg = D._weight_gradient_device(torch.from_numpy(w), torch.from_numpy(sums), torch.from_numpy(lv),
                                                   torch.from_numpy(init), D.ALPHA)
            assert np.allclose(dn.numpy(), norms, rtol=1e-6) and np.allclose(dt.numpy(), target, rtol=2e-5)
            decided = np.abs(norms - target) > 1e-4 * np.abs(target)          # sign(norms - target) away from the tie
            assert np.array_equal(dg.numpy()[decided], grad[decided])
