"""CPU-side checks of the host logic and the C-ABI library (no compute calls: there is no GPU here)."""
import ctypes
import os
import re

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
