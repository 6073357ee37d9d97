"""Pins oracle/cdan.py against the reference's own C_DAN.CDAN / RandomLayer / AdversarialNetworkforCDAN
(tests/golden/cdan_small.npz, written by oracle/make_golden.py from the unmodified reference).  CPU only."""
import numpy as np
import torch

from conftest import rel_err
from oracle import cdan as OC


def _ad_sd(z):
    return {k[3:]: torch.from_numpy(z[k]).clone() for k in z.files if k.startswith("ad/")}


def test_init_replay_of_random_layer_and_critic(tables, cdan_small):
    """One seed reproduces the reference's random matrices and the critic's xavier-normal weights bit for bit."""
    t = tables["cdan_small"]
    torch.manual_seed(t["seed"])
    mats = OC.init_random_layer([t["C"] * t["L"], t["n_class"]])
    sd = OC.init_ad_net(1024, t["hidden"])
    assert np.array_equal(mats[0].numpy(), cdan_small["R0"]) and np.array_equal(mats[1].numpy(), cdan_small["R1"])
    for k, v in sd.items():
        assert np.array_equal(v.numpy(), cdan_small["ad/" + k]), k


def test_coefficient_schedule(tables):
    st = OC.AdNetState()
    seen = []
    for training in (True, True, True, True, False, False):
        seen.append(st.advance(training))
    assert seen[0] == 0.0 and abs(seen[1] - 0.9866142981514305) < 1e-15
    t = tables["cdan_small"]
    assert [seen[1], seen[3], seen[5]] == t["coeff_after_call"]
    assert st.iter_num == t["iter_num_after"] == 3
    for _ in range(40):
        st.advance(True)
    assert st.iter_num == 20.0 and st.coeff == 1.0            # saturates (widgets.py:116-117)


def test_cdan_loss_and_gradients_match_the_reference(cdan_small):
    z = cdan_small
    mats = [torch.from_numpy(z["R0"]), torch.from_numpy(z["R1"])]
    st = OC.AdNetState()
    for call, training in enumerate((True, True, False)):
        sd = _ad_sd(z)
        for v in sd.values():
            v.requires_grad_(True)
        ins = [torch.from_numpy(z[f"c{call}/{n}"]).clone().requires_grad_(True) for n in ("ft", "fs", "lt", "ls")]
        loss = OC.cdan(*ins, sd, st, mats, training=training, dropout_p=0.0)
        loss.backward()
        assert abs(float(loss) - float(z[f"c{call}/loss"])) <= 2e-6 * max(1.0, abs(float(z[f"c{call}/loss"])))
        for t, n in zip(ins, ("dft", "dfs", "dlt", "dls")):
            assert rel_err(t.grad, z[f"c{call}/{n}"]) < 2e-5, (call, n)
        for k, v in sd.items():
            assert rel_err(v.grad, z[f"c{call}/dad/{k}"]) < 2e-5, (call, k)
    # call 0 reverses the critic-input gradient with coeff(0) = 0 for the target half: the features still receive
    # the generated half's gradient (coeff(1) = 0.9866), so neither gradient is identically zero
    assert np.abs(z["c0/dft"]).max() == 0.0 and np.abs(z["c0/dfs"]).max() > 0.0
