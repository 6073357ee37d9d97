#!/usr/bin/env python
"""Generate tests/golden/* by running the UNMODIFIED reference modules on CPU.

TEST INFRASTRUCTURE ONLY.  Run in the build container (where /root/reference exists):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

The reference never seeds and hard-codes ``.cuda()`` in its constructors
(OS_CNN/OS_CNN.py:56), so this script (a) seeds torch itself and (b) installs the
``torch.Tensor.cuda = identity`` shim described in SURVEY.md F2.  Nothing from the
reference is copied: it is imported, executed, and only its numeric outputs are stored.
The GPU box has no /root/reference -- tests there read the committed vectors only.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

REF = os.environ.get("TSC_REFERENCE_ROOT", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def import_reference():
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.insert(0, REF)
    torch.Tensor.cuda = lambda self, *a, **k: self          # SURVEY F2 shim
    torch.nn.Module.cuda = lambda self, *a, **k: self
    from OS_CNN import OS_CNN as ref_os                      # noqa
    from OS_CNN import OS_CNN_Structure_build as ref_sb      # noqa
    return ref_os, ref_sb


def to_np(sd):
    return {k: v.detach().cpu().numpy().copy() for k, v in sd.items()}


def run_pair(ref_os, lpl_ext, n_class, x, y, seed):
    """extractor + classifier, one training forward/backward, then one eval forward."""
    from oracle import os_cnn as O
    torch.manual_seed(seed)
    fe = ref_os.OS_CNN_res(lpl_ext)
    cf = O.feature_channels(lpl_ext)
    lpl_cls = ref_os.layer_parameter_list_input_change(lpl_ext, cf)
    cl = ref_os.OS_CNN(lpl_cls, n_class)
    init_fe, init_cl = to_np(fe.state_dict()), to_np(cl.state_dict())

    fe.train(); cl.train()
    xin = x.clone().requires_grad_(True)
    feat = fe(xin)
    feat.retain_grad()
    logits, pooled = cl(feat)
    loss = torch.nn.functional.cross_entropy(logits, y)
    loss.backward()
    out = dict(x=x.numpy(), y=y.numpy(), feat=feat.detach().numpy(), logits=logits.detach().numpy(),
               pooled=pooled.detach().numpy(), loss=np.float64(loss.item()),
               dfeat=feat.grad.numpy(), dx=xin.grad.numpy())
    grads = {}
    for name, mod in (("fe", fe), ("cl", cl)):
        for k, p in mod.named_parameters():
            grads[f"{name}.{k}"] = p.grad.detach().numpy()
    after_fe, after_cl = to_np(fe.state_dict()), to_np(cl.state_dict())
    fe.eval(); cl.eval()
    with torch.no_grad():
        feat_e = fe(x)
        logits_e, pooled_e = cl(feat_e)
    out.update(feat_eval=feat_e.numpy(), logits_eval=logits_e.numpy(), pooled_eval=pooled_e.numpy())
    return init_fe, init_cl, after_fe, after_cl, grads, out, lpl_cls


def make_cdan_golden():
    """Two consecutive training-mode calls (the schedule advances: coeff 0 / 0.9866 then ~1) and one eval-mode call of
    the reference's CDAN on seeded tensors; dropout p = 0 (its mask is generator-specific)."""
    if not hasattr(np, "float"):
        np.float = float                                      # widgets.py:13 / C_DAN.py:44 (numpy < 1.24 spelling)
    import C_DAN as ref_cdan                                  # noqa
    import widgets as ref_w                                   # noqa
    B, C, L, K, H = 5, 6, 8, 4, 32
    torch.manual_seed(3)
    rl = ref_cdan.RandomLayer([C * L, K], with_nvidia=False)
    ad = ref_w.AdversarialNetworkforCDAN(1024, H)
    ad.dropout1.p = 0.0
    ad.dropout2.p = 0.0
    g = torch.Generator().manual_seed(5)
    out = {"R0": rl.random_matrix[0].numpy().copy(), "R1": rl.random_matrix[1].numpy().copy()}
    out.update({f"ad/{k}": v.detach().numpy().copy() for k, v in ad.state_dict().items()})
    coeffs = []
    for call, training in enumerate((True, True, False)):
        ad.train(training)
        ft = torch.randn(B, C, L, generator=g).requires_grad_(True)
        fs = torch.randn(B, C, L, generator=g).requires_grad_(True)
        lt = torch.randn(B, K, generator=g).requires_grad_(True)
        ls = torch.randn(B, K, generator=g).requires_grad_(True)
        for p in ad.parameters():
            p.grad = None
        loss = ref_cdan.CDAN(ft, fs, lt, ls, ad, rl)
        loss.backward()
        coeffs.append(float(ad.coeff))
        out.update({f"c{call}/ft": ft.detach().numpy(), f"c{call}/fs": fs.detach().numpy(),
                    f"c{call}/lt": lt.detach().numpy(), f"c{call}/ls": ls.detach().numpy(),
                    f"c{call}/loss": np.float64(loss.item()),
                    f"c{call}/dft": ft.grad.numpy(), f"c{call}/dfs": fs.grad.numpy(),
                    f"c{call}/dlt": lt.grad.numpy(), f"c{call}/dls": ls.grad.numpy()})
        out.update({f"c{call}/dad/{k}": p.grad.numpy().copy() for k, p in ad.named_parameters()})
    np.savez_compressed(os.path.join(OUT, "cdan_small.npz"), **out)
    return {"B": B, "C": C, "L": L, "n_class": K, "hidden": H, "seed": 3, "coeff_after_call": coeffs,
            "iter_num_after": float(ad.iter_num)}


def _ref_lines(rel_path, first, last):
    """Source lines [first, last] (1-based, inclusive) of a reference file, dedented -- executed, never stored."""
    import textwrap
    with open(os.path.join(REF, rel_path), encoding="utf-8") as f:
        lines = f.readlines()
    return textwrap.dedent("".join(lines[first - 1:last]))


def make_voting_golden():
    """Runs the reference script's own lines (multi_source_voting.py:294-307 per model, 358-367, 406-424) on seeded
    logits of three models; class 4 is never predicted by any model (the 0/0 -> nan_to_num branch)."""
    from scipy.stats import entropy
    from sklearn.metrics import accuracy_score
    rng = np.random.default_rng(17)
    K, n_train, n_test = 5, 90, 61
    label_list_train = rng.integers(0, K - 1, n_train).astype(np.float64)      # np.concatenate of y.numpy() onto float64
    label_list = rng.integers(0, K - 1, n_test).astype(np.float64)
    ns = {"np": np, "entropy": entropy, "accuracy_score": accuracy_score, "target_num_class": K,
          "label_list_train": label_list_train, "label_list": label_list}
    out = {"label_list_train": label_list_train, "label_list": label_list}
    src_prec = _ref_lines("multi_source_voting.py", 294, 307)
    for m in (1, 2, 3):
        tr = (rng.standard_normal((n_train, K)) * 2.0).astype(np.float32)
        tr[np.arange(n_train), label_list_train.astype(int)] += 1.5 * (m / 3.0)     # models of different quality
        tr[:, K - 1] = -30.0                                                         # never the argmax
        te = (rng.standard_normal((n_test, K)) * 2.0).astype(np.float32)
        te[np.arange(n_test), label_list.astype(int)] += 1.0
        te[:, K - 1] -= 4.0
        out[f"train_logits{m}"], out[f"test_logits{m}"] = tr.copy(), te.copy()
        ns[f"results_of_train{m}"] = tr.copy()
        ns[f"results_of_probs{m}"] = te.copy()
        exec(src_prec.replace("results_of_train1", f"results_of_train{m}").replace("weight_for_1", f"weight_for_{m}"), ns)
        out[f"precision{m}"] = np.array(ns[f"weight_for_{m}"], dtype=np.float64)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        exec(_ref_lines("multi_source_voting.py", 358, 367), ns)
    for m in (1, 2, 3):
        out[f"weight{m}"] = np.asarray(ns[f"weight_{m}"], dtype=np.float64)
    exec(_ref_lines("multi_source_voting.py", 406, 424), ns)
    out["score"] = np.asarray(ns["result_final"])
    out["predict"] = np.asarray(ns["predict_list"])
    out["acc"] = np.float64(ns["acc"])
    np.savez_compressed(os.path.join(OUT, "voting_small.npz"), **out)
    return out


def make_gradnorm_golden():
    """Executes train_and_test.py:500-511 (GradNorm weights, their Adam optimizers) once and :646-766 (loss stacking,
    first backward, per-loss norms over return_last_layer().parameters(), weight gradient, graph-clearing second
    backward, optimizer steps, renormalisation, WGAN clamps) for two consecutive batches over the reference's modules."""
    ref_os, ref_sb = import_reference()
    from oracle import os_cnn as O
    from oracle import grad_norm as GN
    lpl = ref_sb.generate_layer_parameter_list(1, 7, [216, 2160], 3)           # the "small" bank: widths 20 / 30 / 40
    cf = O.feature_channels(lpl)
    lpl_cls = ref_os.layer_parameter_list_input_change(lpl, cf)
    K, B, C, Ln, style_weight = 4, 6, 3, 32, 1.0e3
    torch.manual_seed(5)
    fe_t = ref_os.OS_CNN_res(lpl); cl_t = ref_os.OS_CNN(lpl_cls, K)
    fe_s = ref_os.OS_CNN_res(lpl); cl_s = ref_os.OS_CNN(lpl_cls, K)
    mods = (fe_t, cl_t, fe_s, cl_s)
    names = ("fe_t", "cl_t", "fe_s", "cl_s")
    out = {}
    for nm, m in zip(names, mods):
        m.train()
        out.update({f"init/{nm}/{k}": v.detach().numpy().copy() for k, v in m.state_dict().items()})
    ad_net = torch.nn.Linear(4, 3)               # only their parameters are touched (the WGAN clamps, :763-766)
    feature_discriminator_s = torch.nn.Linear(4, 3)
    dummy = torch.nn.Parameter(torch.zeros(1))
    ns = {"torch": torch, "nn": torch.nn, "np": np, "with_nvidia": False,
          "target_feature_extraction_module": fe_t, "source_feature_extraction_module": fe_s,
          "ad_net": ad_net, "feature_discriminator_s": feature_discriminator_s,
          "optimizer_list": [torch.optim.RMSprop(fe_t.parameters(), lr=0.001), torch.optim.RMSprop(cl_t.parameters(), lr=0.003),
                             torch.optim.RMSprop(fe_s.parameters(), lr=0.001), torch.optim.RMSprop(cl_s.parameters(), lr=0.003)],
          "optimizer_sl_cpc": torch.optim.Adam([dummy], lr=0.002)}
    exec(_ref_lines("train_and_test.py", 500, 511), ns)
    body = _ref_lines("train_and_test.py", 646, 766)
    g = torch.Generator().manual_seed(23)
    for b in range(2):
        xt = torch.randn(B, C, Ln, generator=g); yt = torch.randint(0, K, (B,), generator=g)
        xs = torch.randn(B, C, Ln, generator=g); ys = torch.randint(0, K, (B,), generator=g)
        losses = GN.named_losses(mods, xt, yt, xs, ys, style_weight)
        out.update({f"b{b}/xt": xt.numpy(), f"b{b}/yt": yt.numpy(), f"b{b}/xs": xs.numpy(), f"b{b}/ys": ys.numpy()})
        out.update({f"b{b}/loss/{k}": np.float64(v.item()) for k, v in losses.items()})
        out[f"b{b}/weights_t_before"] = ns["weights_grad_norm_t"].detach().numpy().copy()
        out[f"b{b}/weights_s_before"] = ns["weights_grad_norm_s"].detach().numpy().copy()
        ns.update(losses)
        ns["cur_epoch"] = 0
        exec(body, ns)
        for side in ("t", "s"):
            out[f"b{b}/norms_{side}"] = ns[f"norms_{side}_stack"].detach().numpy().copy()
            out[f"b{b}/target_{side}"] = ns[f"constant_term_{side}"].detach().numpy().copy()
            out[f"b{b}/grad_w_{side}"] = ns[f"grad_for_weight_{side}"].detach().numpy().copy()
            out[f"b{b}/weights_{side}_after"] = ns[f"weights_grad_norm_{side}"].detach().numpy().copy()
            out[f"b{b}/initial_{side}"] = np.asarray(ns[f"initial_loss_{side}"]).copy()
        for nm, m in zip(names, mods):
            out.update({f"b{b}/grad/{nm}/{k}": p.grad.detach().numpy().copy() for k, p in m.named_parameters()})
            # the state every module is in after this batch: RMSprop's first update is sign-like, so rounding-level
            # gradients (conv biases in front of a BatchNorm) move by O(lr) -- a test of batch b+1 starts from here
            out.update({f"b{b}/after/{nm}/{k}": v.detach().numpy().copy() for k, v in m.state_dict().items()
                        if "num_batches" not in k})
    out["meta"] = np.array(json.dumps({"lpl": lpl, "lpl_cls": lpl_cls, "n_class": K, "B": B, "C": C, "L": Ln,
                                       "style_weight": style_weight, "seed": 5, "cur_epoch": 0}))
    np.savez_compressed(os.path.join(OUT, "gradnorm_small.npz"), **out)
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    sys.path.insert(0, os.path.dirname(HERE))
    if len(sys.argv) > 1 and sys.argv[1] == "--drivers":       # only the 8f-row vectors (leaves the other files untouched)
        make_voting_golden()
        make_gradnorm_golden()
        for fn in ("voting_small.npz", "gradnorm_small.npz"):
            print(f"  {fn}: {os.path.getsize(os.path.join(OUT, fn))} bytes")
        return
    ref_os, ref_sb = import_reference()
    from oracle import os_cnn as O

    # ---- integer tables: mask indices, layer lists, state_dict keys ---------------------------
    tables = {"mask_index": {}, "layer_lists": {}, "state_dict": {}}
    for kmax in (2, 7, 31, 89):
        primes = ref_sb.get_Prime_number_in_a_range(1, kmax)
        tables["mask_index"][str(kmax)] = {str(k): list(ref_os.calculate_mask_index(k, kmax)) for k in primes}
    for (C, L) in ((1, 128), (9, 128), (3, 1024), (7, 1152), (2, 64)):
        budgets = [8 * 128 * C, 5 * 128 * 256 + 2 * 256 * 128]           # train_and_test.py:38
        rf = min(int(L / 4), 89)                                           # train_and_test.py:40-42
        lpl = ref_sb.generate_layer_parameter_list(1, rf, budgets, C)
        tables["layer_lists"][f"C{C}_L{L}"] = lpl
    tables["primes_1_31"] = ref_sb.get_Prime_number_in_a_range(1, 31)
    tables["primes_1_89"] = ref_sb.get_Prime_number_in_a_range(1, 89)

    # ---- small model with full weights, outputs and gradients ---------------------------------
    lpl_small = ref_sb.generate_layer_parameter_list(1, 7, [216, 2160], 3)   # widths 20 / 30 / 40
    g = torch.Generator().manual_seed(7)
    x = torch.randn(6, 3, 32, generator=g)
    y = torch.randint(0, 4, (6,), generator=g)
    init_fe, init_cl, after_fe, after_cl, grads, out, lpl_cls = run_pair(ref_os, lpl_small, 4, x, y, seed=0)
    tables["small"] = {"lpl_ext": lpl_small, "lpl_cls": lpl_cls, "n_class": 4, "seed": 0}
    tables["state_dict"]["small_fe"] = {k: list(v.shape) for k, v in init_fe.items()}
    tables["state_dict"]["small_cl"] = {k: list(v.shape) for k, v in init_cl.items()}
    np.savez_compressed(
        os.path.join(OUT, "small_pair.npz"),
        **{f"init_fe/{k}": v for k, v in init_fe.items()}, **{f"init_cl/{k}": v for k, v in init_cl.items()},
        **{f"after_fe/{k}": v for k, v in after_fe.items() if "running" in k or "num_batches" in k},
        **{f"after_cl/{k}": v for k, v in after_cl.items() if "running" in k or "num_batches" in k},
        **{f"grad/{k}": v for k, v in grads.items()}, **{f"out/{k}": v for k, v in out.items()})

    # ---- second small model: even-free odd Kmax=13 bank, univariate, longer series -------------
    lpl_uni = ref_sb.generate_layer_parameter_list(1, 13, [8 * 8 * 1, 5 * 8 * 16 + 2 * 16 * 8], 1)
    g = torch.Generator().manual_seed(11)
    x = torch.randn(4, 1, 160, generator=g)
    y = torch.randint(0, 3, (4,), generator=g)
    init_fe, init_cl, after_fe, after_cl, grads, out, lpl_cls = run_pair(ref_os, lpl_uni, 3, x, y, seed=1)
    tables["uni"] = {"lpl_ext": lpl_uni, "lpl_cls": lpl_cls, "n_class": 3, "seed": 1}
    np.savez_compressed(
        os.path.join(OUT, "uni_pair.npz"),
        **{f"init_fe/{k}": v for k, v in init_fe.items()}, **{f"init_cl/{k}": v for k, v in init_cl.items()},
        **{f"after_fe/{k}": v for k, v in after_fe.items() if "running" in k or "num_batches" in k},
        **{f"after_cl/{k}": v for k, v in after_cl.items() if "running" in k or "num_batches" in k},
        **{f"grad/{k}": v for k, v in grads.items()}, **{f"out/{k}": v for k, v in out.items()})

    # ---- cfg1 at full width, seed-only (weights are reproduced by init replay) ----------------
    C, L, K, B = 1, 128, 5, 16
    lpl = tables["layer_lists"]["C1_L128"]
    lpl = [[tuple(t) for t in layer] for layer in lpl]
    x, y = O.synthetic_batch(B, C, L, K, domain_id=0)
    init_fe, init_cl, after_fe, after_cl, grads, out, lpl_cls = run_pair(ref_os, lpl, K, x, y, seed=0)
    tables["state_dict"]["cfg1_fe"] = {k: list(v.shape) for k, v in init_fe.items()}
    tables["state_dict"]["cfg1_cl"] = {k: list(v.shape) for k, v in init_cl.items()}
    checks = {f"fe.{k}": [float(v.astype(np.float64).sum()), float(np.abs(v.astype(np.float64)).sum())]
              for k, v in init_fe.items()}
    checks.update({f"cl.{k}": [float(v.astype(np.float64).sum()), float(np.abs(v.astype(np.float64)).sum())]
                   for k, v in init_cl.items()})
    tables["cfg1"] = {"param_checksums": checks, "seed": 0, "B": B, "C": C, "L": L, "n_class": K,
                      "loss": float(out["loss"]), "argmax": np.argmax(out["logits"], axis=1).tolist(),
                      "argmax_eval": np.argmax(out["logits_eval"], axis=1).tolist()}
    masks = {f"cl.net.{i}.conv1d.weight": O.build_mask(lpl_cls[i]) for i in range(3)}
    masks.update({f"fe.net_1.net.net.{i}.conv1d.weight": O.build_mask(lpl[i]) for i in range(3)})
    gnorm = {k: float(np.linalg.norm((v * masks[k]) if k in masks else v)) for k, v in grads.items()}
    tables["cfg1"]["masked_grad_norms"] = gnorm
    np.savez_compressed(
        os.path.join(OUT, "cfg1_seeded.npz"),
        logits=out["logits"], pooled=out["pooled"], logits_eval=out["logits_eval"],
        feat_b0=out["feat"][0], feat_eval_b0=out["feat_eval"][0], dfeat_b0=out["dfeat"][0],
        **{f"grad/{k}": v for k, v in grads.items() if v.ndim == 1 or "hidden" in k},
        **{"grad/fe.net_1.net.net.0.conv1d.weight": grads["fe.net_1.net.net.0.conv1d.weight"]},
        **{f"after_fe/{k}": v for k, v in after_fe.items() if "running" in k},
        **{f"after_cl/{k}": v for k, v in after_cl.items() if "running" in k})

    # ---- C-DAN consumer (C_DAN.py + widgets.AdversarialNetworkforCDAN), small shapes -----------
    tables["cdan_small"] = make_cdan_golden()
    make_voting_golden()
    make_gradnorm_golden()

    with open(os.path.join(OUT, "tables.json"), "w") as f:
        json.dump(tables, f, indent=1, sort_keys=True)
    print("golden vectors written to", OUT)
    for fn in sorted(os.listdir(OUT)):
        print(f"  {fn}: {os.path.getsize(os.path.join(OUT, fn))} bytes")


if __name__ == "__main__":
    main()
