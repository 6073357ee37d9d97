"""Autograd operators of the hot path, built on the C-ABI kernels (ops.py).

* ``os_stack``     -- a chain of OS layers (masked multi-size Conv1d -> BatchNorm1d -> ReLU), optionally with
                      the 1x1 shortcut branch and the final ``relu(shortcut + block)`` of ``Res_OS_layer``
                      (reference: OS_CNN/OS_CNN.py:46-77,117-139,155-180).  One autograd node for the whole
                      chain: activations stay in the c8 device layout between layers, in the engine's operand
                      dtype, and only the module boundary is NCL fp32.
* ``adain``        -- per-(b, c) mean/std swap (spec SURVEY.md 8c; site train_and_test.py:552-561).
* ``gram_style_loss`` -- mean((G(a) - G(b))^2) with G = x x^T / (C L).

The backward of ``os_stack`` never mutates what the forward saved, so ``retain_graph=True``, repeated
``backward()`` calls and ``torch.autograd.grad`` on sub-losses (train_and_test.py:678-690,741) work.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import torch

from . import _lib as L
from . import ops

ADAIN_EPS = 1e-5


@dataclass
class LayerSpec:
    """Static description of one conv+BN(+ReLU) layer of a stack."""
    geom: ops.BankGeometry
    relu: bool
    training: bool          # BN mode of this layer for this call
    momentum: float         # effective momentum for the running statistics (0 disables the update)
    eps: float
    running_mean: Optional[torch.Tensor]
    running_var: Optional[torch.Tensor]
    zero_masked: bool = True
    dense_wgrad: bool = False   # weight gradient over ALL taps (the reference's unmasked ``.grad``) instead of grad*mask


@dataclass
class StackSpec:
    layers: List[LayerSpec]
    shortcut: Optional[LayerSpec]     # 1x1 conv + BN branch added before the final ReLU
    final_relu: bool                   # relu(shortcut + block) of Res_OS_layer
    engine: int                        # conv (forward + dgrad) engine
    wgrad_engine: int
    op_dtype: int                      # operand dtype of the stack (bf16 for the tensor-core engine)
    direct_grads: bool = False         # backward adds parameter gradients straight into the existing .grad buffers
    pooled: bool = False               # return mean over L, [B, C] (the classifier's AdaptiveAvgPool1d(1) + squeeze)


class _Saved:
    """Per-call buffers kept for backward (all read-only after forward)."""
    __slots__ = ("x_ops", "ys", "coeffs", "wd", "y_r", "co_r", "wd_r", "B", "Ln", "fused")


_DIRECT_GRADS = False
FUSED_PATH = True          # debugging switch: False runs the tcgen05 engine through the unfused BatchNorm kernels


def set_direct_grads(on: bool) -> None:
    """When on, the backward of ``os_stack`` adds the gradients of parameters that already own a ``.grad`` buffer
    (e.g. views of a flat data-parallel bucket) in place inside the kernels and reports ``None`` to autograd for them:
    the wgrad / BatchNorm-backward kernels write the collective's operand directly and no ``AccumulateGrad`` add
    kernels run.  Off (the default) keeps plain autograd semantics (``torch.autograd.grad``, hooks)."""
    global _DIRECT_GRADS
    _DIRECT_GRADS = bool(on)


def direct_grads() -> bool:
    return _DIRECT_GRADS


_DENSE_WGRAD = False


class dense_wgrad:
    """Context manager / switch: while on, every ``os_stack`` backward returns the weight gradient of the *unmasked* big
    Conv1d -- non-zero on the masked taps, exactly what the reference's autograd produces (OS_CNN.py:68-71) --
    instead of ``grad * mask``.  GradNorm needs it: its per-loss norms ``torch.norm(w_i * g)`` run over the whole
    ``.grad`` tensors of ``return_last_layer().parameters()`` (train_and_test.py:683-690).  Read at *backward* time, so
    ``with dense_wgrad(): torch.autograd.grad(loss_i, block.parameters(), retain_graph=True)`` works on a graph that was
    built outside the context.  Costs the dense/live ratio of the bank (2.2x for cfg2's 72->228 layer) in the wgrad kernel."""

    def __init__(self, on: bool = True):
        self.on = bool(on)

    def __enter__(self):
        global _DENSE_WGRAD
        self.prev, _DENSE_WGRAD = _DENSE_WGRAD, self.on
        return self

    def __exit__(self, *exc):
        global _DENSE_WGRAD
        _DENSE_WGRAD = self.prev
        return False


def _wgrad_geom(ls: "LayerSpec"):
    return ls.geom.dense_twin() if (ls.dense_wgrad or _DENSE_WGRAD) else ls.geom


_AUX_STREAMS = {}
WGRAD_ON_AUX_STREAM = False     # optional: run the weight gradient (off the backward's critical path) on an auxiliary
                                # stream beside dgrad; measured no gain at cfg2 (the two branch streams already fill the
                                # two CTA slots per SM), so it is off by default


def _aux_stream(cur: "torch.cuda.Stream") -> "torch.cuda.Stream":
    """One auxiliary stream per calling stream (the two branches of a step each get their own)."""
    key = (cur.device.index, cur.cuda_stream)
    st = _AUX_STREAMS.get(key)
    if st is None:
        st = _AUX_STREAMS[key] = torch.cuda.Stream(device=cur.device)
    return st


def _fused_forward(spec: StackSpec, sv: "_Saved", x: torch.Tensor, params, need_dgrad_first: bool):
    """tcgen05 engine: conv (+ per-CTA BatchNorm statistics in its epilogue) -> fused merge + apply, one pack launch."""
    eng, dt = spec.engine, spec.op_dtype
    B, _, Ln = x.shape
    nl = len(spec.layers)
    ncta = ops.n_conv_ctas(B, Ln)
    dev = x.device
    h = ops.ncl_to_c8(x, dt)
    jobs = []
    for i, ls in enumerate(spec.layers):
        jobs.append((ls.geom, params[4 * i], ls.zero_masked, i > 0 or need_dgrad_first))
    if spec.shortcut is not None:
        jobs.append((spec.shortcut.geom, params[4 * nl], False, need_dgrad_first))
    with torch.no_grad():
        packs = ops.pack_weights_multi(jobs, dt)

    def branch(ls, y, part, gamma, beta):
        coef = torch.empty((4, ls.geom.cout_p), device=dev, dtype=torch.float32)
        track = ls.training and ls.momentum > 0
        rm = ls.running_mean if (track or not ls.training) else None
        rv = ls.running_var if (track or not ls.training) else None
        return ops.BNLayerFwd(y, part, gamma, beta, rm, rv, ls.momentum if ls.training else 0.0, ls.eps, coef)

    out = None
    for i, ls in enumerate(spec.layers):
        W, bias, gamma, beta = params[4 * i: 4 * i + 4]
        g = ls.geom
        wf, wd = packs[i]
        sv.wd.append(wd)
        part = torch.empty((ncta, g.cout_p, 2), device=dev, dtype=torch.float32) if ls.training else None
        y = ops.osconv(eng, L.DIR_FWD, g, h, wf, bias, stat_partial=part)
        br = branch(ls, y, part, gamma, beta)
        sv.x_ops.append(h)
        sv.ys.append(y)
        sv.coeffs.append(br.coef)
        if i < nl - 1:
            h = ops.bn_apply_fused(br, None, g.cout, ls.relu, L.OUT_C8_BF16)
        elif spec.shortcut is None:
            out = ops.bn_apply_fused(br, None, g.cout, ls.relu, L.OUT_POOLED if spec.pooled else L.OUT_NCL_F32)
        else:
            sc = spec.shortcut
            Wr, br_, gr, betar = params[4 * nl: 4 * nl + 4]
            wfr, sv.wd_r = packs[nl]
            part_r = torch.empty((ncta, sc.geom.cout_p, 2), device=dev, dtype=torch.float32) if sc.training else None
            y_r = ops.osconv(eng, L.DIR_FWD, sc.geom, sv.x_ops[0], wfr, br_, stat_partial=part_r)
            brr = branch(sc, y_r, part_r, gr, betar)
            sv.y_r, sv.co_r = y_r, brr.coef
            out = ops.bn_apply_fused(br, brr, g.cout, spec.final_relu, L.OUT_NCL_F32)
    return out


def _inference_forward(spec: StackSpec, x: torch.Tensor, params):
    """Forward-only pass with every BatchNorm in eval mode (utils.eval_*, multi_source_voting): the
    running-statistics affine, the ReLU, the shortcut add and the pooling are the convolution's epilogue, so a stack is
    ncl_to_c8 + one pack launch + one convolution per layer (which derives the BatchNorm coefficients in its prologue); no
    pre-BN tensor reaches HBM."""
    eng, dt = spec.engine, spec.op_dtype
    B, _, Ln = x.shape
    nl = len(spec.layers)
    h = x0 = ops.ncl_to_c8(x, dt)
    jobs = [(ls.geom, params[4 * i], ls.zero_masked, False) for i, ls in enumerate(spec.layers)]
    if spec.shortcut is not None:
        jobs.append((spec.shortcut.geom, params[4 * nl], False, False))
    packs = ops.pack_weights_multi(jobs, dt)

    def coeffs(ls, k):
        return (params[k + 2], params[k + 3], ls.running_mean, ls.running_var, ls.eps)

    out = None
    for i, ls in enumerate(spec.layers):
        co = coeffs(ls, 4 * i)
        bias = params[4 * i + 1]
        wf = packs[i][0]
        if i < nl - 1:
            h = ops.osconv(eng, L.DIR_FWD, ls.geom, h, wf, bias, affine=(co, ls.relu, L.OUT_C8_BF16, None))
        elif spec.shortcut is None:
            if spec.pooled and Ln > 128:      # the pooled epilogue owns one sample per CTA: longer series take two kernels
                y = ops.osconv(eng, L.DIR_FWD, ls.geom, h, wf, bias)
                out = ops.bn_apply_fused(ops.BNLayerFwd(y, None, params[4 * i + 2], params[4 * i + 3], ls.running_mean,
                                                        ls.running_var, 0.0, ls.eps,
                                                        torch.empty((4, ls.geom.cout_p), device=x.device, dtype=torch.float32)),
                                         None, ls.geom.cout, ls.relu, L.OUT_POOLED)
            else:
                out = ops.osconv(eng, L.DIR_FWD, ls.geom, h, wf, bias,
                                 affine=(co, ls.relu, L.OUT_POOLED if spec.pooled else L.OUT_NCL_F32, None))
        else:
            sc = spec.shortcut
            co_r = coeffs(sc, 4 * nl)
            r = ops.osconv(eng, L.DIR_FWD, sc.geom, x0, packs[nl][0], params[4 * nl + 1],
                           affine=(co_r, False, L.OUT_C8_F32, None))
            out = ops.osconv(eng, L.DIR_FWD, ls.geom, h, wf, bias, affine=(co, spec.final_relu, L.OUT_NCL_F32, r))
    return out


def _fused_backward(spec: StackSpec, sv: "_Saved", params, dout: torch.Tensor, x_requires_grad: bool):
    eng, dt = spec.engine, spec.op_dtype
    nl = len(spec.layers)
    B, Ln = sv.B, sv.Ln
    dev = dout.device
    ncta = ops.n_conv_ctas(B, Ln)
    direct = spec.direct_grads and all(p.grad is not None and p.grad.is_contiguous() for p in params)
    grads: List[Optional[torch.Tensor]] = [None] * len(params)
    zero_bias = None

    def param_grad_targets(k, C, training):
        """(dgamma, dbeta, dbias) buffers for parameter group k (index of W in params)."""
        nonlocal zero_bias
        if direct:
            return params[k + 2].grad, params[k + 3].grad, (None if training else params[k + 1].grad)
        dg = torch.empty(C, device=dev, dtype=torch.float32)
        db = torch.empty(C, device=dev, dtype=torch.float32)
        if training:
            if zero_bias is None:       # d(conv bias) behind a train-mode BN is identically zero
                zero_bias = torch.zeros(max(p.numel() for p in params[1::4]), device=dev, dtype=torch.float32)
            dbias = zero_bias[:C]
            grads[k + 1] = dbias
            dbias_buf = None
        else:
            dbias_buf = torch.empty(C, device=dev, dtype=torch.float32)
            grads[k + 1] = dbias_buf
        grads[k + 2], grads[k + 3] = dg, db
        return dg, db, dbias_buf

    cur_stream = torch.cuda.current_stream()
    aux = _aux_stream(cur_stream) if WGRAD_ON_AUX_STREAM else None

    def wgrad(k, g, dy, x_op):
        # dW depends on (dy, x) only and nothing in this backward depends on it: it runs on the auxiliary stream next to
        # the dgrad -> BatchNorm-backward chain of the layers below (joined before this function returns)
        if aux is not None:
            aux.wait_stream(cur_stream)
            dy.record_stream(aux)
            x_op.record_stream(aux)
        with torch.cuda.stream(aux if aux is not None else cur_stream):
            if direct:
                ops.oswgrad(spec.wgrad_engine, g, dy, x_op, out=params[k].grad, accumulate=True)
            else:
                grads[k] = ops.oswgrad(spec.wgrad_engine, g, dy, x_op)
                if aux is not None:
                    grads[k].record_stream(cur_stream)

    last = spec.layers[-1]
    S_top = ops.bn_fused_splits(B, last.geom.cout, Ln)
    red = torch.empty((S_top, last.geom.cout_p, 2), device=dev, dtype=torch.float32)
    a_top = ops.BNLayerBwd(sv.ys[-1], sv.coeffs[-1], params[4 * (nl - 1) + 2], last.training, red)
    dx_short = None
    if spec.shortcut is not None:
        sc = spec.shortcut
        red_r = torch.empty((S_top, sc.geom.cout_p, 2), device=dev, dtype=torch.float32)
        b_top = ops.BNLayerBwd(sv.y_r, sv.co_r, params[4 * nl + 2], sc.training, red_r)
        d = ops.bn_bwd_top(dout, a_top, b_top, spec.final_relu)
        b_top.dgamma, b_top.dbeta, b_top.dbias = param_grad_targets(4 * nl, sc.geom.cout, sc.training)
        dy_r = ops.bn_bwd_apply_fused(d, b_top, S_top, sc.geom.cout, dt, direct)
        wgrad(4 * nl, _wgrad_geom(sc), dy_r, sv.x_ops[0])
        if x_requires_grad:
            dx_short = ops.osconv(eng, L.DIR_DGRAD, sc.geom, dy_r, sv.wd_r, None)
    elif spec.pooled:
        d = ops.bn_bwd_top_pooled(dout, a_top, last.relu, Ln)
    else:
        d = ops.bn_bwd_top(dout, a_top, None, last.relu)
    cur, n_part = a_top, S_top
    dz = None
    for i in range(nl - 1, -1, -1):
        ls = spec.layers[i]
        g = ls.geom
        cur.dgamma, cur.dbeta, cur.dbias = param_grad_targets(4 * i, g.cout, ls.training)
        dy = ops.bn_bwd_apply_fused(d, cur, n_part, g.cout, dt, direct)
        wgrad(4 * i, _wgrad_geom(ls), dy, sv.x_ops[i])
        if i > 0:
            below = spec.layers[i - 1]
            co = sv.coeffs[i - 1]
            red_b = torch.empty((ncta, below.geom.cout_p, 2), device=dev, dtype=torch.float32)
            mask = (sv.ys[i - 1], co[2] if below.relu else None, co[3] if below.relu else None, co[0], co[1])
            d = ops.osconv(eng, L.DIR_DGRAD, g, dy, sv.wd[i], None, mask=mask, red_partial=red_b)
            cur = ops.BNLayerBwd(sv.ys[i - 1], co, params[4 * (i - 1) + 2], below.training, red_b)
            n_part = ncta
        elif x_requires_grad:
            dz = ops.osconv(eng, L.DIR_DGRAD, g, dy, sv.wd[0], None)
    dx = None
    if x_requires_grad:
        if dx_short is not None:
            dz = dz + dx_short
        dx = ops.c8_to_ncl(dz, spec.layers[0].geom.cin)
    if aux is not None:
        cur_stream.wait_stream(aux)
    return dx, grads


class OSStackFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, spec: StackSpec, x: torch.Tensor, *params: torch.Tensor):
        eng = spec.engine
        dt = spec.op_dtype
        ops._req(x, name="input")
        B, C, Ln = x.shape
        if C != spec.layers[0].geom.cin:
            raise RuntimeError(f"input has {C} channels, the first OS layer expects {spec.layers[0].geom.cin}")
        sv = _Saved()
        sv.B, sv.Ln = B, Ln
        sv.x_ops, sv.ys, sv.coeffs, sv.wd = [], [], [], []
        sv.fused = FUSED_PATH and eng == L.ENGINE_TCGEN05 and dt == L.TSC_BF16
        nl = len(spec.layers)
        if sv.fused:
            out = _fused_forward(spec, sv, x, params, x.requires_grad)
            ctx.spec, ctx.sv = spec, sv
            ctx.save_for_backward(*params)
            ctx.x_requires_grad = x.requires_grad
            return out
        h = ops.ncl_to_c8(x, dt)
        x_op = h
        out = None
        need_dgrad_first = x.requires_grad
        for i, ls in enumerate(spec.layers):
            W, bias, gamma, beta = params[4 * i: 4 * i + 4]
            g = ls.geom
            with torch.no_grad():
                wf, wd = ops.pack_weights_pair(g, W, dt, ls.zero_masked, i > 0 or need_dgrad_first)
                sv.wd.append(wd)
            y = ops.osconv(eng, L.DIR_FWD, g, h, wf, bias)
            if ls.training:
                co = ops.bn_stats(y, g.cout, gamma, beta, ls.running_mean if ls.momentum > 0 else None,
                                  ls.running_var if ls.momentum > 0 else None, ls.momentum, ls.eps)
            else:
                co = ops.bn_eval_coeffs(g.cout, gamma, beta, ls.running_mean, ls.running_var, ls.eps)
            sv.x_ops.append(h)
            sv.ys.append(y)
            sv.coeffs.append(co)
            if i < nl - 1:
                h = ops.bn_apply(y, co, g.cout, ls.relu, L.OUT_C8_BF16 if dt == L.TSC_BF16 else L.OUT_C8_F32)
            elif spec.shortcut is None:
                out = ops.bn_apply(y, co, g.cout, ls.relu, L.OUT_NCL_F32)
            else:
                sc = spec.shortcut
                Wr, br, gr, betar = params[4 * nl: 4 * nl + 4]
                with torch.no_grad():
                    wfr, sv.wd_r = ops.pack_weights_pair(sc.geom, Wr, dt, False, need_dgrad_first)
                y_r = ops.osconv(eng, L.DIR_FWD, sc.geom, x_op, wfr, br)
                if sc.training:
                    co_r = ops.bn_stats(y_r, sc.geom.cout, gr, betar, sc.running_mean if sc.momentum > 0 else None,
                                        sc.running_var if sc.momentum > 0 else None, sc.momentum, sc.eps)
                else:
                    co_r = ops.bn_eval_coeffs(sc.geom.cout, gr, betar, sc.running_mean, sc.running_var, sc.eps)
                sv.y_r, sv.co_r = y_r, co_r
                out = ops.bn_apply(y, co, g.cout, spec.final_relu, L.OUT_NCL_F32, y2=y_r, co2=co_r)
        ctx.spec, ctx.sv = spec, sv
        ctx.save_for_backward(*params)
        ctx.x_requires_grad = need_dgrad_first
        return out

    @staticmethod
    def backward(ctx, dout: torch.Tensor):
        spec, sv = ctx.spec, ctx.sv
        params = ctx.saved_tensors
        eng = spec.engine
        dt = spec.op_dtype
        nl = len(spec.layers)
        dout = dout.contiguous().float()
        if sv.fused:
            dx, grads = _fused_backward(spec, sv, params, dout, ctx.x_requires_grad)
            return (None, dx, *grads)
        dz = ops.ncl_to_c8(dout, L.TSC_F32)
        grads: List[Optional[torch.Tensor]] = [None] * len(params)
        dx_short = None
        # d(conv bias) behind a train-mode BN is identically zero: one zero buffer serves every layer
        zero_bias = torch.zeros(max(p.numel() for p in params[1::4]), device=dout.device, dtype=torch.float32)
        # masks of the top of the stack
        last = spec.layers[-1]
        if spec.shortcut is not None:
            top_mask1 = (sv.ys[-1], sv.coeffs[-1]) if spec.final_relu else None
            top_mask2 = (sv.y_r, sv.co_r) if spec.final_relu else None
            sc = spec.shortcut
            Wr, br, gr, betar = params[4 * nl: 4 * nl + 4]
            s1, s2 = ops.bn_bwd_reduce(dz, sv.y_r, sv.co_r, sc.geom.cout, top_mask1, top_mask2)
            dy_r = ops.bn_bwd_apply(dz, sv.y_r, sv.co_r, gr, s1, s2, sc.training, sc.geom.cout, dt, top_mask1, top_mask2)
            C = sc.geom.cout
            grads[4 * nl + 2] = s2[:C]
            grads[4 * nl + 3] = s1[:C]
            grads[4 * nl + 1] = zero_bias[:C] if sc.training else (gr * sv.co_r.invstd[:C] * s1[:C])
            grads[4 * nl + 0] = ops.oswgrad(spec.wgrad_engine, _wgrad_geom(sc), dy_r, sv.x_ops[0])
            if ctx.x_requires_grad:
                dx_short = ops.osconv(eng, L.DIR_DGRAD, sc.geom, dy_r, sv.wd_r, None)
        else:
            top_mask1 = (sv.ys[-1], sv.coeffs[-1]) if last.relu else None
            top_mask2 = None
        for i in range(nl - 1, -1, -1):
            ls = spec.layers[i]
            g = ls.geom
            W, bias, gamma, beta = params[4 * i: 4 * i + 4]
            y, co = sv.ys[i], sv.coeffs[i]
            if i == nl - 1:
                m1, m2 = top_mask1, top_mask2
            else:
                m1, m2 = ((y, co) if ls.relu else None), None
            s1, s2 = ops.bn_bwd_reduce(dz, y, co, g.cout, m1, m2)
            dy = ops.bn_bwd_apply(dz, y, co, gamma, s1, s2, ls.training, g.cout, dt, m1, m2)
            C = g.cout
            grads[4 * i + 2] = s2[:C]
            grads[4 * i + 3] = s1[:C]
            # d(bias) = sum dY: identically 0 behind a train-mode BN, gamma*invstd*S1 behind an eval-mode BN
            grads[4 * i + 1] = zero_bias[:C] if ls.training else (gamma * co.invstd[:C] * s1[:C])
            grads[4 * i + 0] = ops.oswgrad(spec.wgrad_engine, _wgrad_geom(ls), dy, sv.x_ops[i])
            if i > 0 or ctx.x_requires_grad:
                dz = ops.osconv(eng, L.DIR_DGRAD, g, dy, sv.wd[i], None)
        dx = None
        if ctx.x_requires_grad:
            if dx_short is not None:
                dz = dz + dx_short
            dx = ops.c8_to_ncl(dz, spec.layers[0].geom.cin)
        return (None, dx, *grads)


INFERENCE_PATH = True      # debugging switch: False sends forward-only eval calls through the training kernels


def os_stack(spec: StackSpec, x: torch.Tensor, params: List[torch.Tensor]) -> torch.Tensor:
    no_graph = not torch.is_grad_enabled() or not (x.requires_grad or any(p.requires_grad for p in params))
    if (INFERENCE_PATH and no_graph and FUSED_PATH and spec.engine == L.ENGINE_TCGEN05 and spec.op_dtype == L.TSC_BF16
            and not any(ls.training for ls in spec.layers) and not (spec.shortcut is not None and spec.shortcut.training)):
        ops._req(x, name="input")
        if x.shape[1] != spec.layers[0].geom.cin:
            raise RuntimeError(f"input has {x.shape[1]} channels, the first OS layer expects {spec.layers[0].geom.cin}")
        with torch.no_grad():
            return _inference_forward(spec, x, list(params))
    return OSStackFunction.apply(spec, x, *params)


class AdaINFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, content, style, eps):
        content = content.contiguous()
        style = style.contiguous()
        out, stats = ops.adain_fwd(content, style, eps)
        ctx.save_for_backward(content, style, stats)
        return out

    @staticmethod
    def backward(ctx, dy):
        content, style, stats = ctx.saved_tensors
        dc, ds = ops.adain_bwd(dy.contiguous(), content, style, stats)
        return dc, ds, None


def adain(content: torch.Tensor, style: torch.Tensor, eps: float = ADAIN_EPS) -> torch.Tensor:
    """out = (content - mu_c)/sigma_c * sigma_s + mu_s per (b, c) row over L; sigma = sqrt(var_unbiased + eps).
    content/style: [B, C, L] fp32 CUDA, paired by batch index (truncate to the smaller batch)."""
    if content.shape[0] != style.shape[0]:
        n = min(content.shape[0], style.shape[0])
        content, style = content[:n], style[:n]
    return AdaINFunction.apply(content, style, eps)


class GramStyleLossFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, engine):
        a = a.contiguous()
        b = b.contiguous()
        loss, D = ops.gram_loss_fwd(engine, a, b)
        ctx.save_for_backward(a, b, D)
        ctx.engine = engine
        return loss

    @staticmethod
    def backward(ctx, dloss):
        a, b, D = ctx.saved_tensors
        da, db = ops.gram_loss_bwd(ctx.engine, D, a, b, dloss.contiguous().float())
        return da, db, None


def gram_style_loss(a: torch.Tensor, b: torch.Tensor, engine: Optional[int] = None) -> torch.Tensor:
    """mean over B*C*C of (G(a) - G(b))^2, G(x) = x x^T / (C L); a, b: [B, C, L] fp32 CUDA."""
    if a.shape != b.shape:
        raise RuntimeError(f"gram_style_loss needs equal shapes, got {tuple(a.shape)} and {tuple(b.shape)}")
    return GramStyleLossFunction.apply(a, b, ops.get_engine("gram") if engine is None else engine)


class HeadCEFunction(torch.autograd.Function):
    """Linear head + softmax cross-entropy (ops.head_ce_fwd / head_ce_bwd).  Returns (logits, loss); ``labels`` None gives
    logits only (loss is a zero scalar that must not be used).  With ``direct`` the parameter gradients are added into the
    existing ``.grad`` buffers inside the kernel (flat data-parallel bucket) and reported as None."""

    @staticmethod
    def forward(ctx, pooled, W, bias, labels, direct):
        pooled = pooled.contiguous()
        logits, prob, loss = ops.head_ce_fwd(pooled, W, bias, labels)
        ctx.save_for_backward(pooled, W, bias, prob, labels if labels is not None else torch.empty(0, device=pooled.device))
        ctx.has_labels = labels is not None
        ctx.direct = bool(direct)
        ctx.need_dpooled = pooled.requires_grad
        ctx.set_materialize_grads(False)          # an unused output arrives as None, not as a zero-filled tensor
        return logits, loss

    @staticmethod
    def backward(ctx, dlogits, dloss):
        pooled, W, bias, prob, labels = ctx.saved_tensors
        use_loss = ctx.has_labels and dloss is not None
        dl = dlogits.contiguous().float() if dlogits is not None else None
        if dl is None and not use_loss:
            return None, None, None, None, None
        direct = ctx.direct and W.grad is not None and bias.grad is not None and W.grad.is_contiguous()
        dpooled, dW, db = ops.head_ce_bwd(dloss.contiguous().float() if use_loss else None, dl, prob,
                                          labels if use_loss else None, pooled, W, ctx.need_dpooled,
                                          W.grad if direct else None, bias.grad if direct else None)
        if direct:
            return dpooled, None, None, None, None
        return dpooled, dW, db, None, None


def head_cross_entropy(pooled: torch.Tensor, W: torch.Tensor, bias: torch.Tensor, labels: Optional[torch.Tensor]):
    """(logits, mean cross-entropy) of ``Linear(W, bias)`` over pooled features [B, C]; one kernel forward, one backward."""
    return HeadCEFunction.apply(pooled, W, bias, labels, _DIRECT_GRADS)


class _WeightedSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, weights, *terms):
        ctx.weights = [float(w) for w in weights]
        return ops.weighted_scalar_sum([t.contiguous().float() for t in terms], ctx.weights)

    @staticmethod
    def backward(ctx, dout):
        # terms with weight 1 share the incoming gradient tensor (no launch)
        return (None, *[dout if w == 1.0 else dout * w for w in ctx.weights])


def weighted_loss_sum(terms, weights) -> torch.Tensor:
    """sum_i weights[i] * terms[i] of scalar losses (the total loss of a step) in one launch."""
    return _WeightedSum.apply(list(weights), *terms)
