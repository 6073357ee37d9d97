"""Tensor-level wrappers over the C-ABI (one Python function per entry point of include/tsc_b200.h).

PyTorch here is plumbing only: it owns device memory and the current stream.  Every function takes
CUDA tensors, allocates its outputs with torch, and passes raw pointers + the current stream to
libtsc_b200.so.  Nothing in this module computes on the host and nothing falls back to torch ops.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import torch

from . import _lib as L

# Engine per kernel family plus the operand dtype of the OS stack.
#   "tcgen05"   : conv / wgrad / gram on the tensor cores, bf16 operands, fp32 accumulation   (<= 1e-2)
#   "simt"      : fp32 CUDA-core kernels, fp32 operands                                        (<= 1e-5)
#   "simt_bf16" : the CUDA-core kernels fed the SAME bf16 operands as the tensor-core engine -- not a
#                 product mode: it is the bit-faithful checker of the tcgen05 kernels used by tests.
_TC_READY = ("conv", "wgrad", "gram")   # families whose tcgen05 kernel exists
_CFG = {"conv": L.ENGINE_TCGEN05, "wgrad": L.ENGINE_TCGEN05, "gram": L.ENGINE_TCGEN05, "op_dtype": L.TSC_BF16,
        "name": "tcgen05"}


def set_engine(name: str) -> None:
    if name in ("tcgen05", "bf16"):
        for fam in ("conv", "wgrad", "gram"):
            _CFG[fam] = L.ENGINE_TCGEN05 if fam in _TC_READY else L.ENGINE_SIMT
        _CFG["op_dtype"], _CFG["name"] = L.TSC_BF16, "tcgen05"
    elif name in ("simt", "fp32"):
        for fam in ("conv", "wgrad", "gram"):
            _CFG[fam] = L.ENGINE_SIMT
        _CFG["op_dtype"], _CFG["name"] = L.TSC_F32, "simt"
    elif name == "simt_bf16":
        for fam in ("conv", "wgrad", "gram"):
            _CFG[fam] = L.ENGINE_SIMT
        _CFG["op_dtype"], _CFG["name"] = L.TSC_BF16, "simt_bf16"
    else:
        raise KeyError(name)


def get_engine(family: str = "conv") -> int:
    return _CFG[family]


def engine_name() -> str:
    return _CFG["name"]


def op_dtype() -> int:
    return _CFG["op_dtype"]


def torch_dtype(dt: int):
    return torch.bfloat16 if dt == L.TSC_BF16 else torch.float32


def pad16(c: int) -> int:
    return (c + 15) & ~15


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def _req(t: torch.Tensor, dtype=torch.float32, name="tensor"):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: the tsc_b200 operators have no CPU fallback")
    if t.dtype != dtype:
        raise RuntimeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    return t


# --------------------------------------------------------------------------------------------------
# kernel-bank geometry (host integers; mirrors OS_CNN/OS_CNN.py:9-12,23-43,59 of the reference)
# --------------------------------------------------------------------------------------------------
@dataclass
class BankGeometry:
    cin: int
    cout: int
    kmax: int
    s_of_tap: List[int]
    lo: List[int] = field(default_factory=list)      # per out channel: first live tap
    hi: List[int] = field(default_factory=list)      # per out channel: one past the last live tap

    def __post_init__(self):
        if not (1 <= self.kmax <= L.MAX_TAPS):
            raise ValueError(f"kernel size {self.kmax} outside [1, {L.MAX_TAPS}]")
        if not (1 <= self.cin <= L.MAX_CHANNELS_WIDE and 1 <= self.cout <= L.MAX_CHANNELS_WIDE):
            raise ValueError(f"channel counts ({self.cin}, {self.cout}) outside [1, {L.MAX_CHANNELS_WIDE}]")
        self.cin_p, self.cout_p = pad16(self.cin), pad16(self.cout)
        # wider than one TMEM accumulator tile (the reference's recipe for series shorter than ~80 samples): such a layer --
        # and with it its whole stack -- runs on the fp32 CUDA-core engine
        self.wide = self.cin_p > L.MAX_CHANNELS or self.cout_p > L.MAX_CHANNELS
        self.s_arr = L.int_array(self.s_of_tap)
        self.pad_l, self.pad_r = (self.kmax - 1) // 2, self.kmax // 2
        self._packed_bytes = {}
        self._plans = {}

    def packed_bytes(self, direction: int, dt: int) -> int:
        key = (direction, dt)
        if key not in self._packed_bytes:
            n = L.load().tsc_packed_weight_bytes(direction, dt, self.cin, self.cout, self.kmax, self.s_arr)
            if n == 0:
                raise RuntimeError("tsc_packed_weight_bytes failed: " + L.load().tsc_last_error().decode())
            self._packed_bytes[key] = int(n)
        return self._packed_bytes[key]

    def plan(self, direction: int, device) -> torch.Tensor:
        """Device copy of the tcgen05 issue schedule of this bank (built once per geometry, direction and device)."""
        key = (direction, str(device))
        if key not in self._plans:
            lib = L.load()
            n = int(lib.tsc_osconv_plan_bytes(direction, self.cin, self.cout, self.kmax, self.s_arr))
            if n == 0:
                raise RuntimeError("tsc_osconv_plan_bytes failed: " + lib.tsc_last_error().decode())
            host = torch.empty(n, dtype=torch.uint8)
            L.check(lib.tsc_osconv_plan_build(direction, self.cin, self.cout, self.kmax, self.s_arr,
                                              ctypes.c_void_p(host.data_ptr())), "tsc_osconv_plan_build")
            self._plans[key] = host.to(device)
        return self._plans[key]

    def dense_twin(self) -> "BankGeometry":
        """The same bank with every tap live for every channel: the geometry of the reference's *unmasked* weight gradient
        (``W.grad`` of the big Conv1d is dense, OS_CNN.py:68-71 -- SURVEY F4), used when ``dense_wgrad`` is on."""
        if getattr(self, "_dense", None) is None:
            self._dense = self if all(v == 0 for v in self.s_of_tap) else dense_geometry(self.cin, self.cout, self.kmax)
        return self._dense

    def live_macs_per_position(self) -> int:
        return self.cin * sum(h - l for l, h in zip(self.lo, self.hi))


def mask_interval(k: int, kmax: int):
    """Live taps [left, left+k) of a size-k kernel inside the kmax window (OS_CNN/OS_CNN.py:9-12)."""
    right_zero = -((-(kmax - 1)) // 2) - (-((-(k - 1)) // 2))     # ceil((kmax-1)/2) - ceil((k-1)/2)
    left = kmax - k - right_zero
    return left, left + k


def bank_geometry(layer_parameters: Sequence[Sequence[int]]) -> BankGeometry:
    """Geometry of one OS layer from its [(in_ch, out_ch, kernel), ...] list."""
    kmax = layer_parameters[-1][-1]
    cin = layer_parameters[0][0]
    lo, hi = [], []
    for (_, oc, k) in layer_parameters:
        l, r = mask_interval(k, kmax)
        lo += [l] * oc
        hi += [r] * oc
    cout = len(lo)
    s_of_tap = []
    for t in range(kmax):
        live = [c for c in range(cout) if lo[c] <= t < hi[c]]
        first = live[0] if live else cout
        if live != list(range(first, cout)):
            raise ValueError("kernel bank is not nested (live channels at a tap must form a suffix)")
        s_of_tap.append(first)
    return BankGeometry(cin, cout, kmax, s_of_tap, lo, hi)


def dense_geometry(cin: int, cout: int, kernel: int) -> BankGeometry:
    """A plain Conv1d (every tap live for every channel), e.g. the 1x1 shortcut (OS_CNN.py:155-166)."""
    return BankGeometry(cin, cout, kernel, [0] * kernel, [0] * cout, [kernel] * cout)


# --------------------------------------------------------------------------------------------------
# wrappers
# --------------------------------------------------------------------------------------------------
def ncl_to_c8(x: torch.Tensor, dt: int) -> torch.Tensor:
    _req(x, name="x")
    B, C, Ln = x.shape
    out = torch.empty((B, pad16(C) // 8, Ln, 8), device=x.device, dtype=torch_dtype(dt))
    L.check(L.load().tsc_ncl_to_c8(_ptr(x), _ptr(out), dt, B, C, Ln, _stream()), "tsc_ncl_to_c8")
    return out


def c8_to_ncl(x8: torch.Tensor, C: int) -> torch.Tensor:
    _req(x8, name="x_c8")
    B, _, Ln, _ = x8.shape
    out = torch.empty((B, C, Ln), device=x8.device, dtype=torch.float32)
    L.check(L.load().tsc_c8_to_ncl(_ptr(x8), _ptr(out), B, C, Ln, _stream()), "tsc_c8_to_ncl")
    return out


def pack_weights(g: BankGeometry, W: torch.Tensor, direction: int, dt: int, zero_masked: bool) -> torch.Tensor:
    _req(W, name="weight")
    if tuple(W.shape) != (g.cout, g.cin, g.kmax):
        raise RuntimeError(f"weight shape {tuple(W.shape)} != {(g.cout, g.cin, g.kmax)}")
    nbytes = g.packed_bytes(direction, dt)
    packed = torch.empty(nbytes // (2 if dt == L.TSC_BF16 else 4), device=W.device, dtype=torch_dtype(dt))
    L.check(L.load().tsc_pack_weights(direction, dt, _ptr(W), _ptr(packed), g.cin, g.cout, g.kmax, g.s_arr,
                                      1 if zero_masked else 0, _stream()), "tsc_pack_weights")
    return packed


def pack_weights_pair(g: BankGeometry, W: torch.Tensor, dt: int, zero_masked: bool, want_dgrad: bool):
    """Forward pack, (optionally) dgrad pack and the in-place masking of W in one launch."""
    _req(W, name="weight")
    if tuple(W.shape) != (g.cout, g.cin, g.kmax):
        raise RuntimeError(f"weight shape {tuple(W.shape)} != {(g.cout, g.cin, g.kmax)}")
    esz = 2 if dt == L.TSC_BF16 else 4
    pf = torch.empty(g.packed_bytes(L.DIR_FWD, dt) // esz, device=W.device, dtype=torch_dtype(dt))
    pd = torch.empty(g.packed_bytes(L.DIR_DGRAD, dt) // esz, device=W.device, dtype=torch_dtype(dt)) if want_dgrad else None
    L.check(L.load().tsc_pack_weights_pair(dt, _ptr(W), _ptr(pf), _ptr(pd), g.cin, g.cout, g.kmax, g.s_arr,
                                           1 if zero_masked else 0, _stream()), "tsc_pack_weights_pair")
    return pf, pd


def pack_weights_multi(layers, dt: int):
    """layers: [(geometry, W, zero_masked, want_dgrad)] -> [(packed_fwd, packed_dgrad | None)], one launch per
    L.PACK_MAX_LAYERS layers (forward pack, dgrad pack and the in-place masking of every W)."""
    esz = 2 if dt == L.TSC_BF16 else 4
    out = []
    for i0 in range(0, len(layers), L.PACK_MAX_LAYERS):
        grp = layers[i0:i0 + L.PACK_MAX_LAYERS]
        batch = L.PackBatch()
        batch.n = len(grp)
        for k, (g, W, zero_masked, want_dgrad) in enumerate(grp):
            _req(W, name="weight")
            if tuple(W.shape) != (g.cout, g.cin, g.kmax):
                raise RuntimeError(f"weight shape {tuple(W.shape)} != {(g.cout, g.cin, g.kmax)}")
            pf = torch.empty(g.packed_bytes(L.DIR_FWD, dt) // esz, device=W.device, dtype=torch_dtype(dt))
            pd = (torch.empty(g.packed_bytes(L.DIR_DGRAD, dt) // esz, device=W.device, dtype=torch_dtype(dt))
                  if want_dgrad else None)
            ly = batch.layer[k]
            ly.W, ly.packed_fwd = W.data_ptr(), pf.data_ptr()
            ly.packed_dgrad = pd.data_ptr() if pd is not None else None
            ly.Cin, ly.Cout, ly.Kmax, ly.zero_masked = g.cin, g.cout, g.kmax, 1 if zero_masked else 0
            for t, sv in enumerate(g.s_of_tap):
                ly.s_of_tap[t] = min(int(sv), 32767)
            out.append((pf, pd))
        L.check(L.load().tsc_pack_weights_multi(dt, ctypes.byref(batch), _stream()), "tsc_pack_weights_multi")
    return out


def rmsprop_step(params: torch.Tensor, grads: torch.Tensor, square_avg: torch.Tensor, group_end, group_lr,
                 alpha: float = 0.99, eps: float = 1e-8, grad_scale: float = 1.0, group_clamp=None):
    """Fused RMSprop over flat fp32 buffers (torch.optim.RMSprop defaults; per-group learning rates; optional WGAN
    weight clipping of chosen groups, ``group_clamp[i] > 0``)."""
    for t, nm in ((params, "params"), (grads, "grads"), (square_avg, "square_avg")):
        _req(t, name=nm)
    n = params.numel()
    ends = (ctypes.c_longlong * len(group_end))(*[int(e) for e in group_end])
    lrs = (ctypes.c_float * len(group_lr))(*[float(x) for x in group_lr])
    clamps = None
    if group_clamp is not None and any(c > 0 for c in group_clamp):
        clamps = (ctypes.c_float * len(group_lr))(*[float(c) for c in group_clamp])
    L.check(L.load().tsc_rmsprop_step_clamped(_ptr(params), _ptr(grads), _ptr(square_avg), n, ends, lrs, clamps,
                                              len(group_end), float(alpha), float(eps), float(grad_scale), _stream()),
            "tsc_rmsprop_step_clamped")


def n_conv_ctas(B: int, Ln: int) -> int:
    """CTAs (= rows of the fused-epilogue partial buffers) of one tcgen05 conv launch."""
    return B * ((Ln + 127) // 128)


def osconv(engine: int, direction: int, g: BankGeometry, x8: torch.Tensor, w_packed: torch.Tensor,
           bias: Optional[torch.Tensor], stat_partial: Optional[torch.Tensor] = None, mask=None,
           red_partial: Optional[torch.Tensor] = None, affine=None) -> torch.Tensor:
    """mask = (y_c8, scale|None, shift|None, mean, invstd) of the layer below (dgrad with red_partial only).
    affine = (coef, relu, out_kind, residual_c8_f32 | None): the inference epilogue of the tcgen05 engine (forward only) --
    returns act(scale * (conv + bias) + shift [+ residual]) in layout ``out_kind``; the pre-BN y is not written.  ``coef`` is
    (scale, shift) with Cout_p entries each, or (gamma, beta, running_mean, running_var, eps): eval-mode BatchNorm, the
    coefficients are then derived in the kernel's prologue."""
    dt = L.TSC_BF16 if x8.dtype == torch.bfloat16 else L.TSC_F32
    _req(x8, torch_dtype(dt), "x_c8")
    _req(w_packed, torch_dtype(dt), "w_packed")
    B, kc, Ln, _ = x8.shape
    cin_side = g.cin_p if direction == L.DIR_FWD else g.cout_p
    cout_side = g.cout_p if direction == L.DIR_FWD else g.cin_p
    if kc * 8 != cin_side:
        raise RuntimeError(f"x_c8 has {kc * 8} padded channels, expected {cin_side}")
    if bias is not None:
        _req(bias, name="bias")
    plan = g.plan(direction, x8.device) if engine == L.ENGINE_TCGEN05 else None
    if affine is not None:
        if direction != L.DIR_FWD or engine != L.ENGINE_TCGEN05 or stat_partial is not None:
            raise RuntimeError("the affine (inference) epilogue exists for the forward tcgen05 convolution only")
        coef, relu, out_kind, residual = affine
        for t in coef[:4]:
            _req(t, name="affine coefficient")
            if t.numel() < (g.cout_p if len(coef) == 2 else g.cout):
                raise RuntimeError("affine coefficient tensor is too short")
        if out_kind == L.OUT_NCL_F32:
            out = torch.empty((B, g.cout, Ln), device=x8.device, dtype=torch.float32)
        elif out_kind == L.OUT_POOLED:
            out = torch.empty((B, g.cout), device=x8.device, dtype=torch.float32)
        else:
            out = torch.empty((B, cout_side // 8, Ln, 8), device=x8.device,
                              dtype=torch.bfloat16 if out_kind == L.OUT_C8_BF16 else torch.float32)
        epi = L.ConvEpilogue()
        if len(coef) == 2:
            epi.affine_scale, epi.affine_shift = coef[0].data_ptr(), coef[1].data_ptr()
        else:
            epi.bn_gamma, epi.bn_beta, epi.bn_mean, epi.bn_var = (t.data_ptr() for t in coef[:4])
            epi.bn_eps = float(coef[4])
        epi.affine_out = out.data_ptr()
        epi.affine_out_kind, epi.affine_relu = int(out_kind), 1 if relu else 0
        if residual is not None:
            _req(residual, name="residual")
            if tuple(residual.shape) != (B, cout_side // 8, Ln, 8):
                raise RuntimeError(f"residual shape {tuple(residual.shape)} != {(B, cout_side // 8, Ln, 8)}")
            epi.residual = residual.data_ptr()
        L.check(L.load().tsc_osconv(engine, direction, _ptr(x8), dt, _ptr(w_packed), _ptr(plan), _ptr(bias), None,
                                    ctypes.byref(epi), B, Ln, g.cin, g.cout, g.kmax, g.s_arr, _stream()), "tsc_osconv")
        return out
    y = torch.empty((B, cout_side // 8, Ln, 8), device=x8.device, dtype=torch.float32)
    epi = None
    if stat_partial is not None or red_partial is not None:
        epi = L.ConvEpilogue()
        if stat_partial is not None:
            epi.stat_partial = stat_partial.data_ptr()
        if red_partial is not None:
            my, msc, msh, mmean, minv = mask
            epi.mask_y, epi.mask_mean, epi.mask_invstd = my.data_ptr(), mmean.data_ptr(), minv.data_ptr()
            if msc is not None:
                epi.mask_scale, epi.mask_shift = msc.data_ptr(), msh.data_ptr()
            epi.red_partial = red_partial.data_ptr()
    L.check(L.load().tsc_osconv(engine, direction, _ptr(x8), dt, _ptr(w_packed), _ptr(plan), _ptr(bias), _ptr(y),
                                ctypes.byref(epi) if epi is not None else None, B, Ln,
                                g.cin, g.cout, g.kmax, g.s_arr, _stream()), "tsc_osconv")
    return y


def oswgrad(engine: int, g: BankGeometry, dy8: torch.Tensor, x8: torch.Tensor,
            out: Optional[torch.Tensor] = None, accumulate: bool = False) -> torch.Tensor:
    """dW of the bank; with ``out`` (fp32 [Cout, Cin, Kmax], e.g. a view of the flat gradient bucket) the result is
    written -- or, with ``accumulate``, added -- in place."""
    dt = L.TSC_BF16 if x8.dtype == torch.bfloat16 else L.TSC_F32
    _req(x8, torch_dtype(dt), "x_c8")
    _req(dy8, torch_dtype(dt), "dy_c8")
    B, _, Ln, _ = x8.shape
    lib = L.load()
    ws = torch.empty(int(lib.tsc_oswgrad_workspace_bytes(engine, B, Ln, g.cin, g.cout, g.kmax)) // 4 + 4,
                     device=x8.device, dtype=torch.float32)
    if out is None:
        if accumulate:
            raise RuntimeError("accumulate needs an output tensor")
        dW = torch.empty((g.cout, g.cin, g.kmax), device=x8.device, dtype=torch.float32)
    else:
        dW = _req(out, name="dW")
        if tuple(dW.shape) != (g.cout, g.cin, g.kmax):
            raise RuntimeError(f"dW shape {tuple(dW.shape)} != {(g.cout, g.cin, g.kmax)}")
    L.check(lib.tsc_oswgrad(engine, _ptr(dy8), _ptr(x8), dt, _ptr(dW), _ptr(ws), 1 if accumulate else 0, B, Ln, g.cin,
                            g.cout, g.kmax, g.s_arr, _stream()), "tsc_oswgrad")
    return dW


@dataclass
class BNCoeffs:
    mean: torch.Tensor
    invstd: torch.Tensor
    scale: torch.Tensor
    shift: torch.Tensor


def _bn_ws(B, C, Ln, device):
    n = int(L.load().tsc_bn_workspace_bytes(B, C, Ln)) // 4 + 4
    return torch.empty(n, device=device, dtype=torch.float32)


def bn_stats(y8: torch.Tensor, C: int, gamma, beta, running_mean, running_var, momentum: float, eps: float) -> BNCoeffs:
    _req(y8, name="y_c8")
    B, cpc, Ln, _ = y8.shape
    cp = cpc * 8
    buf = torch.empty((4, cp), device=y8.device, dtype=torch.float32)
    co = BNCoeffs(buf[0], buf[1], buf[2], buf[3])
    L.check(L.load().tsc_bn_stats(_ptr(y8), _ptr(gamma), _ptr(beta), _ptr(_bn_ws(B, C, Ln, y8.device)), _ptr(co.mean),
                                  _ptr(co.invstd), _ptr(co.scale), _ptr(co.shift), _ptr(running_mean),
                                  _ptr(running_var), float(momentum), float(eps), B, C, Ln, _stream()), "tsc_bn_stats")
    return co


def bn_eval_coeffs(C: int, gamma, beta, running_mean, running_var, eps: float) -> BNCoeffs:
    cp = pad16(C)
    buf = torch.empty((4, cp), device=gamma.device, dtype=torch.float32)
    co = BNCoeffs(buf[0], buf[1], buf[2], buf[3])
    L.check(L.load().tsc_bn_eval_coeffs(_ptr(gamma), _ptr(beta), _ptr(running_mean), _ptr(running_var), float(eps),
                                        _ptr(co.mean), _ptr(co.invstd), _ptr(co.scale), _ptr(co.shift), C, _stream()),
            "tsc_bn_eval_coeffs")
    return co


def bn_apply(y8, co: BNCoeffs, C: int, relu: bool, out_kind: int, y2=None, co2: Optional[BNCoeffs] = None):
    _req(y8, name="y_c8")
    B, cpc, Ln, _ = y8.shape
    if out_kind == L.OUT_NCL_F32:
        out = torch.empty((B, C, Ln), device=y8.device, dtype=torch.float32)
    else:
        out = torch.empty((B, cpc, Ln, 8), device=y8.device,
                          dtype=torch.bfloat16 if out_kind == L.OUT_C8_BF16 else torch.float32)
    L.check(L.load().tsc_bn_apply(_ptr(y8), _ptr(co.scale), _ptr(co.shift), _ptr(y2),
                                  _ptr(co2.scale if co2 else None), _ptr(co2.shift if co2 else None),
                                  1 if relu else 0, _ptr(out), out_kind, B, C, Ln, _stream()), "tsc_bn_apply")
    return out


def bn_bwd_reduce(dz8, y8, co: BNCoeffs, C: int, mask1=None, mask2=None):
    """mask1/mask2: optional (y_c8, BNCoeffs) pre-activation operands the ReLU mask is recomputed from."""
    B, cpc, Ln, _ = y8.shape
    buf = torch.empty((2, cpc * 8), device=y8.device, dtype=torch.float32)
    m1y, m1c = mask1 if mask1 else (None, None)
    m2y, m2c = mask2 if mask2 else (None, None)
    L.check(L.load().tsc_bn_bwd_reduce(_ptr(dz8), _ptr(y8), _ptr(co.mean), _ptr(co.invstd),
                                       _ptr(m1y), _ptr(m1c.scale if m1c else None), _ptr(m1c.shift if m1c else None),
                                       _ptr(m2y), _ptr(m2c.scale if m2c else None), _ptr(m2c.shift if m2c else None),
                                       _ptr(_bn_ws(B, C, Ln, y8.device)), _ptr(buf[0]), _ptr(buf[1]), B, C, Ln,
                                       _stream()), "tsc_bn_bwd_reduce")
    return buf[0], buf[1]


def bn_bwd_apply(dz8, y8, co: BNCoeffs, gamma, s1, s2, training: bool, C: int, dt: int, mask1=None, mask2=None):
    B, cpc, Ln, _ = y8.shape
    dy = torch.empty((B, cpc, Ln, 8), device=y8.device, dtype=torch_dtype(dt))
    m1y, m1c = mask1 if mask1 else (None, None)
    m2y, m2c = mask2 if mask2 else (None, None)
    L.check(L.load().tsc_bn_bwd_apply(_ptr(dz8), _ptr(y8), _ptr(co.mean), _ptr(co.invstd), _ptr(gamma), _ptr(s1), _ptr(s2),
                                      1 if training else 0,
                                      _ptr(m1y), _ptr(m1c.scale if m1c else None), _ptr(m1c.shift if m1c else None),
                                      _ptr(m2y), _ptr(m2c.scale if m2c else None), _ptr(m2c.shift if m2c else None),
                                      _ptr(dy), dt, B, C, Ln, _stream()), "tsc_bn_bwd_apply")
    return dy


# ---- fused BatchNorm path (tcgen05 engine) ----------------------------------------------------------
@dataclass
class BNLayerFwd:
    """One conv+BN branch of a fused apply: y (c8 fp32), the conv's per-CTA statistics (None = eval mode)."""
    y8: torch.Tensor
    stat_partial: Optional[torch.Tensor]
    gamma: torch.Tensor
    beta: torch.Tensor
    running_mean: Optional[torch.Tensor]
    running_var: Optional[torch.Tensor]
    momentum: float
    eps: float
    coef: torch.Tensor          # [4, Cp] out: mean, invstd, scale, shift

    def c(self) -> "L.BNBranch":
        b = L.BNBranch()
        b.y_c8 = self.y8.data_ptr()
        b.stat_partial = self.stat_partial.data_ptr() if self.stat_partial is not None else None
        b.gamma, b.beta = self.gamma.data_ptr(), self.beta.data_ptr()
        b.running_mean = self.running_mean.data_ptr() if self.running_mean is not None else None
        b.running_var = self.running_var.data_ptr() if self.running_var is not None else None
        b.momentum, b.eps = float(self.momentum), float(self.eps)
        b.coef = self.coef.data_ptr()
        return b


def bn_apply_fused(a: BNLayerFwd, b: Optional[BNLayerFwd], C: int, relu: bool, out_kind: int):
    B, cpc, Ln, _ = a.y8.shape
    if out_kind == L.OUT_NCL_F32:
        out = torch.empty((B, C, Ln), device=a.y8.device, dtype=torch.float32)
    elif out_kind == L.OUT_POOLED:
        out = torch.empty((B, C), device=a.y8.device, dtype=torch.float32)
    else:
        out = torch.empty((B, cpc, Ln, 8), device=a.y8.device,
                          dtype=torch.bfloat16 if out_kind == L.OUT_C8_BF16 else torch.float32)
    ca = a.c()
    cb = b.c() if b is not None else None
    L.check(L.load().tsc_bn_apply_fused(ctypes.byref(ca), ctypes.byref(cb) if cb is not None else None,
                                        n_conv_ctas(B, Ln), 1 if relu else 0, _ptr(out), out_kind, B, C, Ln, _stream()),
            "tsc_bn_apply_fused")
    return out


@dataclass
class BNLayerBwd:
    y8: torch.Tensor
    coef: torch.Tensor
    gamma: torch.Tensor
    training: bool
    red_partial: torch.Tensor
    dgamma: Optional[torch.Tensor] = None
    dbeta: Optional[torch.Tensor] = None
    dbias: Optional[torch.Tensor] = None

    def c(self) -> "L.BNBwdBranch":
        b = L.BNBwdBranch()
        b.y_c8, b.coef, b.gamma = self.y8.data_ptr(), self.coef.data_ptr(), self.gamma.data_ptr()
        b.training = 1 if self.training else 0
        b.red_partial = self.red_partial.data_ptr()
        b.dgamma = self.dgamma.data_ptr() if self.dgamma is not None else None
        b.dbeta = self.dbeta.data_ptr() if self.dbeta is not None else None
        b.dbias = self.dbias.data_ptr() if self.dbias is not None else None
        return b


def bn_fused_splits(B: int, C: int, Ln: int) -> int:
    return int(L.load().tsc_bn_fused_splits(B, C, Ln))


def bn_bwd_top(dout: torch.Tensor, a: BNLayerBwd, b: Optional[BNLayerBwd], relu: bool) -> torch.Tensor:
    _req(dout, name="dout")
    B, C, Ln = dout.shape
    d8 = torch.empty_like(a.y8)
    ca = a.c()
    cb = b.c() if b is not None else None
    L.check(L.load().tsc_bn_bwd_top(_ptr(dout), ctypes.byref(ca), ctypes.byref(cb) if cb is not None else None,
                                    1 if relu else 0, _ptr(d8), B, C, Ln, _stream()), "tsc_bn_bwd_top")
    return d8


def bn_bwd_top_pooled(dpooled: torch.Tensor, a: BNLayerBwd, relu: bool, Ln: int) -> torch.Tensor:
    _req(dpooled, name="dpooled")
    B, C = dpooled.shape
    d8 = torch.empty_like(a.y8)
    ca = a.c()
    L.check(L.load().tsc_bn_bwd_top_pooled(_ptr(dpooled), ctypes.byref(ca), 1 if relu else 0, _ptr(d8), B, C, Ln, _stream()),
            "tsc_bn_bwd_top_pooled")
    return d8


def bn_bwd_apply_fused(d8: torch.Tensor, a: BNLayerBwd, n_part: int, C: int, dt: int, accumulate: bool) -> torch.Tensor:
    B, cpc, Ln, _ = d8.shape
    dy = torch.empty((B, cpc, Ln, 8), device=d8.device, dtype=torch_dtype(dt))
    ca = a.c()
    L.check(L.load().tsc_bn_bwd_apply_fused(_ptr(d8), ctypes.byref(ca), n_part, 1 if accumulate else 0, _ptr(dy), dt,
                                            B, C, Ln, _stream()), "tsc_bn_bwd_apply_fused")
    return dy


def rowstats(x: torch.Tensor):
    _req(x, name="x")
    Ln = x.shape[-1]
    R = x.numel() // Ln
    mean = torch.empty(x.shape[:-1], device=x.device, dtype=torch.float32)
    var = torch.empty_like(mean)
    L.check(L.load().tsc_rowstats_welford(_ptr(x), _ptr(mean), _ptr(var), R, Ln, _stream()), "tsc_rowstats_welford")
    return mean, var


def adain_fwd(content, style, eps: float):
    _req(content, name="content"); _req(style, name="style")
    if content.shape != style.shape:
        raise RuntimeError(f"content {tuple(content.shape)} and style {tuple(style.shape)} must have the same shape")
    Ln = content.shape[-1]
    R = content.numel() // Ln
    out = torch.empty_like(content)
    stats = torch.empty((R, 4), device=content.device, dtype=torch.float32)
    L.check(L.load().tsc_adain_fwd(_ptr(content), _ptr(style), _ptr(out), _ptr(stats), float(eps), R, Ln, _stream()),
            "tsc_adain_fwd")
    return out, stats


def adain_bwd(dy, content, style, stats):
    _req(dy, name="dy")
    Ln = content.shape[-1]
    R = content.numel() // Ln
    dc, ds = torch.empty_like(content), torch.empty_like(style)
    L.check(L.load().tsc_adain_bwd(_ptr(dy), _ptr(content), _ptr(style), _ptr(stats), _ptr(dc), _ptr(ds), R, Ln, _stream()),
            "tsc_adain_bwd")
    return dc, ds


def gram_loss_fwd(engine: int, a, s):
    _req(a, name="a"); _req(s, name="s")
    B, C, Ln = a.shape
    D = torch.empty((B, C, C), device=a.device, dtype=torch.float32)
    loss = torch.empty((), device=a.device, dtype=torch.float32)
    ws = torch.empty(int(L.load().tsc_gram_workspace_bytes(B, C, Ln)) // 4 + 4, device=a.device, dtype=torch.float32)
    L.check(L.load().tsc_gram_loss_fwd(engine, _ptr(a), _ptr(s), _ptr(D), _ptr(loss), _ptr(ws), B, C, Ln, _stream()),
            "tsc_gram_loss_fwd")
    return loss, D


def gram_loss_bwd(engine: int, D, a, s, dloss):
    B, C, Ln = a.shape
    _req(dloss, name="dloss")
    da, ds = torch.empty_like(a), torch.empty_like(s)
    L.check(L.load().tsc_gram_loss_bwd(engine, _ptr(D), _ptr(a), _ptr(s), _ptr(dloss), _ptr(da), _ptr(ds), B, C, Ln,
                                       _stream()), "tsc_gram_loss_bwd")
    return da, ds


def cdan_fuse_fwd(y0, logits, r1, scale_div: float):
    """C_DAN.py:20-25,32-37,53-54 for the stacked (target, generated) rows: returns (fusion [M,D], prob [M,K], u [M])."""
    _req(y0, name="y0"); _req(logits, name="logits"); _req(r1, name="r1")
    M, D = y0.shape
    K = logits.shape[1]
    if logits.shape[0] != M or tuple(r1.shape) != (K, D):
        raise RuntimeError(f"cdan_fuse: y0 {tuple(y0.shape)}, logits {tuple(logits.shape)}, r1 {tuple(r1.shape)} do not agree")
    fusion = torch.empty_like(y0)
    prob = torch.empty((M, K), device=y0.device, dtype=torch.float32)
    u = torch.empty((M,), device=y0.device, dtype=torch.float32)
    L.check(L.load().tsc_cdan_fuse_fwd(_ptr(y0), _ptr(logits), _ptr(r1), _ptr(fusion), _ptr(prob), _ptr(u), M, K, D,
                                       float(scale_div), _stream()), "tsc_cdan_fuse_fwd")
    return fusion, prob, u


def cdan_fuse_bwd(dfusion, y0, prob, r1, u, du, coeff, scale_div: float):
    _req(dfusion, name="dfusion"); _req(coeff, name="coeff")
    M, D = y0.shape
    K = prob.shape[1]
    dy0 = torch.empty_like(y0)
    dlogits = torch.empty_like(prob)
    L.check(L.load().tsc_cdan_fuse_bwd(_ptr(dfusion), _ptr(y0), _ptr(prob), _ptr(r1), _ptr(u), _ptr(du), _ptr(coeff),
                                       _ptr(dy0), _ptr(dlogits), M // 2, K, D, float(scale_div), _stream()),
            "tsc_cdan_fuse_bwd")
    return dy0, dlogits


def cdan_distance_fwd(u, critic_out):
    _req(u, name="u"); _req(critic_out, name="critic_out")
    B = u.numel() // 2
    loss = torch.empty((), device=u.device, dtype=torch.float32)
    saved = torch.empty((8,), device=u.device, dtype=torch.float32)
    L.check(L.load().tsc_cdan_distance_fwd(_ptr(u), _ptr(critic_out), _ptr(loss), _ptr(saved), B, _stream()),
            "tsc_cdan_distance_fwd")
    return loss, saved


def cdan_distance_bwd(dloss, saved, B: int):
    _req(dloss, name="dloss")
    du = torch.empty((2 * B,), device=saved.device, dtype=torch.float32)
    dcritic = torch.empty((2 * B, 1), device=saved.device, dtype=torch.float32)
    L.check(L.load().tsc_cdan_distance_bwd(_ptr(dloss), _ptr(saved), _ptr(du), _ptr(dcritic), B, _stream()),
            "tsc_cdan_distance_bwd")
    return du, dcritic


_HEAD_WS = {}


def _head_workspace(B: int, device, owner: int) -> torch.Tensor:
    """Zero-initialised scratch of the head kernel, one per (head weight, batch): the kernel leaves it zeroed, so it is
    allocated (and cleared) once -- no fill launch per step.  Keyed by the weight's address: the classifiers of a step run
    on different streams, but one head never runs twice at the same time."""
    key = (str(device), owner, B)
    ws = _HEAD_WS.get(key)
    if ws is None:
        n = int(L.load().tsc_head_ce_workspace_bytes(B)) // 4 + 4
        ws = _HEAD_WS[key] = torch.zeros(n, device=device, dtype=torch.float32)
    return ws


def head_ce_fwd(pooled: torch.Tensor, W: torch.Tensor, bias: torch.Tensor, labels: Optional[torch.Tensor]):
    """(logits [B,K], prob [B,K], loss () | None) of Linear + softmax cross-entropy (mean over the batch), one launch."""
    _req(pooled, name="pooled"); _req(W, name="head weight"); _req(bias, name="head bias")
    B, C = pooled.shape
    K = W.shape[0]
    if tuple(W.shape) != (K, C) or bias.numel() != K:
        raise RuntimeError(f"head shapes do not agree: pooled {tuple(pooled.shape)}, W {tuple(W.shape)}, bias {tuple(bias.shape)}")
    logits = torch.empty((B, K), device=pooled.device, dtype=torch.float32)
    prob = torch.empty((B, K), device=pooled.device, dtype=torch.float32)
    loss = ws = None
    if labels is not None:
        _req(labels, torch.int64, "labels")
        if labels.numel() != B:
            raise RuntimeError(f"{labels.numel()} labels for {B} rows")
        loss = torch.empty((), device=pooled.device, dtype=torch.float32)
        ws = _head_workspace(B, pooled.device, W.data_ptr())
    L.check(L.load().tsc_head_ce_fwd(_ptr(pooled), _ptr(W), _ptr(bias), _ptr(labels), _ptr(logits), _ptr(prob), _ptr(loss),
                                     _ptr(ws), B, C, K, _stream()), "tsc_head_ce_fwd")
    return logits, prob, loss


def head_ce_bwd(dloss, dlogits, prob, labels, pooled, W, need_dpooled: bool, dW_out=None, dbias_out=None):
    """(dpooled | None, dW, dbias); with dW_out / dbias_out the parameter gradients are ADDED in place."""
    B, C = pooled.shape
    K = W.shape[0]
    dpooled = torch.empty_like(pooled) if need_dpooled else None
    acc = dW_out is not None
    dW = dW_out if acc else torch.empty_like(W)
    db = dbias_out if acc else torch.empty(K, device=W.device, dtype=torch.float32)
    L.check(L.load().tsc_head_ce_bwd(_ptr(dloss), _ptr(dlogits), _ptr(prob), _ptr(labels), _ptr(pooled), _ptr(W), _ptr(dpooled),
                                     _ptr(dW), _ptr(db), 1 if acc else 0, B, C, K, _stream()), "tsc_head_ce_bwd")
    return dpooled, dW, db


def weighted_scalar_sum(terms: Sequence[torch.Tensor], weights: Sequence[float]) -> torch.Tensor:
    """sum_i weights[i] * terms[i] for up to 8 fp32 CUDA scalars, one launch."""
    n = len(terms)
    if not 1 <= n <= 8 or len(weights) != n:
        raise RuntimeError("weighted_scalar_sum takes 1..8 scalars and as many weights")
    ptrs = (ctypes.c_void_p * n)(*[_req(t, name=f"term {i}").data_ptr() for i, t in enumerate(terms)])
    ws = (ctypes.c_float * n)(*[float(w) for w in weights])
    out = torch.empty((), device=terms[0].device, dtype=torch.float32)
    L.check(L.load().tsc_weighted_scalar_sum(ptrs, ws, n, _ptr(out), _stream()), "tsc_weighted_scalar_sum")
    return out


def multi_l2norm(tensors: Sequence[torch.Tensor]) -> torch.Tensor:
    """[||t_0||, ..., ||t_{n-1}||, sum of them] (fp32, device) for up to L.MAX_LIST fp32 CUDA tensors, two launches."""
    if not 1 <= len(tensors) <= L.MAX_LIST:
        raise RuntimeError(f"multi_l2norm takes 1..{L.MAX_LIST} tensors, got {len(tensors)}")
    lst = L.TensorList()
    lst.count = len(tensors)
    keep = []
    for i, t in enumerate(tensors):
        t = _req(t.detach(), name=f"tensor {i}") if t.is_contiguous() else _req(t.detach().contiguous(), name=f"tensor {i}")
        keep.append(t)
        lst.p[i], lst.n[i] = t.data_ptr(), t.numel()
    dev = keep[0].device
    norms = torch.empty(len(tensors) + 1, device=dev, dtype=torch.float32)
    ws = torch.empty(int(L.load().tsc_multi_l2norm_workspace_bytes(len(tensors))) // 4 + 4, device=dev, dtype=torch.float32)
    L.check(L.load().tsc_multi_l2norm(ctypes.byref(lst), _ptr(norms), _ptr(ws), _stream()), "tsc_multi_l2norm")
    return norms


def class_precision(logits: torch.Tensor, labels: Optional[torch.Tensor] = None):
    """(pred [N] int32, counts [2, K] int32, precision [K] fp64) of multi_source_voting.py:296-311."""
    _req(logits, name="logits")
    N, K = logits.shape
    if labels is not None:
        _req(labels, torch.int64, "labels")
        if labels.numel() != N:
            raise RuntimeError(f"{labels.numel()} labels for {N} rows")
    pred = torch.empty(N, device=logits.device, dtype=torch.int32)
    counts = torch.empty((2, K), device=logits.device, dtype=torch.int32)
    prec = torch.empty(K, device=logits.device, dtype=torch.float64)
    L.check(L.load().tsc_class_precision(_ptr(logits), _ptr(labels), _ptr(pred), _ptr(counts), _ptr(prec), N, K, _stream()),
            "tsc_class_precision")
    return pred, counts, prec


def entropy_vote(logits: torch.Tensor, precision: torch.Tensor, entropy_gain: float = 120.0, weight_base: float = 9.0):
    """logits [M, N, K] fp32, precision [M, K] fp64 -> (score [N, K] fp32, pred [N] int32); multi_source_voting.py:357-407."""
    _req(logits, name="logits")
    _req(precision, torch.float64, "precision")
    M, N, K = logits.shape
    if tuple(precision.shape) != (M, K):
        raise RuntimeError(f"precision {tuple(precision.shape)} != {(M, K)}")
    score = torch.empty((N, K), device=logits.device, dtype=torch.float32)
    pred = torch.empty(N, device=logits.device, dtype=torch.int32)
    L.check(L.load().tsc_entropy_vote(_ptr(logits), _ptr(precision), _ptr(score), _ptr(pred), M, N, K, float(entropy_gain),
                                      float(weight_base), _stream()), "tsc_entropy_vote")
    return score, pred


def read_watchdog() -> int:
    code = ctypes.c_int(0)
    L.check(L.load().tsc_debug_read_and_clear_watchdog(ctypes.byref(code)), "tsc_debug_read_and_clear_watchdog")
    return code.value
