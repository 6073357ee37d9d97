"""ctypes binding of libtsc_b200.so (C-ABI in include/tsc_b200.h).

The library is built in-tree by ``build()`` (also called from ``__graft_entry__.build()``) and is the
only compute path of this package: there is no CPU or alternate-backend fallback.  Importing the
package without the built library works (so CPU-only tooling can introspect it), but every operator
raises ``RuntimeError`` the moment it is asked to compute.
"""
from __future__ import annotations

import ctypes
import glob
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, "libtsc_b200.so")
CSRC = os.path.join(_HERE, "csrc")
HEADER = os.path.join(_ROOT, "include", "tsc_b200.h")

TSC_F32, TSC_BF16 = 0, 1
ENGINE_SIMT, ENGINE_TCGEN05 = 0, 1
DIR_FWD, DIR_DGRAD = 0, 1
OUT_C8_F32, OUT_C8_BF16, OUT_NCL_F32, OUT_POOLED = 0, 1, 2, 3
MAX_TAPS, MAX_CHANNELS, MAX_CHANNELS_WIDE = 96, 256, 2048

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC"]

_lock = threading.Lock()
_lib = None

_p, _i, _f, _sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t
_ip = ctypes.POINTER(ctypes.c_int)


class ConvEpilogue(ctypes.Structure):
    """tsc_conv_epilogue of include/tsc_b200.h (device pointers; all nullable)."""
    _fields_ = [("stat_partial", ctypes.c_void_p), ("mask_y", ctypes.c_void_p), ("mask_scale", ctypes.c_void_p),
                ("mask_shift", ctypes.c_void_p), ("mask_mean", ctypes.c_void_p), ("mask_invstd", ctypes.c_void_p),
                ("red_partial", ctypes.c_void_p), ("affine_scale", ctypes.c_void_p), ("affine_shift", ctypes.c_void_p),
                ("residual", ctypes.c_void_p), ("affine_out", ctypes.c_void_p), ("affine_out_kind", ctypes.c_int),
                ("affine_relu", ctypes.c_int), ("bn_gamma", ctypes.c_void_p), ("bn_beta", ctypes.c_void_p),
                ("bn_mean", ctypes.c_void_p), ("bn_var", ctypes.c_void_p), ("bn_eps", ctypes.c_float)]


_ep = ctypes.POINTER(ConvEpilogue)
PACK_MAX_LAYERS = 8


class PackLayer(ctypes.Structure):
    """tsc_pack_layer"""
    _fields_ = [("W", ctypes.c_void_p), ("packed_fwd", ctypes.c_void_p), ("packed_dgrad", ctypes.c_void_p),
                ("Cin", ctypes.c_int), ("Cout", ctypes.c_int), ("Kmax", ctypes.c_int), ("zero_masked", ctypes.c_int),
                ("s_of_tap", ctypes.c_short * 96)]


class PackBatch(ctypes.Structure):
    """tsc_pack_batch"""
    _fields_ = [("n", ctypes.c_int), ("pad_", ctypes.c_int), ("layer", PackLayer * PACK_MAX_LAYERS)]


class BNBranch(ctypes.Structure):
    """tsc_bn_branch"""
    _fields_ = [("y_c8", ctypes.c_void_p), ("stat_partial", ctypes.c_void_p), ("gamma", ctypes.c_void_p),
                ("beta", ctypes.c_void_p), ("running_mean", ctypes.c_void_p), ("running_var", ctypes.c_void_p),
                ("momentum", ctypes.c_float), ("eps", ctypes.c_float), ("coef", ctypes.c_void_p)]


class BNBwdBranch(ctypes.Structure):
    """tsc_bn_bwd_branch"""
    _fields_ = [("y_c8", ctypes.c_void_p), ("coef", ctypes.c_void_p), ("gamma", ctypes.c_void_p),
                ("training", ctypes.c_int), ("red_partial", ctypes.c_void_p), ("dgamma", ctypes.c_void_p),
                ("dbeta", ctypes.c_void_p), ("dbias", ctypes.c_void_p)]


_bp, _bbp = ctypes.POINTER(BNBranch), ctypes.POINTER(BNBwdBranch)
MAX_LIST, MAX_CLASSES, MAX_VOTERS = 32, 64, 8


class TensorList(ctypes.Structure):
    """tsc_tensor_list"""
    _fields_ = [("p", ctypes.c_void_p * MAX_LIST), ("n", ctypes.c_longlong * MAX_LIST), ("count", ctypes.c_int),
                ("pad_", ctypes.c_int)]


# name -> (restype, argtypes); must list every symbol declared in include/tsc_b200.h
SIGNATURES = {
    "tsc_version": (_i, []),
    "tsc_last_error": (ctypes.c_char_p, []),
    "tsc_pad_channels": (_i, [_i]),
    "tsc_device_supports_tcgen05": (_i, []),
    "tsc_ncl_to_c8": (_i, [_p, _p, _i, _i, _i, _i, _p]),
    "tsc_c8_to_ncl": (_i, [_p, _p, _i, _i, _i, _p]),
    "tsc_packed_weight_bytes": (_sz, [_i, _i, _i, _i, _i, _ip]),
    "tsc_pack_weights": (_i, [_i, _i, _p, _p, _i, _i, _i, _ip, _i, _p]),
    "tsc_pack_weights_pair": (_i, [_i, _p, _p, _p, _i, _i, _i, _ip, _i, _p]),
    "tsc_pack_weights_multi": (_i, [_i, ctypes.POINTER(PackBatch), _p]),
    "tsc_bn_apply_fused": (_i, [_bp, _bp, _i, _i, _p, _i, _i, _i, _i, _p]),
    "tsc_bn_fused_splits": (_i, [_i, _i, _i]),
    "tsc_bn_bwd_top": (_i, [_p, _bbp, _bbp, _i, _p, _i, _i, _i, _p]),
    "tsc_bn_bwd_top_pooled": (_i, [_p, _bbp, _i, _p, _i, _i, _i, _p]),
    "tsc_bn_bwd_apply_fused": (_i, [_p, _bbp, _i, _i, _p, _i, _i, _i, _i, _p]),
    "tsc_rmsprop_step": (_i, [_p, _p, _p, ctypes.c_longlong, ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(ctypes.c_float),
                         _i, _f, _f, _f, _p]),
    "tsc_rmsprop_step_clamped": (_i, [_p, _p, _p, ctypes.c_longlong, ctypes.POINTER(ctypes.c_longlong),
                                 ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float), _i, _f, _f, _f, _p]),
    "tsc_osconv_plan_bytes": (_sz, [_i, _i, _i, _i, _ip]),
    "tsc_osconv_plan_build": (_i, [_i, _i, _i, _i, _ip, _p]),
    "tsc_osconv": (_i, [_i, _i, _p, _i, _p, _p, _p, _p, _ep, _i, _i, _i, _i, _i, _ip, _p]),
    "tsc_oswgrad_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i]),
    "tsc_oswgrad": (_i, [_i, _p, _p, _i, _p, _p, _i, _i, _i, _i, _i, _i, _ip, _p]),
    "tsc_bn_workspace_bytes": (_sz, [_i, _i, _i]),
    "tsc_bn_stats": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _f, _f, _i, _i, _i, _p]),
    "tsc_bn_eval_coeffs": (_i, [_p, _p, _p, _p, _f, _p, _p, _p, _p, _i, _p]),
    "tsc_bn_apply": (_i, [_p, _p, _p, _p, _p, _p, _i, _p, _i, _i, _i, _i, _p]),
    "tsc_bn_bwd_reduce": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _p]),
    "tsc_bn_bwd_apply": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "tsc_rowstats_welford": (_i, [_p, _p, _p, _i, _i, _p]),
    "tsc_adain_fwd": (_i, [_p, _p, _p, _p, _f, _i, _i, _p]),
    "tsc_adain_bwd": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _p]),
    "tsc_gram_workspace_bytes": (_sz, [_i, _i, _i]),
    "tsc_gram_loss_fwd": (_i, [_i, _p, _p, _p, _p, _p, _i, _i, _i, _p]),
    "tsc_gram_loss_bwd": (_i, [_i, _p, _p, _p, _p, _p, _p, _i, _i, _i, _p]),
    "tsc_cdan_fuse_fwd": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _f, _p]),
    "tsc_cdan_fuse_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _f, _p]),
    "tsc_cdan_distance_fwd": (_i, [_p, _p, _p, _p, _i, _p]),
    "tsc_cdan_distance_bwd": (_i, [_p, _p, _p, _p, _i, _p]),
    "tsc_multi_l2norm_workspace_bytes": (_sz, [_i]),
    "tsc_multi_l2norm": (_i, [ctypes.POINTER(TensorList), _p, _p, _p]),
    "tsc_class_precision": (_i, [_p, _p, _p, _p, _p, _i, _i, _p]),
    "tsc_entropy_vote": (_i, [_p, _p, _p, _p, _i, _i, _i, _f, _f, _p]),
    "tsc_head_ce_workspace_bytes": (_sz, [_i]),
    "tsc_head_ce_fwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _p]),
    "tsc_head_ce_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "tsc_weighted_scalar_sum": (_i, [ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_float), _i, _p, _p]),
    "tsc_debug_read_and_clear_watchdog": (_i, [_ip]),
    "tsc_debug_set_timeline": (_i, [_p]),
}


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into libtsc_b200.so (nvcc cross-compiles without a GPU): one object per source,
    compiled in parallel and re-used while the source (and every header) is older, then one link."""
    from concurrent.futures import ThreadPoolExecutor
    srcs = sources()
    hdrs = glob.glob(os.path.join(CSRC, "*.cuh")) + [HEADER]
    deps = srcs + hdrs
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(d) for d in deps):
        return LIB_PATH
    objdir = os.path.join(CSRC, "_build")
    os.makedirs(objdir, exist_ok=True)
    hdr_time = max(os.path.getmtime(h) for h in hdrs)
    flags = [f for f in NVCC_FLAGS if f != "-shared"]

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(src), hdr_time):
            return obj
        cmd = ["nvcc"] + flags + ["-c", "-o", obj, src]
        if verbose:
            print(" ".join(cmd))
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, srcs))
    cmd = ["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_PATH] + objs
    if verbose:
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + res.stdout + res.stderr)
    return LIB_PATH


def load():
    """Load the library (once).  Raises RuntimeError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback for the OS-CNN / style-transfer kernels)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.tsc_version() != 100:
            raise RuntimeError(f"libtsc_b200.so version {lib.tsc_version()} does not match the Python host (100)")
        _lib = lib
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        lib = load()
        msg = lib.tsc_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (rc={rc}): {msg}")


def int_array(values):
    return (ctypes.c_int * len(values))(*[int(v) for v in values])
