"""ctypes binding of libirb200.so (include/irb200.h).  There is no CPU or PyTorch fallback:
if the CUDA library is missing or the tensors are not on a CUDA device the call fails loudly."""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("IRB200_LIB") or os.path.join(_HERE, "libirb200.so")   # override: A/B builds when tuning

IR_OK, IR_ERR_INVALID, IR_ERR_WORKSPACE, IR_ERR_CUDA, IR_ERR_OOM = 0, -1, -2, -3, -4
MODE_FP32, MODE_HALF, MODE_FP32_SIMT, MODE_FP32_STRICT, MODE_BF16 = 0, 1, 2, 3, 4
ABI_VERSION = 3


class IrRestormerCfg(C.Structure):
    _fields_ = [
        ("inp_channels", C.c_int32), ("out_channels", C.c_int32), ("dim", C.c_int32),
        ("num_blocks", C.c_int32 * 4), ("num_refinement_blocks", C.c_int32), ("heads", C.c_int32 * 4),
        ("ffn_expansion_factor", C.c_double), ("bias", C.c_int32), ("layernorm_with_bias", C.c_int32),
        ("dual_pixel_task", C.c_int32),
    ]


class IrDncnnCfg(C.Structure):
    _fields_ = [("in_nc", C.c_int32), ("out_nc", C.c_int32), ("nc", C.c_int32), ("nb", C.c_int32),
                ("has_bn", C.c_int32)]


class IrKernelStat(C.Structure):
    _fields_ = [("tag", C.c_int32), ("launches", C.c_int32), ("ms", C.c_double), ("bytes", C.c_double),
                ("flops", C.c_double)]


_PP = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); must list every symbol include/irb200.h declares
SIGNATURES = {
    "ir_abi_version": (C.c_int, []),
    "ir_last_error": (C.c_char_p, []),
    "ir_restormer_param_count": (C.c_int, [C.POINTER(IrRestormerCfg)]),
    "ir_restormer_param_numel": (C.c_longlong, [C.POINTER(IrRestormerCfg), C.c_int]),
    "ir_restormer_packed_bytes": (C.c_size_t, [C.POINTER(IrRestormerCfg), C.c_int]),
    "ir_restormer_pack_weights": (C.c_int, [C.POINTER(IrRestormerCfg), _PP, C.c_int, C.c_void_p, C.c_size_t, C.c_int,
                                            C.c_void_p]),
    "ir_restormer_workspace_bytes": (C.c_size_t, [C.POINTER(IrRestormerCfg), C.c_int, C.c_int, C.c_int, C.c_int]),
    "ir_restormer_forward": (C.c_int, [C.POINTER(IrRestormerCfg), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                       C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]),
    "ir_restormer_graph_workspace_bytes": (C.c_size_t, [C.POINTER(IrRestormerCfg), C.c_int, C.c_int, C.c_int, C.c_int]),
    "ir_restormer_forward_graph": (C.c_int, [C.POINTER(IrRestormerCfg), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                             C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]),
    "ir_graph_cache_clear": (C.c_int, []),
    "ir_graph_cache_stats": (C.c_int, [C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
    "ir_restormer_launch_count": (C.c_int, [C.POINTER(IrRestormerCfg)]),
    "ir_dncnn_param_count": (C.c_int, [C.POINTER(IrDncnnCfg)]),
    "ir_dncnn_param_numel": (C.c_longlong, [C.POINTER(IrDncnnCfg), C.c_int]),
    "ir_dncnn_packed_bytes": (C.c_size_t, [C.POINTER(IrDncnnCfg), C.c_int]),
    "ir_dncnn_pack_weights": (C.c_int, [C.POINTER(IrDncnnCfg), _PP, C.c_int, C.c_void_p, C.c_size_t, C.c_int,
                                        C.c_void_p]),
    "ir_dncnn_workspace_bytes": (C.c_size_t, [C.POINTER(IrDncnnCfg), C.c_int, C.c_int, C.c_int, C.c_int]),
    "ir_dncnn_forward": (C.c_int, [C.POINTER(IrDncnnCfg), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                   C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]),
    "ir_dncnn_graph_workspace_bytes": (C.c_size_t, [C.POINTER(IrDncnnCfg), C.c_int, C.c_int, C.c_int, C.c_int]),
    "ir_dncnn_forward_graph": (C.c_int, [C.POINTER(IrDncnnCfg), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                         C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]),
    "ir_dncnn_launch_count": (C.c_int, [C.POINTER(IrDncnnCfg)]),
    "ir_block_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int]),
    "ir_block_packed_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int]),
    "ir_block_pack_weights": (C.c_int, [C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, _PP, C.c_int, C.c_void_p,
                                        C.c_size_t, C.c_int, C.c_void_p]),
    "ir_block_forward": (C.c_int, [C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int,
                                   C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]),
    "ir_nchw_to_nhwc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "ir_nhwc_to_nchw": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "ir_test_conv1x1": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                  C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                  C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                  C.c_void_p]),
    "ir_test_conv3x3": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                  C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                  C.c_size_t, C.c_void_p]),
    "ir_probe_shifted_descriptor": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "ir_tile_gather": (C.c_int, [C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                 C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ir_tile_blend": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float,
                                C.c_void_p]),
    "ir_image_metrics_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "ir_image_metrics": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p,
                                   C.c_void_p, C.c_size_t, C.c_void_p]),
    "ir_profile_begin": (C.c_int, []),
    "ir_profile_end": (C.c_int, [C.POINTER(IrKernelStat), C.c_int]),
    "ir_profile_tag_name": (C.c_char_p, [C.c_int]),
}

_lib = None
_lock = threading.Lock()


def lib():
    """Load the CUDA library once.  Raises RuntimeError (never falls back) if it is missing."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        f"{LIB_PATH} is missing: build it with `python -m image_restoration_models_b200.build` "
                        "(needs nvcc); this package has no CPU / PyTorch fallback")
                l = C.CDLL(LIB_PATH)
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(l, name)      # AttributeError if the .so lacks a declared symbol
                    fn.restype, fn.argtypes = res, args
                if l.ir_abi_version() != ABI_VERSION:
                    raise RuntimeError("libirb200.so ABI version mismatch; rebuild the library")
                _lib = l
    return _lib


def last_error() -> str:
    return lib().ir_last_error().decode("utf-8", "replace")


def check(status: int):
    """Map an IrStatus to the exception the reference's callers expect."""
    if status == IR_OK:
        return
    msg = last_error()
    if status == IR_ERR_INVALID:
        raise ValueError(msg)
    if status == IR_ERR_OOM and "out of memory" not in msg:
        msg = "CUDA out of memory: " + msg
    raise RuntimeError(msg)


def ptr_array(tensors):
    """Host array of device pointers for a list of tensors (None -> NULL)."""
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr


def require_cuda(x, what: str):
    import torch
    if not isinstance(x, torch.Tensor):
        raise TypeError(f"{what}: expected a torch.Tensor")
    if not x.is_cuda:
        raise RuntimeError(f"{what}: tensor is on {x.device}; this implementation runs on CUDA (sm_100a) only and has "
                           "no CPU fallback")
    if x.dtype != torch.float32:
        raise ValueError(f"{what}: expected float32, got {x.dtype}")


class kernel_profile:
    """Context manager: per-kernel-family CUDA-event times of every launch issued inside the block.

    >>> with kernel_profile() as prof: model(x)
    >>> prof.rows   # [{'name', 'launches', 'ms', 'bytes', 'flops'}, ...]
    """

    def __enter__(self):
        check(lib().ir_profile_begin())
        self.rows = []
        return self

    def __exit__(self, *exc):
        buf = (IrKernelStat * 64)()
        n = lib().ir_profile_end(buf, 64)
        if n < 0:
            check(n)
        self.rows = [dict(name=lib().ir_profile_tag_name(buf[i].tag).decode(), launches=buf[i].launches,
                          ms=buf[i].ms, bytes=buf[i].bytes, flops=buf[i].flops) for i in range(n)]
        return False
