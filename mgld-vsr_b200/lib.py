"""ctypes binding of libmgld.so (the C ABI in include/mgld.h).

There is deliberately NO fallback: if the shared library is missing or no sm_100a GPU is visible, every op raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MGLD_LIB") or os.path.join(_HERE, "libmgld.so")   # MGLD_LIB: development A/B builds only

_lib = None
_inited = set()
ABI_VERSION = 2     # include/mgld.h MGLD_ABI_VERSION


class MgldError(RuntimeError):
    pass


class ConvGemmDesc(ctypes.Structure):
    """Mirror of ``struct mgld_conv_gemm_desc`` (include/mgld.h)."""
    _fields_ = [
        ("a", ctypes.c_void_p), ("a2", ctypes.c_void_p),
        ("T", ctypes.c_int32), ("H", ctypes.c_int32), ("W", ctypes.c_int32),
        ("C1", ctypes.c_int32), ("C2", ctypes.c_int32),
        ("lda", ctypes.c_int32), ("lda2", ctypes.c_int32),
        ("w", ctypes.c_void_p),
        ("N", ctypes.c_int32), ("taps", ctypes.c_int32), ("block_n", ctypes.c_int32),
        ("epilogue", ctypes.c_int32), ("act", ctypes.c_int32),
        ("bias", ctypes.c_void_p),
        ("alpha", ctypes.c_float), ("beta", ctypes.c_float),
        ("res", ctypes.c_void_p), ("ldres", ctypes.c_int32),
        ("h", ctypes.c_void_p), ("ldh", ctypes.c_int32),
        ("gn_stats", ctypes.c_void_p), ("gn_weight", ctypes.c_void_p), ("gn_bias", ctypes.c_void_p),
        ("groups", ctypes.c_int32),
        ("out", ctypes.c_void_p), ("ldout", ctypes.c_int32), ("out_col0", ctypes.c_int32),
        ("out_f32", ctypes.c_int32),
        ("stats_out", ctypes.c_void_p), ("stats_groups", ctypes.c_int32),
        ("workspace", ctypes.c_void_p), ("workspace_bytes", ctypes.c_int64),
    ]


def load():
    """Load libmgld.so (once).  Raises MgldError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MgldError(f"{LIB_PATH} not found: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU or PyTorch fallback for the mgld kernels)")
    lib = ctypes.CDLL(LIB_PATH)
    lib.mgld_last_error.restype = ctypes.c_char_p
    lib.mgld_abi_version.restype = ctypes.c_int
    if lib.mgld_abi_version() != ABI_VERSION:
        raise MgldError(f"{LIB_PATH} has ABI version {lib.mgld_abi_version()}, this package expects {ABI_VERSION}: rebuild it")
    lib.mgld_conv_gemm_workspace_bytes.restype = ctypes.c_longlong
    _lib = lib
    return lib


def lib():
    """The loaded library, initialised for the current CUDA device."""
    l = load()
    if not torch.cuda.is_available():
        raise MgldError("no CUDA device: the mgld kernels are sm_100a-only and have no CPU fallback")
    dev = torch.cuda.current_device()
    if dev not in _inited:
        check(l.mgld_init(ctypes.c_int(dev)))
        _inited.add(dev)
    return l


def check(rc):
    if rc != 0:
        raise MgldError(f"libmgld error {rc}: {load().mgld_last_error().decode()}")


def stream_ptr():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


class AttentionDesc(ctypes.Structure):
    """Mirror of ``struct mgld_attention_desc`` (include/mgld.h)."""
    _fields_ = [
        ("q", ctypes.c_void_p), ("k", ctypes.c_void_p), ("v", ctypes.c_void_p),
        ("ldq", ctypes.c_int32), ("ldk", ctypes.c_int32), ("ldv", ctypes.c_int32),
        ("q_col0", ctypes.c_int32), ("k_col0", ctypes.c_int32), ("v_col0", ctypes.c_int32),
        ("q_head_stride", ctypes.c_int32), ("k_head_stride", ctypes.c_int32), ("v_head_stride", ctypes.c_int32),
        ("batch", ctypes.c_int32), ("heads", ctypes.c_int32), ("head_dim", ctypes.c_int32),
        ("nq", ctypes.c_int32), ("nkv", ctypes.c_int32), ("kv_batched", ctypes.c_int32),
        ("scale", ctypes.c_float),
        ("out", ctypes.c_void_p), ("ldo", ctypes.c_int32),
    ]
