"""ctypes binding of lib/libldpc_b200.so (C ABI: include/ldpc_b200.h).

PyTorch is only the carrier of device memory and streams: every call hands the
library raw ``tensor.data_ptr()`` values and ``torch.cuda.current_stream()``.
There is no CPU fallback -- if the CUDA library is missing or no device is
present the product fails loudly (``NativeLibraryError`` / ``LdpcError``).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", os.environ.get("LDPC_LIB_NAME", "libldpc_b200.so"))

LDPC_F64, LDPC_F32, LDPC_F32_FAST = 0, 1, 2
FLAG_EARLY_TERM, FLAG_COMPACT, FLAG_FIX_ODD_SIGN, FLAG_FORCE_GENERIC, FLAG_TABLE_KERNEL, FLAG_NO_JIT = 0x1, 0x2, 0x4, 0x8, 0x10, 0x20
FLAG_NORM_LLR, FLAG_NO_REPLAY, FLAG_ONE_FRAME, FLAG_PAIR_REGS, FLAG_PAIR_SCATTER, FLAG_PAIR_GATHER = 0x40, 0x80, 0x100, 0x200, 0x400, 0x800
FLAG_LLR_F16, FLAG_LLR_I8, FLAG_ONE_GATHER = 0x1000, 0x2000, 0x4000
CHANNEL_SIGMA_SQ, CHANNEL_AMP_07 = 0x1, 0x2
KERNEL_KINDS = ("generic", "qc_table", "qc_registered", "qc_jit")      # ldpc_kernel_kind
ABI_VERSION = 4

EXPORTS = [
    "ldpc_host_edge_index", "ldpc_host_detect_qc", "ldpc_host_standard_form",
    "ldpc_graph_create_csr", "ldpc_graph_create_qc", "ldpc_graph_info", "ldpc_graph_qc_shifts",
    "ldpc_graph_prepare", "ldpc_host_jit_compile",
    "ldpc_graph_destroy", "ldpc_workspace_bytes", "ldpc_workspace_bytes_ex", "ldpc_mc_workspace_bytes_ex", "ldpc_decode_batch", "ldpc_decode_batch_host",
    "ldpc_mc_run", "ldpc_mc_run_ex", "ldpc_mc_workspace_bytes", "ldpc_channel_llr", "ldpc_channel_llr_ex", "ldpc_encoder_create", "ldpc_encoder_destroy",
    "ldpc_encode_batch", "ldpc_kernel_launch_count",
    "ldpc_measure_mufu_peak", "ldpc_last_error", "ldpc_abi_version",
]


class ChannelDesc(C.Structure):
    """struct ldpc_channel (include/ldpc_b200.h)."""
    _fields_ = [("mode", C.c_int), ("modulation", C.c_int), ("sigma_sq_quirk", C.c_int), ("speed", C.c_double),
                ("snr_db", C.c_double), ("interference_snr_db", C.c_double), ("p", C.c_double)]


class NativeLibraryError(RuntimeError):
    """The CUDA library is not built / cannot be loaded."""


class LdpcError(RuntimeError):
    """A library call returned a negative ldpc_status."""

    def __init__(self, code, message):
        super().__init__(f"libldpc_b200 error {code}: {message}")
        self.code = code


_lib = None


def lib():
    """Load the shared library once.  Never builds implicitly and never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryError(
            f"{LIB_PATH} is missing. Build it with `python ldpc-simulator_b200/build_native.py` "
            "(needs nvcc). This package has no CPU fallback.")
    try:
        l = C.CDLL(LIB_PATH)
    except OSError as e:  # pragma: no cover
        raise NativeLibraryError(f"cannot load {LIB_PATH}: {e}") from e
    vp, i32p, i16p, u8p, u64p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int16), C.POINTER(C.c_uint8), C.POINTER(C.c_uint64)
    ip, i64 = C.POINTER(C.c_int), C.c_int64
    sig = {
        "ldpc_host_edge_index": (C.c_int, [C.c_int, C.c_int, i32p, i32p, i32p, i32p, i32p]),
        "ldpc_host_detect_qc": (C.c_int, [C.c_int, C.c_int, i32p, i32p, ip, ip, ip, i16p, i64]),
        "ldpc_host_standard_form": (C.c_int, [C.c_int, C.c_int, i32p, i32p, u64p, i32p, i32p]),
        "ldpc_graph_create_csr": (C.c_int, [C.c_int, C.c_int, i64, i32p, i32p, C.POINTER(vp)]),
        "ldpc_graph_create_qc": (C.c_int, [C.c_int, C.c_int, C.c_int, i16p, C.POINTER(vp)]),
        "ldpc_graph_info": (C.c_int, [vp, ip, ip, C.POINTER(i64), ip, ip, ip, ip, ip]),
        "ldpc_graph_qc_shifts": (C.c_int, [vp, i16p, i64]),
        "ldpc_graph_prepare": (C.c_int, [vp, C.c_int, C.c_uint, ip]),
        "ldpc_host_jit_compile": (C.c_int, [C.c_int, C.c_int, C.c_int, i16p, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]),
        "ldpc_graph_destroy": (None, [vp]),
        "ldpc_workspace_bytes": (C.c_size_t, [vp, i64, C.c_int]),
        "ldpc_workspace_bytes_ex": (C.c_size_t, [vp, i64, C.c_int, C.c_uint, C.c_int]),
        "ldpc_mc_workspace_bytes_ex": (C.c_size_t, [vp, i64, C.c_int, C.c_uint]),
        "ldpc_decode_batch": (C.c_int, [vp, C.c_int, i64, C.c_int, C.c_uint, vp, vp, vp, vp, vp, vp, C.c_int,
                                        vp, C.c_size_t, vp]),
        "ldpc_decode_batch_host": (C.c_int, [vp, C.c_int, i64, C.c_int, C.c_uint, vp, vp, vp, vp, vp, vp, vp, C.c_int]),
        "ldpc_mc_run": (C.c_int, [vp, C.c_int, i64, C.c_int, C.c_uint, C.c_double, C.c_double, C.c_int,
                                  C.c_uint64, C.c_uint32, C.c_uint64, vp, i64, vp, C.c_int, vp, vp, C.c_size_t, vp]),
        "ldpc_mc_run_ex": (C.c_int, [vp, C.c_int, i64, C.c_int, C.c_uint, C.POINTER(ChannelDesc),
                                     C.c_uint64, C.c_uint32, C.c_uint64, vp, i64, vp, C.c_int, vp, vp, C.c_size_t, vp]),
        "ldpc_channel_llr_ex": (C.c_int, [C.c_int, C.c_int, i64, C.POINTER(ChannelDesc), C.c_uint64,
                                          C.c_uint32, C.c_uint64, vp, i64, vp, vp]),
        "ldpc_mc_workspace_bytes": (C.c_size_t, [vp, i64, C.c_int]),
        "ldpc_channel_llr": (C.c_int, [C.c_int, C.c_int, i64, C.c_double, C.c_double, C.c_int, C.c_uint64,
                                       C.c_uint32, C.c_uint64, vp, i64, vp, vp]),
        "ldpc_encoder_create": (C.c_int, [C.c_int, C.c_int, u64p, i32p, C.POINTER(vp)]),
        "ldpc_encoder_destroy": (None, [vp]),
        "ldpc_encode_batch": (C.c_int, [vp, i64, vp, C.c_uint64, C.c_uint32, C.c_uint64, vp, vp, vp]),
        "ldpc_kernel_launch_count": (C.c_uint64, []),
        "ldpc_measure_mufu_peak": (C.c_int, [C.POINTER(C.c_double), vp]),
        "ldpc_last_error": (C.c_char_p, []),
        "ldpc_abi_version": (C.c_int, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(l, name)
        fn.restype = res
        fn.argtypes = args
    if l.ldpc_abi_version() != ABI_VERSION:
        raise NativeLibraryError("libldpc_b200.so ABI version mismatch; rebuild it")
    _lib = l
    return l


def check(rc):
    if rc < 0:
        raise LdpcError(rc, lib().ldpc_last_error().decode("utf-8", "replace"))
    return rc


def launches() -> int:
    return int(lib().ldpc_kernel_launch_count())
