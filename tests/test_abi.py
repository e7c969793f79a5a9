"""The C-ABI library loads on a CPU box and exports every symbol include/ldpc_b200.h declares."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import _native
from conftest import REPO, has_gpu


def declared_functions():
    text = open(os.path.join(REPO, "include", "ldpc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ldpc_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_what_python_binds():
    assert declared_functions() == sorted(_native.EXPORTS)


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(_native.LIB_PATH)
    for name in declared_functions():
        assert hasattr(lib, name), name


def test_abi_version_and_error_string():
    lib = _native.lib()
    assert lib.ldpc_abi_version() == _native.ABI_VERSION
    assert isinstance(lib.ldpc_last_error(), bytes)
    assert lib.ldpc_kernel_launch_count() >= 0


def test_missing_library_fails_loudly(monkeypatch):
    monkeypatch.setattr(_native, "_lib", None)
    monkeypatch.setattr(_native, "LIB_PATH", "/nonexistent/libldpc_b200.so")
    with pytest.raises(_native.NativeLibraryError):
        _native.lib()


def test_bad_arguments_are_error_codes_not_crashes():
    lib = _native.lib()
    rp = np.array([0, 2, 1], dtype=np.int32)          # non-monotone
    ci = np.array([0, 1], dtype=np.int32)
    cp = np.zeros(3, dtype=np.int32)
    ce = np.zeros(2, dtype=np.int32)
    p = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
    rc = lib.ldpc_host_edge_index(2, 2, p(rp), p(ci), p(cp), p(ce), None)
    assert rc == -1 and b"monotone" in lib.ldpc_last_error()
    with pytest.raises(_native.LdpcError):
        _native.check(rc)
    assert lib.ldpc_workspace_bytes(None, 10, 0) == 0


@pytest.mark.skipif(has_gpu(), reason="checks the no-device behaviour")
def test_no_cpu_fallback_without_a_device():
    """Graph creation needs a CUDA device; on a CPU box it must fail, not decode on the host."""
    from matrix_sparse import DeviceGraph
    from scipy import sparse
    h = sparse.csr_matrix(np.array([[1, 1, 0], [0, 1, 1]], dtype=np.int32))
    with pytest.raises(_native.LdpcError) as e:
        DeviceGraph.from_csr(h)
    assert e.value.code == -2
