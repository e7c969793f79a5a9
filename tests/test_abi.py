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


# ---- run-time specialisation (csrc/qc_jit.cu): host side, no GPU needed ---------------------------
def _jit(z, shift, compile_it=False):
    lib = _native.lib()
    sh = np.ascontiguousarray(shift, dtype=np.int16)
    text = C.create_string_buffer(1 << 16)
    size = C.c_size_t(0)
    rc = lib.ldpc_host_jit_compile(int(z), sh.shape[0], sh.shape[1], sh.ctypes.data_as(C.POINTER(C.c_int16)),
                                   text, len(text), C.byref(size) if compile_it else None)
    return rc, text.value.decode(), size.value


def test_run_time_schedule_equals_the_build_time_generator():
    """qc_jit.cu re-implements build_native.schedule_rows in C++: for the registered codes it must
    write exactly the Code<...> type the build generated into qc_codes_gen.cuh."""
    import json
    csrc = os.path.join(REPO, "ldpc-simulator_b200", "csrc")
    gen = open(os.path.join(csrc, "qc_codes_gen.cuh")).read()
    for code in json.load(open(os.path.join(csrc, "qc_registry.json"))):
        rc, text, _ = _jit(code["z"], np.array(code["shift"]))
        assert rc == 0
        ident = "Code_" + "".join(ch if ch.isalnum() else "_" for ch in code["name"])
        built = re.search(r"using %s = (Code<.*?>);\nstatic" % ident, gen, re.S).group(1)
        assert re.sub(r"\s", "", built) == re.sub(r"\s", "", text), code["name"]


def test_nvrtc_specialises_an_unregistered_code_for_sm_100a(tmp_path, monkeypatch):
    """wimax_1152_0.66B (z=48, 8x24 base, check degree 10/11) is in neither the registry nor the shapes
    of the table-driven kernel.  NVRTC cross-compiles without a device; the cubin lands in the cache."""
    from conftest import load_code
    z, shift = load_code("wimax_1152_0.66B").sparse_matrix().detect_qc()
    assert (z, shift.shape) == (48, (8, 24))
    monkeypatch.setenv("LDPC_JIT_CACHE", str(tmp_path))
    rc, text, size = _jit(z, shift, compile_it=True)
    if rc != 0 and b"libnvrtc not found" in _native.lib().ldpc_last_error():
        pytest.skip("no NVRTC on this machine")
    assert rc == 0, _native.lib().ldpc_last_error()
    assert text.startswith("Code<48, 1152,") and size > 100_000
    cached = [f for f in os.listdir(tmp_path) if f.endswith(".cubin")]
    assert len(cached) == 1 and os.path.getsize(tmp_path / cached[0]) == size
    head = open(tmp_path / cached[0], "rb").read(4)
    assert head == b"\x7fELF"
    rc2, _, size2 = _jit(z, shift, compile_it=True)          # second call: served from the disk cache
    assert rc2 == 0 and size2 == size


def test_nvrtc_module_carries_the_gather_kernel_for_narrow_rows(tmp_path, monkeypatch):
    """A base matrix outside the registry whose rows have at most 8 edges (WiMAX r1/2 scaled to z = 28) gets the
    two-frames-per-thread gather kernel in its NVRTC module as well (compiled with explicit 32-bit shared-window
    addressing, csrc/qc_jit.cu); LDPC_JIT_GATHER=0 leaves it out.  Host only: NVRTC cross-compiles without a device."""
    import json
    reg = json.load(open(os.path.join(REPO, "ldpc-simulator_b200", "csrc", "qc_registry.json")))
    base = np.array(reg[0]["shift"])
    shift = np.where(base < 0, -1, base * 28 // 96)
    monkeypatch.setenv("LDPC_JIT_CACHE", str(tmp_path))
    monkeypatch.setenv("LDPC_JIT_GATHER", "0")
    rc, _, small = _jit(28, shift, compile_it=True)
    if rc != 0 and b"libnvrtc not found" in _native.lib().ldpc_last_error():
        pytest.skip("no NVRTC on this machine")
    assert rc == 0, _native.lib().ldpc_last_error()
    monkeypatch.delenv("LDPC_JIT_GATHER")
    rc, _, full = _jit(28, shift, compile_it=True)
    assert rc == 0, _native.lib().ldpc_last_error()
    assert full > small + 200_000                             # a third entry point of ~30 KB of SASS per team path + line info
    cubins = sorted(os.path.getsize(tmp_path / f) for f in os.listdir(tmp_path) if f.endswith(".cubin"))
    assert cubins == sorted([small, full])                    # two cache entries: the source text differs
    blob = b"".join(open(tmp_path / f, "rb").read() for f in os.listdir(tmp_path) if f.endswith(".cubin"))
    assert blob.count(b"ldpc_jit_gather") >= 1 and blob.count(b"ldpc_jit_fixed") >= 2


def test_jit_rejects_what_the_kernel_cannot_run():
    lib = _native.lib()
    rc, _, _ = _jit(8, np.array([[0, 9]]))                    # shift >= z
    assert rc == -1 and b"out of range" in lib.ldpc_last_error()
    rc, _, _ = _jit(8, np.array([[0, -1], [1, -1]]), compile_it=True)     # degree-1 checks, empty column block
    assert rc == -4


def test_the_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under ldpc-simulator_b200/ (the package, its CUDA sources, the
    build recipe) may import, link or execute it; only tests/, bench.py's CPU arms and smoke() do."""
    pkg = os.path.join(REPO, "ldpc-simulator_b200")
    offenders = []
    for root, _dirs, files in os.walk(pkg):
        if os.path.basename(root) in ("build", "lib", "__pycache__"):
            continue
        for name in files:
            if not name.endswith((".py", ".cu", ".cuh", ".json")) or name == "qc_jit_src_gen.cuh":
                continue
            text = open(os.path.join(root, name), encoding="utf-8", errors="replace").read()
            for line in text.splitlines():
                code = line.split("#")[0].split("//")[0]
                if re.search(r"\boracle\b", code) or "spa_oracle" in code:
                    offenders.append((name, line.strip()))
    assert not offenders, offenders
    shared = open(_native.LIB_PATH, "rb").read()
    assert b"spa_oracle" not in shared
