"""Shared fixtures.  ``-m "not gpu"``: oracle vs golden vectors, host logic, C-ABI symbol
check (runs on a CPU box).  ``-m gpu``: parity of the CUDA path against the oracle and the
golden vectors, through the C ABI (needs a B200)."""
import json
import os
import sys

import numpy as np
import pytest

# run-time specialised kernels (csrc/qc_jit.cu) cache their cubins here instead of ~/.cache during tests
os.environ.setdefault("LDPC_JIT_CACHE", os.path.join(__import__("tempfile").gettempdir(), "ldpc_b200_jit_tests"))

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(REPO, "ldpc-simulator_b200")
GOLDEN = os.path.join(REPO, "tests", "golden")
for p in (PKG, REPO):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200)")
    config.addinivalue_line("markers", "slow: long running")
    # the oracle is plain C: build it once (gcc); the CUDA library is built in-tree (nvcc)
    from oracle import spa_oracle
    spa_oracle.build()
    import build_native
    if build_native.is_stale():
        build_native.build()


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


class Code:
    """A parity-check pattern stored under tests/golden/codes (CSR index arrays)."""

    def __init__(self, name):
        d = np.load(os.path.join(GOLDEN, "codes", name + ".npz"))
        self.name = name
        self.m, self.n = int(d["m"]), int(d["n"])
        self.row_ptr = d["row_ptr"].astype(np.int32)
        self.col_idx = d["col_idx"].astype(np.int32)

    @property
    def nnz(self):
        return int(self.col_idx.size)

    def csr(self):
        from scipy import sparse
        return sparse.csr_matrix((np.ones(self.nnz, dtype=np.int32), self.col_idx, self.row_ptr),
                                 shape=(self.m, self.n))

    def sparse_matrix(self):
        from matrix_sparse import SparseMatrix
        return SparseMatrix(sparse_matrix=self.csr())


def load_code(name):
    return Code(name)


def load_golden(name):
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    out = {k: d[k] for k in d.files}
    if "z" in out and "n" in out and out["z"].ndim == 2 and out["z"].shape[1] != int(out["n"]):
        out["z"] = np.unpackbits(out["z"], axis=1)[:, : int(out["n"])]
    return out


GOLDEN_DECODE_SETS = ["bch74_std_random", "ccsds128_alist", "ccsds128_std", "tanner155_std", "wifi648_alist",
                      "wimax576_alist", "wimax576_alist_cw", "wimax576_std", "wimax2304_alist",
                      "wimax2304_075B_alist", "wimax2304_083_alist", "wimax2304_std",
                      # round 2: a thicker live-reference pin (2 048 / 256 frames; H_std of the headline code, 20 passes)
                      "wimax576_alist_2k", "wimax2304_alist_256", "wimax2304_std_20it"]


def posterior_violations(got, ref, rel=1e-4, abs_tol=1e-5):
    """Entries outside the north-star tolerance: 1e-4 relative, or 1e-5 absolute near zero."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    err = np.abs(got - ref)
    return (err > abs_tol) & (err > rel * np.abs(ref))


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def stdform_index():
    with open(os.path.join(GOLDEN, "codes", "stdform_index.json")) as f:
        return json.load(f)
