"""Which kernel family each parity-check matrix of a code database gets (host only, no GPU).

    python tools/catalog_kernels.py <Channel_Codes_Database dir> [--compile]

For every matrix the catalog lists: quasi-cyclic structure (z, base matrix), and whether the SM-resident
kernel specialised for the base matrix applies (registered at build time, or compiled at run time with
NVRTC -- with --compile the NVRTC compilation for sm_100a is actually run and timed).
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "ldpc-simulator_b200"))
import _native  # noqa: E402
from matrix_catalog import MatrixCatalog  # noqa: E402
from utils import read_parity_check_matrix  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("database")
    ap.add_argument("--compile", action="store_true")
    ap.add_argument("--out")
    a = ap.parse_args()
    lib = _native.lib()
    reg = json.load(open(os.path.join(os.path.dirname(_native.LIB_PATH), "..", "csrc", "qc_registry.json")))
    registered = {(c["z"], c["mb"], c["nb"], tuple(np.array(c["shift"]).ravel().tolist())) for c in reg}
    rows = []
    for info in MatrixCatalog(a.database).matrices:
        mat = read_parity_check_matrix(info.path)
        h = mat.get_sparse_matrix()
        qc = mat.detect_qc()
        row = {"name": info.name, "n": int(h.shape[1]), "m": int(h.shape[0]), "nnz": int(h.nnz), "kernel": "generic"}
        if qc is not None:
            z, sh = qc
            row.update(z=int(z), base=f"{sh.shape[0]}x{sh.shape[1]}")
            flat = np.ascontiguousarray(sh, np.int16).ravel()
            if (int(z), sh.shape[0], sh.shape[1], tuple(flat.tolist())) in registered:
                row["kernel"] = "qc_registered"
            else:
                size = C.c_size_t(0)
                t0 = time.time()
                rc = lib.ldpc_host_jit_compile(int(z), sh.shape[0], sh.shape[1], flat.ctypes.data_as(C.POINTER(C.c_int16)),
                                               None, 0, C.byref(size) if a.compile else None)
                if rc == 0:
                    row["kernel"] = "qc_jit"
                    if a.compile:
                        row.update(cubin_bytes=int(size.value), compile_s=round(time.time() - t0, 2))
                else:
                    row["why_not"] = lib.ldpc_last_error().decode()[:120]
        rows.append(row)
        print(json.dumps(row), flush=True)
    kinds = {}
    for r in rows:
        kinds[r["kernel"]] = kinds.get(r["kernel"], 0) + 1
    print(json.dumps({"matrices": len(rows), "by_kernel": kinds}))
    if a.out:
        with open(a.out, "w") as f:
            json.dump({"by_kernel": kinds, "matrices": rows}, f, indent=1)


if __name__ == "__main__":
    main()
