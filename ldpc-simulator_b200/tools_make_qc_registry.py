"""Write csrc/qc_registry.json: the base matrices the resident kernel is specialised for at build time.

    python ldpc-simulator_b200/tools_make_qc_registry.py

Sources are the code definitions stored under tests/golden/codes (CSR index arrays derived from the
reference's ALIST database by tests/golden/make_golden.py); the shift tables are found by the library's
own quasi-cyclic detector (ldpc_host_detect_qc).  Any other quasi-cyclic code still runs on the
table-driven resident kernel; adding an entry here only makes it faster.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, HERE)

CODES = ["wimax_2304_0.5", "wimax_576_0.5", "wimax_2304_0.75B", "wimax_2304_0.83"]


def main():
    from scipy import sparse
    from matrix_sparse import SparseMatrix
    out = []
    for name in CODES:
        d = np.load(os.path.join(REPO, "tests", "golden", "codes", name + ".npz"))
        m, n = int(d["m"]), int(d["n"])
        h = sparse.csr_matrix((np.ones(d["col_idx"].size, dtype=np.int32), d["col_idx"], d["row_ptr"]), shape=(m, n))
        z, sh = SparseMatrix(sparse_matrix=h).detect_qc()
        out.append(dict(name=name, z=int(z), mb=int(sh.shape[0]), nb=int(sh.shape[1]),
                        shift=[[int(v) for v in row] for row in sh]))
    with open(os.path.join(HERE, "csrc", "qc_registry.json"), "w") as f:
        f.write("[\n" + ",\n".join(
            " {" + f'"name": "{c["name"]}", "z": {c["z"]}, "mb": {c["mb"]}, "nb": {c["nb"]}, "shift": [\n  '
            + ",\n  ".join(json.dumps(row) for row in c["shift"]) + "]}" for c in out) + "\n]\n")
    print("wrote", len(out), "codes")


if __name__ == "__main__":
    main()
