"""ALIST parity-check files -> ``SparseMatrix`` (and back).

Mirror of the reference's reader (python_ldpc_app/utils.py:21-113), same
observable behaviour:

* line 1 ``N M``; line 2 (max weights) ignored; N column weights; M row weights;
  the N per-column lists are skipped; the M per-row lists define H;
* entries are 1-based, ``0`` is padding; an empty row line is an empty row;
* on ANY problem the error is printed and an EMPTY matrix is returned (the
  caller detects ``n == 0``, encoder_decoder_data.py:194-195).

``write_alist`` is an addition used by tools and tests.
"""
from __future__ import annotations

import numpy as np
from scipy import sparse

from matrix_sparse import SparseMatrix


def parse_string_to_int_array(s):
    """Whitespace separated integers of one line."""
    return [int(tok) for tok in s.split()] if s and s.strip() else []


class _AlistError(ValueError):
    pass


def _parse_alist(handle):
    def next_line(what):
        line = handle.readline()
        if not line:
            raise _AlistError(f"Unexpected end of file: missing {what}")
        return line

    head = handle.readline().strip()
    if not head:
        raise _AlistError("Empty file or missing dimensions")
    dims = parse_string_to_int_array(head)
    if len(dims) < 2:
        raise _AlistError("Invalid format: missing dimensions")
    n_cols, n_rows = dims[0], dims[1]
    if n_cols <= 0 or n_rows <= 0:
        raise _AlistError(f"Invalid dimensions: cols={n_cols}, rows={n_rows}")
    handle.readline()                                   # max column / row weight: unused
    col_w = parse_string_to_int_array(next_line("column weights"))
    if len(col_w) != n_cols:
        raise _AlistError(f"Column weights count mismatch: expected {n_cols}, got {len(col_w)}.")
    row_w = parse_string_to_int_array(next_line("row weights"))
    if len(row_w) != n_rows:
        raise _AlistError(f"Row weights count mismatch: expected {n_rows}, got {len(row_w)}.")
    for c in range(n_cols):                             # per-column lists are redundant
        next_line(f"column {c}")
    rows, cols = [], []
    for r in range(n_rows):
        for idx in parse_string_to_int_array(next_line(f"row {r}").strip()):
            if idx == 0:
                continue
            if not 1 <= idx <= n_cols:
                raise _AlistError(f"Invalid column index {idx} in row {r} (valid range: 1-{n_cols})")
            rows.append(r)
            cols.append(idx - 1)
    return n_rows, n_cols, rows, cols


def read_parity_check_matrix(file_name):
    """Read an ALIST file; returns an empty ``SparseMatrix`` on failure (reference behaviour)."""
    try:
        with open(file_name, "r") as fh:
            n_rows, n_cols, rows, cols = _parse_alist(fh)
        ones = np.ones(len(rows), dtype=np.int32)
        h = sparse.coo_matrix((ones, (rows, cols)), shape=(n_rows, n_cols), dtype=np.int32).tocsr()
        return SparseMatrix(sparse_matrix=h)
    except Exception as exc:  # noqa: BLE001 - the reference swallows everything here
        print(f"Error: Could not read parity check matrix from file {file_name}: {exc}")
        return SparseMatrix()


def write_alist(file_name, h):
    """Write a ``SparseMatrix`` / scipy matrix as ALIST text (zero padded lists)."""
    csr = sparse.csr_matrix(h.get_sparse_matrix() if hasattr(h, "get_sparse_matrix") else h)
    csr.sort_indices()
    csc = csr.tocsc()
    csc.sort_indices()
    m, n = csr.shape
    rw = np.diff(csr.indptr)
    cw = np.diff(csc.indptr)
    max_c, max_r = int(cw.max(initial=0)), int(rw.max(initial=0))
    with open(file_name, "w") as fh:
        fh.write(f"{n} {m}\n{max_c} {max_r}\n")
        fh.write(" ".join(str(int(v)) for v in cw) + " \n")
        fh.write(" ".join(str(int(v)) for v in rw) + " \n")
        for j in range(n):
            ent = [int(i) + 1 for i in csc.indices[csc.indptr[j]:csc.indptr[j + 1]]]
            fh.write(" ".join(str(v) for v in ent + [0] * (max_c - len(ent))) + " \n")
        for i in range(m):
            ent = [int(j) + 1 for j in csr.indices[csr.indptr[i]:csr.indptr[i + 1]]]
            fh.write(" ".join(str(v) for v in ent + [0] * (max_r - len(ent))) + " \n")
