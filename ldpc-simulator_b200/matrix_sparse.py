"""GF(2) sparse matrix carrier + the device edge-index builder.

API mirror of the reference's ``SparseMatrix`` (python_ldpc_app/matrix_sparse.py:15-270;
same method names and argument meaning, scipy CSR int32 storage) with the two
additions the B200 decode path needs:

* ``edge_index()`` -- CSR/CSC edge permutations of the pattern (what the
  reference rebuilds as Python dicts per decoder, spa_decoder.py:44-61);
* ``detect_qc()`` / ``device_graph()`` -- quasi-cyclic structure detection
  (z, base matrix of shifts) and the immutable device-resident graph handle the
  CUDA kernels decode on.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
from scipy import sparse

import _native


def _as_csr(mat) -> sparse.csr_matrix:
    out = sparse.csr_matrix(mat, dtype=np.int32)
    out.sum_duplicates()
    out.eliminate_zeros()
    out.sort_indices()
    return out


class DeviceGraph:
    """Owner of an ``ldpc_graph*`` (include/ldpc_b200.h).  Immutable once built."""

    def __init__(self, handle, m, n, nnz):
        self._h = handle
        self.m, self.n, self.nnz = m, n, nnz
        lib = _native.lib()
        vals = [C.c_int() for _ in range(7)]
        nn = C.c_int64()
        _native.check(lib.ldpc_graph_info(self._h, C.byref(vals[0]), C.byref(vals[1]), C.byref(nn),
                                          C.byref(vals[2]), C.byref(vals[3]), C.byref(vals[4]),
                                          C.byref(vals[5]), C.byref(vals[6])))
        self.max_check_degree, self.max_var_degree = vals[2].value, vals[3].value
        self.qc_z, self.qc_mb, self.qc_nb = vals[4].value, vals[5].value, vals[6].value

    @property
    def handle(self):
        if self._h is None:
            raise RuntimeError("graph handle already destroyed")
        return self._h

    @property
    def is_qc(self) -> bool:
        return self.qc_z > 0

    def qc_shifts(self):
        if not self.is_qc:
            return None
        out = np.zeros(self.qc_mb * self.qc_nb, dtype=np.int16)
        _native.check(_native.lib().ldpc_graph_qc_shifts(self.handle, out.ctypes.data_as(C.POINTER(C.c_int16)), out.size))
        return out.reshape(self.qc_mb, self.qc_nb)

    def prepare(self, precision="f32_fast", flags=0) -> str:
        """Select (and, for an unregistered quasi-cyclic base matrix, compile with NVRTC and load) the
        kernels a decode with this precision / flags will run; returns the kernel family name."""
        dt = {"f64": _native.LDPC_F64, "f32": _native.LDPC_F32, "f32_fast": _native.LDPC_F32_FAST}[precision]
        kind = C.c_int(0)
        _native.check(_native.lib().ldpc_graph_prepare(self.handle, dt, int(flags), C.byref(kind)))
        return _native.KERNEL_KINDS[kind.value]

    def close(self):
        if self._h is not None:
            _native.lib().ldpc_graph_destroy(self._h)
            self._h = None

    def __del__(self):  # pragma: no cover - best effort
        try:
            self.close()
        except Exception:
            pass

    @classmethod
    def from_csr(cls, csr) -> "DeviceGraph":
        csr = _as_csr(csr)
        m, n = csr.shape
        rp = np.ascontiguousarray(csr.indptr, dtype=np.int32)
        ci = np.ascontiguousarray(csr.indices, dtype=np.int32)
        h = C.c_void_p()
        _native.check(_native.lib().ldpc_graph_create_csr(
            m, n, int(ci.size), rp.ctypes.data_as(C.POINTER(C.c_int32)), ci.ctypes.data_as(C.POINTER(C.c_int32)),
            C.byref(h)))
        return cls(h, m, n, int(ci.size))

    @classmethod
    def from_qc(cls, z, shifts) -> "DeviceGraph":
        sh = np.ascontiguousarray(shifts, dtype=np.int16)
        mb, nb = sh.shape
        h = C.c_void_p()
        _native.check(_native.lib().ldpc_graph_create_qc(int(z), mb, nb, sh.ctypes.data_as(C.POINTER(C.c_int16)), C.byref(h)))
        g = cls(h, mb * z, nb * z, int((sh >= 0).sum()) * int(z))
        return g


class SparseMatrix:
    """Sparse GF(2) matrix (rows = checks M, columns = code bits N)."""

    def __init__(self, rows=0, cols=0, default_value=0, values=None, sparse_matrix=None):
        if sparse_matrix is not None:
            self._matrix = sparse_matrix
        elif values is not None:
            arr = np.asarray(values, dtype=np.int32)
            if arr.ndim != 2:
                arr = arr.reshape(len(values), -1)
            self._matrix = sparse.csr_matrix(arr)
        else:
            self._matrix = self._filled(rows, cols, default_value)
        self._rows, self._cols = self._matrix.shape
        if values is not None and len(values) == 0:
            self._rows, self._cols = 0, 0

    @staticmethod
    def _filled(rows, cols, value):
        if value == 0:
            return sparse.csr_matrix((rows, cols), dtype=np.int32)
        return sparse.csr_matrix(np.full((rows, cols), value, dtype=np.int32))

    # ---- shape ---------------------------------------------------------------
    def get_rows(self):
        return self._rows

    def get_cols(self):
        return self._cols

    def get_message_bit_length(self):
        return self._cols - self._rows

    # ---- access --------------------------------------------------------------
    def get_data(self):
        return self._matrix.toarray().tolist()

    def get_sparse_matrix(self):
        return self._matrix

    def update_from_sparse_matrix(self, sparse_matrix):
        self._matrix = sparse_matrix

    def _check_index(self, i_row, i_col):
        if not (0 <= i_row < self._rows and 0 <= i_col < self._cols):
            raise IndexError("Index out of bounds")

    def set_element(self, i_row, i_col, i_value):
        self._check_index(i_row, i_col)
        work = self._matrix.tolil()
        work[i_row, i_col] = i_value
        self._matrix = work.tocsr()

    def get_element(self, i_row, i_col):
        self._check_index(i_row, i_col)
        return int(self._matrix[i_row, i_col])

    # ---- GF(2) algebra -------------------------------------------------------
    def multiply(self, mtx):
        if self._cols != mtx.get_rows():
            raise ValueError("Number of columns of the first matrix must be equal to the number of rows "
                             "of the second matrix.")
        prod = (self._matrix @ mtx.get_sparse_matrix()).tocsr()
        prod.data = prod.data % 2
        prod.eliminate_zeros()
        return SparseMatrix(sparse_matrix=prod)

    def transpose(self):
        return SparseMatrix(sparse_matrix=self._matrix.transpose())

    def permute_columns(self, permutation):
        """New column c is old column ``permutation[c]``."""
        if len(permutation) != self._cols:
            raise ValueError("Invalid permutation size")
        perm = np.asarray(permutation, dtype=np.int64)
        self._matrix = self._matrix.tocsc()[:, perm].tocsr()

    def permute_rows(self, permutation):
        """New row r is old row ``permutation[r]``."""
        if len(permutation) != self._rows:
            raise ValueError("Invalid permutation size")
        perm = np.asarray(permutation, dtype=np.int64)
        self._matrix = self._matrix.tocsr()[perm, :]

    def swap_rows(self, row1, row2):
        if not (0 <= row1 < self._rows and 0 <= row2 < self._rows):
            raise IndexError("Row index out of bounds")
        order = np.arange(self._rows)
        order[[row1, row2]] = order[[row2, row1]]
        self._matrix = self._matrix.tocsr()[order, :]

    def extract_sub_matrix(self, start_row, start_col, sub_rows, sub_cols):
        if (start_row < 0 or start_col < 0 or start_row + sub_rows > self._rows
                or start_col + sub_cols > self._cols):
            raise IndexError("Sub-matrix dimensions are out of bounds")
        block = self._matrix.tocsr()[start_row:start_row + sub_rows, start_col:start_col + sub_cols]
        return SparseMatrix(sparse_matrix=block)

    @staticmethod
    def create_identity_matrix(size):
        return SparseMatrix(sparse_matrix=sparse.identity(size, dtype=np.int32, format="csr"))

    def concatenate_horizontally(self, mtx):
        if self._rows != mtx.get_rows():
            raise ValueError("Row counts must match for horizontal concatenation")
        return SparseMatrix(sparse_matrix=sparse.hstack([self._matrix, mtx.get_sparse_matrix()]))

    def concatenate_vertically(self, mtx):
        if self._cols != mtx.get_cols():
            raise ValueError("Column counts must match for vertical concatenation")
        return SparseMatrix(sparse_matrix=sparse.vstack([self._matrix, mtx.get_sparse_matrix()]))

    def add_row(self, row):
        if self._cols != len(row):
            raise ValueError("Must have the same dimensions for addition.")
        self._matrix = sparse.vstack([self._matrix, sparse.csr_matrix([row], dtype=np.int32)])
        self._rows += 1

    def print(self):
        for line in self._matrix.toarray():
            print(" ".join(str(int(v)) for v in line))

    def init(self, rows, cols, default_value=0):
        self._rows, self._cols = rows, cols
        self._matrix = self._filled(rows, cols, default_value)

    # ---- B200 additions ------------------------------------------------------
    def csr_pattern(self):
        """(row_ptr, col_idx) int32 with ascending columns per row -- the edge numbering."""
        csr = _as_csr(self._matrix)
        return (np.ascontiguousarray(csr.indptr, dtype=np.int32),
                np.ascontiguousarray(csr.indices, dtype=np.int32))

    def edge_index(self):
        """CSR/CSC edge permutations: dict(row_ptr, col_idx, col_ptr, csc_edge, edge_row).

        Edges are numbered in CSR order (the COO order the reference iterates,
        spa_decoder.py:41-61); ``csc_edge`` lists each column's edges by ascending row.
        Pure host code (ldpc_host_edge_index), no device needed.
        """
        rp, ci = self.csr_pattern()
        m, n = self._rows, self._cols
        cp = np.zeros(n + 1, dtype=np.int32)
        ce = np.zeros(ci.size, dtype=np.int32)
        er = np.zeros(ci.size, dtype=np.int32)
        p = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
        _native.check(_native.lib().ldpc_host_edge_index(m, n, p(rp), p(ci), p(cp), p(ce), p(er)))
        return dict(row_ptr=rp, col_idx=ci, col_ptr=cp, csc_edge=ce, edge_row=er)

    def detect_qc(self):
        """(z, shifts[mb, nb]) when H is an array of z x z circulants, else None."""
        rp, ci = self.csr_pattern()
        m, n = self._rows, self._cols
        z, mb, nb = C.c_int(), C.c_int(), C.c_int()
        cap = max(1, m * n // 4)
        sh = np.zeros(cap, dtype=np.int16)
        p = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
        rc = _native.check(_native.lib().ldpc_host_detect_qc(
            m, n, p(rp), p(ci), C.byref(z), C.byref(mb), C.byref(nb), sh.ctypes.data_as(C.POINTER(C.c_int16)), cap))
        if rc == 0:
            return None
        return z.value, sh[:mb.value * nb.value].reshape(mb.value, nb.value).copy()

    def device_graph(self) -> DeviceGraph:
        """Upload the edge index (and the QC shift table if any) to the current CUDA device."""
        return DeviceGraph.from_csr(self._matrix)
