"""Dense list-of-lists twin of ``SparseMatrix``.

The reference carries this class (python_ldpc_app/matrix.py:1-173) but never
imports it; it is kept here only as an API-compatible carrier so that code
written against either class keeps working.  No decode-path compute lives here:
``to_sparse()`` hands the pattern to the sparse/device path.
"""
from __future__ import annotations


class Matrix:
    def __init__(self, rows=0, cols=0, default_value=0, values=None):
        if values is not None:
            self._data = [list(r) for r in values]
            self._rows = len(self._data)
            self._cols = len(self._data[0]) if self._data else 0
        else:
            self._rows, self._cols = rows, cols
            self._data = [[default_value] * cols for _ in range(rows)]

    def get_rows(self):
        return self._rows

    def get_cols(self):
        return self._cols

    def get_message_bit_length(self):
        return self._cols - self._rows

    def get_data(self):
        return self._data

    def _check(self, i_row, i_col):
        if not (0 <= i_row < self._rows and 0 <= i_col < self._cols):
            raise IndexError("Index out of bounds")

    def set_element(self, i_row, i_col, i_value):
        self._check(i_row, i_col)
        self._data[i_row][i_col] = i_value

    def get_element(self, i_row, i_col):
        self._check(i_row, i_col)
        return self._data[i_row][i_col]

    def multiply(self, mtx):
        if self._cols != mtx.get_rows():
            raise ValueError("Number of columns of the first matrix must be equal to the number of rows "
                             "of the second matrix.")
        other = mtx.get_data()
        width = mtx.get_cols()
        out = Matrix(self._rows, width, 0)
        for i, row in enumerate(self._data):
            acc = [0] * width
            for k, v in enumerate(row):
                if v & 1:
                    ok = other[k]
                    for j in range(width):
                        acc[j] ^= ok[j] & 1
            out._data[i] = acc
        return out

    def transpose(self):
        return Matrix(values=[list(col) for col in zip(*self._data)]) if self._data else Matrix()

    def permute_columns(self, permutation):
        if len(permutation) != self._cols:
            raise ValueError("Invalid permutation size")
        self._data = [[row[p] for p in permutation] for row in self._data]

    def swap_rows(self, row1, row2):
        if not (0 <= row1 < self._rows and 0 <= row2 < self._rows):
            raise IndexError("Row index out of bounds")
        self._data[row1], self._data[row2] = self._data[row2], self._data[row1]

    def extract_sub_matrix(self, start_row, start_col, sub_rows, sub_cols):
        if (start_row < 0 or start_col < 0 or start_row + sub_rows > self._rows
                or start_col + sub_cols > self._cols):
            raise IndexError("Sub-matrix dimensions are out of bounds")
        return Matrix(values=[r[start_col:start_col + sub_cols]
                              for r in self._data[start_row:start_row + sub_rows]])

    @staticmethod
    def create_identity_matrix(size):
        return Matrix(values=[[1 if i == j else 0 for j in range(size)] for i in range(size)])

    def concatenate_horizontally(self, mtx):
        if self._rows != mtx.get_rows():
            raise ValueError("Row counts must match for horizontal concatenation")
        return Matrix(values=[a + b for a, b in zip(self._data, mtx.get_data())])

    def concatenate_vertically(self, mtx):
        if self._cols != mtx.get_cols():
            raise ValueError("Column counts must match for vertical concatenation")
        return Matrix(values=self._data + mtx.get_data())

    def add_row(self, row):
        if self._cols != len(row):
            raise ValueError("Must have the same dimensions for addition.")
        self._data.append(list(row))
        self._rows += 1

    def print(self):
        for row in self._data:
            print(" ".join(str(v) for v in row))

    def init(self, rows, cols, default_value=0):
        self.__init__(rows, cols, default_value)

    def to_sparse(self):
        """Hand the pattern to the sparse / device path (B200 addition)."""
        from matrix_sparse import SparseMatrix
        return SparseMatrix(values=self._data) if self._data else SparseMatrix()
