"""Per-frame carrier object of the pipeline (data -> codeword -> LLRs -> decisions).

Field names and the ``encode`` contract follow python_ldpc_app/data_buffer.py:15-82:
``_data`` (k info bits), ``_encoded_data`` (n bits, ``G^T u mod 2``),
``_channel_data`` (n LLRs, APPENDED by ``Channel.process``), ``_decoded_data`` (what the
decoder writes: z, the complement of the decided bits).  ``FrameBatch`` is the batched
twin used by the GPU path.  The Richardson-Urbanke encoder and the S-random
interleaver of the reference are outside the decode path and not provided.
"""
from __future__ import annotations

import math
import random

import numpy as np

from enums import InterleaverType
from generator import Generator


class DataBuffer:
    def __init__(self, size=0):
        self._size = size
        self._data = Generator.generate_bit_sequence(size) if size > 0 else []
        self._encoded_data = []
        self._decoded_data = []
        self._channel_data = []
        self._interleaving_pos_indexes = []

    def get_size(self):
        return self._size

    def get_data(self):
        return self._data

    def print(self):
        print("Bit Sequence: " + "".join(str(b) for b in self._data))
        print("Encode Sequence: " + "".join(str(b) for b in self._encoded_data))

    def encode(self, G_or_G_transpose):
        """codeword = G^T u mod 2; accepts G (k x n) or its transpose (n x k)."""
        k = len(self._data)
        mat = G_or_G_transpose.get_sparse_matrix()
        if G_or_G_transpose.get_rows() == k:
            mat = mat.transpose()
        elif G_or_G_transpose.get_cols() != k:
            raise ValueError("Matrix dimensions don't match data length")
        u = np.asarray(self._data, dtype=np.int64)
        self._encoded_data = [int(v) for v in (mat.dot(u) % 2)]

    def encode_richardson_urbanke(self, ru_data):
        raise NotImplementedError("Richardson-Urbanke encoding is outside the B200 decode path "
                                  "(the reference's own gap>0 branch is unfinished, data_buffer.py:344-345)")

    # -- interleaving (memoryless AWGN: no effect on statistics; kept for API parity) --
    def calculate_rows_and_cols_for_regular_interleaver(self):
        total = len(self._encoded_data)
        rows = int(math.sqrt(total))
        while rows > 0 and total % rows:
            rows -= 1
        return (rows, total // rows if rows else 0)

    def interleave(self, interleaver_type):
        n = len(self._encoded_data)
        if interleaver_type == InterleaverType.REGULAR:
            rows, cols = self.calculate_rows_and_cols_for_regular_interleaver()
            if rows == 0 or cols == 0:
                return
            # element (r, c) of the row-major block moves to column-major position c*rows + r
            dest = [c * rows + r for r in range(rows) for c in range(cols)]
            out = [0] * n
            for src, d in enumerate(dest):
                out[d] = self._encoded_data[src]
            self._encoded_data, self._interleaving_pos_indexes = out, dest
        elif interleaver_type == InterleaverType.RANDOM:
            order = list(range(n))
            random.shuffle(order)
            self._encoded_data = [self._encoded_data[i] for i in order]
            self._interleaving_pos_indexes = order

    def deinterleave(self, interleaver_type):
        pos = self._interleaving_pos_indexes
        if interleaver_type == InterleaverType.REGULAR:
            self._channel_data = [self._channel_data[p] for p in pos]
        elif interleaver_type == InterleaverType.RANDOM:
            out = [0.0] * len(pos)
            for i, p in enumerate(pos):
                out[p] = self._channel_data[i]
            self._channel_data = out


class FrameBatch:
    """F frames at once: ``data`` [F,k], ``encoded`` [F,n], ``channel`` [F,n] LLRs, ``decoded`` [F,n]."""

    def __init__(self, data=None, encoded=None, channel=None):
        self.data, self.encoded, self.channel = data, encoded, channel
        self.decoded = None
        self.ok = None
        self.convergence_iteration = None
        self.posterior = None

    @classmethod
    def random(cls, frames, edd, rng=None):
        rng = rng or np.random.default_rng()
        data = rng.integers(0, 2, size=(frames, edd._k), dtype=np.uint8)
        return cls(data=data, encoded=edd.encode_batch(data))
