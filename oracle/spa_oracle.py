"""ctypes front-end of oracle/spa_oracle.c -- TEST INFRASTRUCTURE ONLY.

Parity status: pinned against the unmodified reference through
tests/golden/*.npz (see tests/golden/make_golden.py and
tests/test_oracle_golden.py).  Every wrapper names the reference lines
(python_ldpc_app/...) its C counterpart restates.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libspa_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile libspa_oracle.so with the committed Makefile (gcc only)."""
    src = os.path.join(_HERE, "spa_oracle.c")
    stale = (not os.path.exists(_LIB_PATH)) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libspa_oracle.so"])
    return _LIB_PATH


def _load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        build()
    lib = C.CDLL(_LIB_PATH)
    i32p, u8p, f64p, u64p = (C.POINTER(C.c_int32), C.POINTER(C.c_uint8),
                             C.POINTER(C.c_double), C.POINTER(C.c_uint64))
    lib.spa_oracle_decode_batch.restype = C.c_int
    lib.spa_oracle_decode_batch.argtypes = [C.c_int, C.c_int, i32p, i32p, C.c_int64, f64p, C.c_int,
                                            C.c_int, C.c_int, u8p, i32p, u8p, f64p, f64p, C.c_int]
    lib.spa_oracle_decode_batch_ex.restype = C.c_int
    lib.spa_oracle_decode_batch_ex.argtypes = [C.c_int, C.c_int, i32p, i32p, C.c_int64, f64p, C.c_int,
                                               C.c_int, C.c_int, u8p, i32p, u8p, f64p, f64p, C.c_int,
                                               C.c_int, f64p]
    lib.spa_oracle_decode_trace.restype = C.c_int
    lib.spa_oracle_decode_trace.argtypes = [C.c_int, C.c_int, i32p, i32p, f64p, C.c_int, u8p, i32p, f64p]
    lib.spa_oracle_sigma.restype = C.c_double
    lib.spa_oracle_sigma.argtypes = [C.c_double, C.c_double]
    lib.spa_oracle_channel_llr.restype = None
    lib.spa_oracle_channel_llr.argtypes = [C.c_int64, u8p, f64p, C.c_double, C.c_int, C.c_double, f64p]
    lib.spa_oracle_count_errors.restype = None
    lib.spa_oracle_count_errors.argtypes = [C.c_int64, C.c_int, C.c_int, u8p, u8p, i32p, u8p, u64p]
    lib.spa_oracle_standard_form.restype = C.c_int
    lib.spa_oracle_standard_form.argtypes = [C.c_int, C.c_int, u8p, u8p, i32p, i32p]
    lib.spa_oracle_encode.restype = None
    lib.spa_oracle_encode.argtypes = [C.c_int, C.c_int, u8p, u8p, u8p]
    _lib = lib
    return lib


def _p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct)) if a is not None else None


def _csr(row_ptr, col_idx):
    rp = np.ascontiguousarray(row_ptr, dtype=np.int32)
    ci = np.ascontiguousarray(col_idx, dtype=np.int32)
    return rp, ci


def decode_batch(row_ptr, col_idx, n, llr, max_iter, *, want_post=True, calc_norm=False,
                 k_norm=None, nthreads=0, fix_odd_check_sign=False, want_trace=False):
    """SPA_Decoder.decode (spa_decoder.py:63-280) over a batch of frames.

    ``llr`` is [F, n] (or [n]); returns dict(z, conv_it, ok, post, norm) where
    ``z`` is the reference's complemented hard decision (_decoded_data).
    ``fix_odd_check_sign`` (NOT the reference; default off) mirrors the product's non-default switch of
    the same name; ``want_trace`` adds ``post_trace`` [F, max_iter, n] (NaN where a pass did not run).
    """
    lib = _load()
    rp, ci = _csr(row_ptr, col_idx)
    m = rp.size - 1
    llr = np.ascontiguousarray(np.atleast_2d(np.asarray(llr, dtype=np.float64)))
    F = llr.shape[0]
    assert llr.shape[1] == n
    z = np.zeros((F, n), dtype=np.uint8)
    conv = np.zeros(F, dtype=np.int32)
    ok = np.zeros(F, dtype=np.uint8)
    post = np.zeros((F, n), dtype=np.float64) if want_post else None
    norm = np.zeros(F, dtype=np.float64) if calc_norm else None
    if k_norm is None:
        k_norm = n - m
    trace = np.full((F, int(max_iter), n), np.nan) if want_trace else None
    rc = lib.spa_oracle_decode_batch_ex(m, n, _p(rp, C.c_int32), _p(ci, C.c_int32), F,
                                        _p(llr, C.c_double), int(max_iter), int(calc_norm), int(k_norm),
                                        _p(z, C.c_uint8), _p(conv, C.c_int32), _p(ok, C.c_uint8),
                                        _p(post, C.c_double), _p(norm, C.c_double), int(nthreads),
                                        int(bool(fix_odd_check_sign)), _p(trace, C.c_double))
    if rc != 0:
        raise ValueError("spa_oracle_decode_batch rejected its arguments")
    out = dict(z=z, conv_it=conv, ok=ok, post=post, norm=norm)
    if want_trace:
        out["post_trace"] = trace
    return out


def decode_trace(row_ptr, col_idx, n, llr, max_iter):
    """One frame, returning the posterior after every executed check-node pass."""
    lib = _load()
    rp, ci = _csr(row_ptr, col_idx)
    m = rp.size - 1
    llr = np.ascontiguousarray(llr, dtype=np.float64)
    z = np.zeros(n, dtype=np.uint8)
    conv = np.zeros(1, dtype=np.int32)
    trace = np.full((max_iter, n), np.nan)
    good = lib.spa_oracle_decode_trace(m, n, _p(rp, C.c_int32), _p(ci, C.c_int32), _p(llr, C.c_double),
                                       int(max_iter), _p(z, C.c_uint8), _p(conv, C.c_int32),
                                       _p(trace, C.c_double))
    if good < 0:
        raise ValueError("spa_oracle_decode_trace rejected its arguments")
    passes = (conv[0] + 1) if good else max_iter
    return dict(z=z, conv_it=int(conv[0]), ok=bool(good), post_trace=trace[:passes])


def sigma(speed, snr_db):
    """channel.py:113."""
    return float(_load().spa_oracle_sigma(float(speed), float(snr_db)))


def channel_llr(bits, unit_normals, sig, sigma_sq_quirk=True, amp=1.0):
    """channel.py:49-51,68,76,80 with caller-supplied unit normals."""
    bits = np.ascontiguousarray(bits, dtype=np.uint8)
    g = np.ascontiguousarray(unit_normals, dtype=np.float64)
    out = np.empty(g.shape, dtype=np.float64)
    _load().spa_oracle_channel_llr(g.size, _p(bits, C.c_uint8), _p(g, C.c_double), float(sig),
                                   int(bool(sigma_sq_quirk)), float(amp), _p(out, C.c_double))
    return out


def count_errors(z, ok, conv_it, k, data=None):
    """main.py:314-339 counter fold -> uint64[5]."""
    z = np.ascontiguousarray(z, dtype=np.uint8)
    F, n = z.shape
    ok = np.ascontiguousarray(ok, dtype=np.uint8)
    conv_it = np.ascontiguousarray(conv_it, dtype=np.int32)
    d = None if data is None else np.ascontiguousarray(data, dtype=np.uint8)
    cnt = np.zeros(5, dtype=np.uint64)
    _load().spa_oracle_count_errors(F, n, int(k), _p(z, C.c_uint8), _p(ok, C.c_uint8),
                                    _p(conv_it, C.c_int32), _p(d, C.c_uint8), _p(cnt, C.c_uint64))
    return cnt


def standard_form(dense):
    """encoder_decoder_data.py:269-317 -> (H_std dense [rank, n], permutation, rank)."""
    dense = np.ascontiguousarray(dense, dtype=np.uint8)
    m, n = dense.shape
    out = np.zeros((m, n), dtype=np.uint8)
    perm = np.zeros(n, dtype=np.int32)
    rank = np.zeros(1, dtype=np.int32)
    rc = _load().spa_oracle_standard_form(m, n, _p(dense, C.c_uint8), _p(out, C.c_uint8),
                                          _p(perm, C.c_int32), _p(rank, C.c_int32))
    if rc != 0:
        raise MemoryError("spa_oracle_standard_form")
    r = int(rank[0])
    return out[:r].copy(), perm, r


def encode(h_std_dense, u):
    """data_buffer.py:47-82 with G = [I | A^T] (encoder_decoder_data.py:319-344)."""
    h = np.ascontiguousarray(h_std_dense, dtype=np.uint8)
    r, n = h.shape
    u = np.ascontiguousarray(u, dtype=np.uint8)
    cw = np.zeros(n, dtype=np.uint8)
    _load().spa_oracle_encode(r, n, _p(h, C.c_uint8), _p(u, C.c_uint8), _p(cw, C.c_uint8))
    return cw
