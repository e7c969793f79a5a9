"""Sum-product (SPA) decoder entry point -- CUDA kernels behind the reference's class.

Drop-in for python_ldpc_app/spa_decoder.py: ``SPA_Decoder(encoder_decoder_data, settings)``,
``.decode(data_buffer) -> Result``, ``.convergence_iteration``, plus the same public
attributes (:17-24).  The flooding iteration itself runs in hand-written sm_100a kernels
(csrc/spa_generic.cu for any graph, csrc/spa_qc_resident.cu for quasi-cyclic codes)
reached through the C ABI in include/ldpc_b200.h; this file only marshals buffers.

Conventions kept from the reference (DESIGN.md "conventions"):
* the decoding graph is ``encoder_decoder_data._h_sparse_cached`` (= H_std) when present,
  else ``_h_std.get_sparse_matrix()`` (:28-31); LLR index j is a column of that matrix;
* ``_decoded_data`` receives z = (posterior < 0), the COMPLEMENT of the decided bits
  (:188,233,245); ``Result.OK`` iff the syndrome of z^1 is zero (:191-204,241,253);
* ``convergence_iteration`` is the 0-based pass index, -1 when not converged (:65,232);
* non-convergence is a return value, never an exception.

There is no CPU path: without the CUDA library / a GPU every call raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

import _native
from enums import Result
from matrix_sparse import DeviceGraph

_PRECISIONS = {"f64": _native.LDPC_F64, "f32": _native.LDPC_F32, "f32_fast": _native.LDPC_F32_FAST}


def quantize_llr_i8(llr):
    """LLRs -> the int8 fixed-point format of LDPC_FLAG_LLR_I8: q = round(4 LLR), clipped to [-127, 127]."""
    if hasattr(llr, "data_ptr"):
        import torch
        return torch.clamp(torch.round(llr.to(torch.float32) * 4.0), -127, 127).to(torch.int8)
    return np.clip(np.rint(np.asarray(llr, dtype=np.float32) * 4.0), -127, 127).astype(np.int8)


class BatchResult:
    """Outputs of ``decode_batch``: arrays over frames."""
    __slots__ = ("z", "zbits", "ok", "conv_it", "post", "norm")

    def __init__(self, z, zbits, ok, conv_it, post, norm):
        self.z, self.zbits, self.ok, self.conv_it, self.post, self.norm = z, zbits, ok, conv_it, post, norm

    @property
    def bits(self):
        """Decided bits (un-complemented, as main.py:328 does before counting errors)."""
        return self.z ^ 1


class SPA_Decoder:
    def __init__(self, encoder_decoder_data, settings, graph=None):
        self.m_pData = encoder_decoder_data
        self.m_pSettings = settings
        self._arr_changed_by_iterations = []
        self._normalized_llr_by_iterations = []
        self._normalized_llr_by_iterations_soft = []
        self._d_summarize_normalized_llr = 0.0
        self._arr_aposteriori_llrs = []
        self.convergence_iteration = -1

        if graph is not None:                                  # B200 addition: explicit matrix
            self.H_sparse = graph
        elif hasattr(encoder_decoder_data, "_h_sparse_cached"):
            self.H_sparse = encoder_decoder_data._h_sparse_cached
        else:
            self.H_sparse = encoder_decoder_data._h_std.get_sparse_matrix()
        self._neighbours = None
        self._graph = None

    # ---- lazily built views the reference exposes as attributes (:36-61) ---------
    def _build_neighbours(self):
        if self._neighbours is None:
            d = self.m_pData
            if getattr(d, "_decoder_structures_initialized", False) and self.H_sparse is getattr(d, "_h_sparse_cached", None):
                self._neighbours = d.get_decoder_structures()
            else:
                coo = self.H_sparse.tocoo()
                v2c, c2v = {}, {}
                for i, j in zip(coo.row.tolist(), coo.col.tolist()):
                    v2c.setdefault(j, []).append(i)
                    c2v.setdefault(i, []).append(j)
                self._neighbours = (coo, v2c, c2v)
        return self._neighbours

    @property
    def H_coo(self):
        return self._build_neighbours()[0]

    @property
    def var_to_check(self):
        return self._build_neighbours()[1]

    @property
    def check_to_var(self):
        return self._build_neighbours()[2]

    # ---- device graph ---------------------------------------------------------------
    @property
    def graph(self) -> DeviceGraph:
        if self._graph is None:
            d = self.m_pData
            cache = getattr(d, "_device_graphs", None)
            if cache is not None and self.H_sparse is getattr(d, "_h_sparse_cached", None):
                self._graph = d.device_graph("std")
            else:
                self._graph = DeviceGraph.from_csr(self.H_sparse)
        return self._graph

    def _mode(self, precision=None):
        name = precision or getattr(self.m_pSettings, "get_precision", lambda: "f64")()
        if name not in _PRECISIONS:
            raise ValueError(f"unknown precision {name!r}")
        return name, _PRECISIONS[name]

    def _flags(self, early_termination=None, compact=None, table_kernel=False, jit=True, replay=True):
        s = self.m_pSettings
        early = getattr(s, "is_early_termination", lambda: True)() if early_termination is None else early_termination
        flags = _native.FLAG_EARLY_TERM if early else 0
        if compact or (compact is None and early):     # same results; never slower than masking
            flags |= _native.FLAG_COMPACT
        if table_kernel:
            flags |= _native.FLAG_TABLE_KERNEL
        if not jit:
            flags |= _native.FLAG_NO_JIT
        if not replay:
            flags |= _native.FLAG_NO_REPLAY
        if getattr(s, "is_fix_odd_check_sign", lambda: False)():
            flags |= _native.FLAG_FIX_ODD_SIGN
        if getattr(s, "is_one_frame_kernel", lambda: False)():
            flags |= _native.FLAG_ONE_FRAME
        if getattr(s, "is_pair_regs_kernel", lambda: False)():
            flags |= _native.FLAG_PAIR_REGS
        if getattr(s, "is_pair_scatter_kernel", lambda: False)():
            flags |= _native.FLAG_PAIR_SCATTER
        if getattr(s, "is_pair_gather_kernel", lambda: False)():
            flags |= _native.FLAG_PAIR_GATHER
        if getattr(s, "is_one_gather_kernel", lambda: False)():
            flags |= _native.FLAG_ONE_GATHER
        return flags

    # ---- batched decode, host buffers (the end-to-end call) -------------------------
    def decode_batch(self, llr, *, precision=None, early_termination=None, compact=None, want_z=True,
                     want_bits=False, want_posterior=False, normalized_llr=None, max_iterations=None,
                     table_kernel=False, jit=True, replay=True, llr_f16=False, llr_i8=False):
        """Decode F frames given as a host array ``llr`` [F, n] (numpy, or a pinned CPU torch tensor).

        ``table_kernel`` / ``jit=False`` pick the table-driven resident kernel instead of the one
        specialised for the base matrix (at build time, or with NVRTC at run time); for tests.
        Calls of <= 32 frames on the generic kernels are replayed from a CUDA graph captured on the first
        call with the same configuration (``replay=False`` launches the kernels one by one).

        ``llr_f16=True`` (fp32 precisions only; not the reference's input type): ``llr`` is sent to the device in
        IEEE half precision -- half the PCIe bytes -- and widened to fp32 there (LDPC_FLAG_LLR_F16); a float16
        array / tensor is taken as is, anything else is rounded to float16 first.  ``llr_i8=True``: int8 fixed point
        q = round(4 LLR) clipped to +-127 (``quantize_llr_i8``), a quarter of the bytes (LDPC_FLAG_LLR_I8).

        Returns a ``BatchResult`` of host numpy arrays.  Host<->device copies are pipelined inside
        ``ldpc_decode_batch_host`` (pinned staging, several streams).
        """
        g = self.graph
        name, dtype = self._mode(precision)
        ndt = np.float64 if dtype == _native.LDPC_F64 else np.float32
        odt = ndt                                    # type of the posterior output
        if llr_f16 or llr_i8:
            if dtype == _native.LDPC_F64:
                raise ValueError("llr_f16 / llr_i8 need precision 'f32' or 'f32_fast'")
            if llr_f16 and llr_i8:
                raise ValueError("llr_f16 and llr_i8 exclude each other")
            ndt = np.float16 if llr_f16 else np.int8
            if llr_i8 and not (hasattr(llr, "dtype") and str(llr.dtype) in ("int8", "torch.int8")):
                llr = quantize_llr_i8(llr)
        keep = llr                                   # keep the owner alive during the call
        if hasattr(llr, "data_ptr"):                 # torch CPU tensor (possibly pinned)
            import torch
            want_t = {np.float64: torch.float64, np.float32: torch.float32, np.float16: torch.float16, np.int8: torch.int8}[ndt]
            if llr.device.type != "cpu":
                raise ValueError("decode_batch takes host buffers; use decode_batch_device for CUDA tensors")
            if llr.dtype != want_t or not llr.is_contiguous():
                llr = llr.to(want_t).contiguous()
            keep = llr
            if llr.dim() == 1:
                llr = llr.unsqueeze(0)
            frames, n = llr.shape
            in_ptr = llr.data_ptr()
        else:
            arr = np.ascontiguousarray(np.atleast_2d(np.asarray(llr, dtype=ndt)))
            keep = arr
            frames, n = arr.shape
            in_ptr = arr.ctypes.data
        if n != g.n:
            raise ValueError(f"LLR frames have {n} entries, the graph has {g.n} columns")
        max_it = self.m_pSettings.get_max_iterations() if max_iterations is None else max_iterations
        calc_norm = self.m_pSettings.is_normalized_llr_calculate() if normalized_llr is None else normalized_llr
        words = (n + 31) // 32
        z = np.empty((frames, n), dtype=np.uint8) if want_z else None
        zbits = np.empty((frames, words * 4), dtype=np.uint8) if (want_bits or not want_z) else None
        conv = np.empty(frames, dtype=np.int32)
        ok = np.empty(frames, dtype=np.uint8)
        post = np.empty((frames, n), dtype=odt) if want_posterior else None
        norm = np.empty(frames, dtype=np.float32) if calc_norm else None
        k_info = int(self.m_pData._n - self.m_pData._m) if calc_norm else 0
        ptr = lambda a: a.ctypes.data if a is not None else None
        _native.check(_native.lib().ldpc_decode_batch_host(
            g.handle, dtype, frames, int(max_it),
            self._flags(early_termination, compact, table_kernel, jit, replay) | (_native.FLAG_LLR_F16 if llr_f16 else 0) | (_native.FLAG_LLR_I8 if llr_i8 else 0), in_ptr,
            ptr(z), ptr(zbits), ptr(conv), ptr(ok), ptr(post), ptr(norm), k_info))
        del keep
        return BatchResult(z, zbits, ok, conv, post, norm)

    # ---- batched decode, device tensors (async on the current stream) ---------------
    def decode_batch_device(self, llr, *, precision=None, early_termination=None, compact=None,
                            want_posterior=False, normalized_llr=False, max_iterations=None, workspace=None,
                            table_kernel=False, jit=True, force_generic=False):
        """``llr``: CUDA tensor [F, n] (float64 for 'f64', float32 otherwise).  Returns CUDA tensors."""
        import torch
        g = self.graph
        name, dtype = self._mode(precision)
        want_t = torch.float64 if dtype == _native.LDPC_F64 else torch.float32
        if llr.dtype != want_t or not llr.is_contiguous():
            llr = llr.to(want_t).contiguous()
        frames, n = llr.shape
        if n != g.n:
            raise ValueError(f"LLR frames have {n} entries, the graph has {g.n} columns")
        dev = llr.device
        max_it = self.m_pSettings.get_max_iterations() if max_iterations is None else max_iterations
        z = torch.empty((frames, n), dtype=torch.uint8, device=dev)
        conv = torch.empty(frames, dtype=torch.int32, device=dev)
        ok = torch.empty(frames, dtype=torch.uint8, device=dev)
        post = torch.empty((frames, n), dtype=want_t, device=dev) if want_posterior else None
        norm = torch.empty(frames, dtype=torch.float32, device=dev) if normalized_llr else None
        flags = self._flags(early_termination, compact, table_kernel, jit) | (_native.FLAG_FORCE_GENERIC if force_generic else 0)
        if workspace is None:
            need = int(_native.lib().ldpc_workspace_bytes_ex(g.handle, frames, dtype, flags, int(bool(normalized_llr))))
            if need > (64 << 20):             # (cudaMemGetInfo costs milliseconds: only when the request is large)
                free_b, _tot = torch.cuda.mem_get_info(dev)
                need = min(need, int(free_b * 0.8))
            need = max(need, 256)
            workspace = torch.empty(need, dtype=torch.uint8, device=dev)
        dp = lambda t: t.data_ptr() if t is not None else None
        _native.check(_native.lib().ldpc_decode_batch(
            g.handle, dtype, frames, int(max_it), flags, llr.data_ptr(),
            z.data_ptr(), conv.data_ptr(), ok.data_ptr(), dp(post), dp(norm),
            int(self.m_pData._n - self.m_pData._m), workspace.data_ptr(), workspace.numel(),
            torch.cuda.current_stream(dev).cuda_stream))
        return BatchResult(z, None, ok, conv, post, norm)

    # ---- the reference's per-frame entry point (:63-280) ----------------------------
    def decode(self, p_data_buffer):
        self.convergence_iteration = -1
        llr = np.asarray(p_data_buffer._channel_data, dtype=np.float64)
        calc_norm = bool(self.m_pSettings.is_normalized_llr_calculate())
        res = self.decode_batch(llr[None, :], want_posterior=False, normalized_llr=calc_norm)
        p_data_buffer._decoded_data = res.z[0].astype(np.int32).tolist()
        if calc_norm:
            # The reference appends one count and one value per executed pass (:226-228) and leaves the last one in
            # _d_summarize_normalized_llr (:237-239,249-251).  The kernels produce the value of the EXIT pass; the value
            # after pass p is the exit value of the same decode cut off after p+1 passes (early termination cannot
            # strike earlier: the frame has not converged before its exit pass), so the history is p short decodes.
            passes = (int(res.conv_it[0]) + 1) if res.ok[0] else int(self.m_pSettings.get_max_iterations())
            k = int(self.m_pData._n - self.m_pData._m)
            for p in range(1, passes):
                part = self.decode_batch(llr[None, :], want_z=False, want_bits=True, normalized_llr=True, max_iterations=p)
                value = float(part.norm[0])
                self._arr_changed_by_iterations.append(int(round(value * k)))
                self._normalized_llr_by_iterations.append(value)
            value = float(res.norm[0])
            self._arr_changed_by_iterations.append(int(round(value * k)))
            self._normalized_llr_by_iterations.append(value)
            self._d_summarize_normalized_llr = value
        if res.ok[0]:
            self.convergence_iteration = int(res.conv_it[0])
            return Result.OK
        return Result.DATA_TRANSFER_NOT_OK
