"""Monte-Carlo BER/FER engine on the GPU(s).

Replaces the frame loop of ``run_simulation`` / ``process_block``
(python_ldpc_app/main.py:178-442, 43-146).  The reference fans frames over OS processes
(``ProcessPoolExecutor``, main.py:248-256) and folds tuples in the parent; here

* every rank (one process per GPU) generates its own frames on the device with an
  independent Philox stream (``stream_id = rank``), decodes them and folds the five
  integer counters in-kernel (``ldpc_mc_run``);
* only those counters cross GPUs: ONE ``all_reduce(SUM)`` over NCCL per reporting
  interval; the stopping rule is evaluated on the reduced counters so all ranks stop together.

Counter semantics are the reference's (main.py:314-339,357-369): FER = failed frames /
frames; BER = info-bit errors counted only in failed frames / (k * frames); the average
convergence iteration is taken over converged frames only.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

import _native

_PRECISIONS = {"f64": _native.LDPC_F64, "f32": _native.LDPC_F32, "f32_fast": _native.LDPC_F32_FAST}


@dataclass
class PointCounters:
    frames: int = 0
    frame_errors: int = 0
    bit_errors: int = 0
    conv_sum: int = 0
    conv_count: int = 0
    norm_sum: int = 0          # sign changes of the exit pass (normalized-LLR metric), all frames

    def fer(self):
        return self.frame_errors / self.frames if self.frames else 0.0

    def ber(self, k):
        return self.bit_errors / (k * self.frames) if self.frames and k else 0.0

    def avg_conv(self):
        return self.conv_sum / self.conv_count if self.conv_count else 0.0

    def avg_normalized_llr(self, k):
        """main.py:332-334,357: mean over the frames of (sign changes over the first k bits) / k."""
        return self.norm_sum / (k * self.frames) if self.frames and k else 0.0


def wilson_interval(errors, trials, z=1.959963984540054):
    """95 % Wilson score interval (the interval BASELINE.md quotes for the reference's anchors)."""
    if trials == 0:
        return 0.0, 1.0
    p = errors / trials
    den = 1.0 + z * z / trials
    mid = (p + z * z / (2 * trials)) / den
    half = z * math.sqrt(p * (1 - p) / trials + z * z / (4.0 * trials * trials)) / den
    return max(0.0, mid - half), min(1.0, mid + half)


def split_frames(total, rank, world):
    """Frames [lo, hi) of rank ``rank``: contiguous, sizes differ by at most one."""
    base, extra = divmod(int(total), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class MonteCarloEngine:
    def __init__(self, edd, *, graph="std", precision="f64", max_iterations=20, early_termination=True,
                 fix_odd_check_sign=False, sigma_sq_quirk=True, seed=0x5EED, device=None, group=None,
                 kernel_flags=0, normalized_llr=False, modulation=1, mode=1, p=0.1, interference_snr=0.0):
        import torch
        self.torch = torch
        self.edd = edd
        self.graph_name = graph
        self.precision = precision
        self.dtype = _PRECISIONS[precision]
        self.max_iterations = int(max_iterations)
        self.flags = ((_native.FLAG_EARLY_TERM | _native.FLAG_COMPACT) if early_termination else 0) | \
                     (_native.FLAG_FIX_ODD_SIGN if fix_odd_check_sign else 0) | int(kernel_flags)
        if normalized_llr:
            # the metric (spa_decoder.py:210-228) is carried by the generic kernels only
            self.flags |= _native.FLAG_NORM_LLR | _native.FLAG_FORCE_GENERIC
        # the channel of Channel.create_channel(speed, snr, interference_snr, mode, p, modulation) (channel.py:102-125)
        self.quirk = int(bool(sigma_sq_quirk))
        self.modulation, self.mode, self.p, self.interference_snr = int(modulation), int(mode), float(p), float(interference_snr)
        self.seed = int(seed)
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        self.group = group
        dist = torch.distributed
        self.distributed = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if self.distributed else 0
        self.world = dist.get_world_size(group) if self.distributed else 1
        with torch.cuda.device(self.device):
            self.graph = edd.device_graph(graph)
            # compile / load now (NVRTC for an unregistered quasi-cyclic base matrix), not in the first interval
            self.kernel = self.graph.prepare(precision, self.flags)
        self.k = int(edd._k)
        self.n = int(edd._n)
        self._mask = torch.as_tensor(edd.info_mask(graph)).to(self.device)
        self._ws = None
        self._ws_capped = False
        self._frame_cursor = 0           # global frame index: keeps Philox counters unique across calls

    def _workspace(self, frames):
        torch = self.torch
        need = max(256, int(_native.lib().ldpc_mc_workspace_bytes_ex(self.graph.handle, frames, self.dtype, self.flags)))
        if self._ws is not None and (self._ws.numel() >= need or self._ws_capped):
            return self._ws                   # (no cudaMemGetInfo on the hot path: it costs milliseconds)
        self._ws = None
        free_b, _ = torch.cuda.mem_get_info(self.device)
        self._ws_capped = need > int(free_b * 0.7)      # the library then works through the frames in chunks
        self._ws = torch.empty(min(need, max(256, int(free_b * 0.7))), dtype=torch.uint8, device=self.device)
        return self._ws

    def codeword(self, rng):
        """A random codeword in the decoding graph's column order (host encode, [u | A u])."""
        u = rng.integers(0, 2, size=(1, self.k), dtype=np.uint8)
        cw = self.edd.encode_batch(u)[0]
        return cw if self.graph_name == "std" else self.edd.to_alist_order(cw)

    def launch(self, frames_local, speed, snr_db, counters, *, codeword=None, frame_offset=0):
        """Enqueue ``frames_local`` frames on the current stream, accumulating into ``counters`` (cuda int64[5],
        or [6] with ``normalized_llr``).
        ``codeword``: cuda uint8 [n] (sent by every frame) or [frames_local, n] (one per frame) or None (all-zero)."""
        torch = self.torch
        if frames_local <= 0:
            return
        ws = self._workspace(frames_local)
        import ctypes as C
        desc = _native.ChannelDesc(self.mode, self.modulation, self.quirk, float(speed), float(snr_db),
                                   self.interference_snr, self.p)
        _native.check(_native.lib().ldpc_mc_run_ex(
            self.graph.handle, self.dtype, int(frames_local), self.max_iterations, self.flags,
            C.byref(desc), self.seed, int(self.rank), int(frame_offset),
            codeword.data_ptr() if codeword is not None else None,
            (self.n if codeword is not None and codeword.dim() == 2 else 0), self._mask.data_ptr(), self.k,
            counters.data_ptr(), ws.data_ptr(), ws.numel(), torch.cuda.current_stream(self.device).cuda_stream))

    def run_point(self, snr_db, speed, *, frames=None, min_frame_errors=None, max_frames=None,
                  interval_frames=None, random_codewords="frame", rng=None) -> PointCounters:
        """Simulate one SNR point.

        ``frames``: exact total (reference behaviour: ``--blocks`` per point), or
        ``min_frame_errors`` / ``max_frames``: run whole intervals until the reduced counters
        show enough frame errors or the frame budget is spent.
        ``random_codewords``: "frame" = a fresh random codeword per frame (device encoder, what the
        reference does), "interval"/True = one random codeword per interval (host encode), False = all-zero.
        """
        torch = self.torch
        rng = rng or np.random.default_rng(self.seed ^ 0xC0DE)

        def launch(frames_local, counters, frame_offset):
            cw_dev = None
            if random_codewords == "frame" and frames_local > 0:
                # a fresh random codeword per frame, drawn and encoded on the device (main.py:296-303)
                cw_dev = self.edd.device_encoder(self.graph_name).encode(
                    frames_local, seed=self.seed, stream_id=self.rank, frame_offset=frame_offset)
            elif random_codewords:
                cw_dev = torch.as_tensor(self.codeword(rng)).to(self.device)
            self.launch(frames_local, speed, snr_db, counters, codeword=cw_dev, frame_offset=frame_offset)

        with torch.cuda.device(self.device):
            total, self._frame_cursor = run_intervals(
                launch, device=self.device, rank=self.rank, world=self.world, group=self.group,
                distributed=self.distributed, frames=frames, min_frame_errors=min_frame_errors,
                max_frames=max_frames, interval_frames=interval_frames, frame_cursor=self._frame_cursor)
        return total


def run_intervals(launch, *, device, rank, world, group, distributed, frames=None, min_frame_errors=None,
                  max_frames=None, interval_frames=None, frame_cursor=0, timers=None):
    """Host logic of one SNR point, independent of the device that runs ``launch``.

    ``launch(frames_local, counters, frame_offset)`` must accumulate this rank's counters into
    ``counters`` (int64[6] on ``device``: the five of main.py:314-339 plus the normalized-LLR sum).  Per interval: shard frames over ranks, launch, ONE
    all-reduce of the counters, then evaluate the stopping rule on the reduced values (so every rank
    takes the same decision).  Returns (PointCounters, new frame cursor).
    """
    import torch
    dist = torch.distributed
    if frames is None and max_frames is None:
        raise ValueError("give frames= or max_frames=")
    budget = int(frames if frames is not None else max_frames)
    interval = int(interval_frames or budget)
    total = PointCounters()
    done = 0
    while done < budget:
        chunk = min(interval, budget - done)
        lo, hi = split_frames(chunk, rank, world)
        counters = torch.zeros(6, dtype=torch.int64, device=device)
        timed = timers is not None and device.type == "cuda"
        if timed:      # bench.py: where an interval's time goes (CUDA events on the launching stream + wall clock)
            import time
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            t0 = time.perf_counter()
            ev[0].record()
        # ranks number their frames from the shared cursor; their Philox streams differ by stream_id = rank
        launch(hi - lo, counters, frame_cursor + lo)
        if timed:
            ev[1].record()
        if distributed:
            dist.all_reduce(counters, op=dist.ReduceOp.SUM, group=group)      # the one exchange step
        if timed:
            ev[2].record()
        c = counters.cpu().tolist()
        if timed:
            wall = time.perf_counter() - t0
            timers.append({"kernel_ms": ev[0].elapsed_time(ev[1]), "allreduce_us": 1e3 * ev[1].elapsed_time(ev[2]),
                           "host_us": 1e6 * wall - 1e3 * ev[0].elapsed_time(ev[2])})
        total.frames += c[0]; total.frame_errors += c[1]; total.bit_errors += c[2]
        total.conv_sum += c[3]; total.conv_count += c[4]; total.norm_sum += c[5]
        done += chunk
        frame_cursor += chunk
        if frames is None and min_frame_errors is not None and total.frame_errors >= min_frame_errors:
            break
    return total, frame_cursor
