#!/usr/bin/env python3
"""Headline benchmark: decoded info Gbit/s at 20 SPA iterations, WiMAX 802.16e n=2304 r1/2.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--frames F] [--impl reference]

One "step" = one pass of the hot path over one batch of F synthetic frames per GPU
(BASELINE.json configs[2]: raw ALIST H, quasi-cyclic z=96 fast path, 20 fixed iterations, early
termination off).  One JSON line is printed by rank 0:

  value      whole-job info Gbit/s with the LLR batch already resident in HBM
             (ldpc_decode_batch on device pointers; one resident-kernel launch per step)
  e2e        the same metric through the reference-facing call SPA_Decoder.decode_batch with
             pinned HOST buffers: H2D of the LLRs and D2H of the packed decisions are inside the
             timed region (ldpc_decode_batch_host)
  roofline   the resident kernel is bound by the SFU (MUFU) pipe, not by HBM:
             achieved = algorithmic transcendentals (2 per edge and pass) per second,
             peak = MUFU ops/s measured on this GPU by a saturating ex2 micro-kernel in this run;
             the HBM view (algorithmic bytes vs MEASURED_PEAKS.json) is reported beside it
  cpu_baseline  the CPU oracle (a C port of the reference's algorithm) on a bounded sample

Under torchrun (N > 1) every rank decodes its own F frames (frames are independent: weak scaling,
no data-path collective); time = max over ranks of the CUDA-event time of the K steps.
``--impl reference`` times the reference's algorithm on the host cores (C port, all threads).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(REPO, "ldpc-simulator_b200")
for _p in (PKG, REPO):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

CODE = "wimax_2304_0.5"
MAX_ITER = 20
EBN0_DB = 2.0
SPEED = 0.5
METRIC = "decoded_info_gbit_per_s_20_spa_iters_wimax_n2304_r12"


def load_code():
    d = np.load(os.path.join(REPO, "tests", "golden", "codes", CODE + ".npz"))
    return int(d["m"]), int(d["n"]), d["row_ptr"].astype(np.int32), d["col_idx"].astype(np.int32)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.first = 0
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def mark(self):
        """Forget what was sampled so far (warm-up); keep sampling."""
        self.first = len(self.rows)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, power = [], [], set(), []
        for r in self.rows[self.first:]:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "power_w_max": max(power), "samples": len(sm)}


def pin_to_gpu_numa_node(gpu_index):
    """Bind this rank to the CPUs NVML reports as local to its GPU (multi-GPU end-to-end runs are limited by
    host-memory locality of the pinned LLR buffers).  Best effort: silently skipped when NVML is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        ncpu = os.cpu_count() or 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        allowed = os.sched_getaffinity(0)
        cpus = {c for c in range(ncpu) if (mask[c // 64] >> (c % 64)) & 1} & allowed
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass


def cpu_reference_run(frames, nthreads, seed=1):
    """Time the oracle port on `frames` frames of the bench workload; returns (info bit/s, seconds)."""
    from oracle import spa_oracle as so
    m, n, rp, ci = load_code()
    rng = np.random.default_rng(seed)
    sig = 1.0 / np.sqrt(2.0 * SPEED * 10 ** (EBN0_DB / 10.0))
    llr = 2.0 * (-1.0 + sig * rng.standard_normal((frames, n))) / sig ** 2
    t0 = time.perf_counter()
    so.decode_batch(rp, ci, n, llr, MAX_ITER, want_post=False, nthreads=nthreads)
    dt = time.perf_counter() - t0
    return frames * (n - m) / dt, dt


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = len(os.sched_getaffinity(0))
    from oracle import spa_oracle as so
    so.build()
    per_step = 192 * cores
    for _ in range(args.warmup):
        cpu_reference_run(max(cores, per_step // 8), cores)
    t_total, bits = 0.0, 0.0
    m, n, _, _ = load_code()
    for s in range(args.steps):
        rate, dt = cpu_reference_run(per_step, cores, seed=100 + s)
        t_total += dt
        bits += per_step * (n - m)
    value = bits / t_total / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Gbit/s", "n_gpus": args.gpus,
        "cores": cores, "frames_per_step": per_step,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"WiMAX 802.16e n=2304 r1/2 raw ALIST H, SPA {MAX_ITER} fixed iterations, "
                               f"{per_step} frames per step (bounded sample of the GPU arm's batch: a rate metric, same code, "
                               f"iterations and channel), Eb/N0 {EBN0_DB} dB, all-zero codeword; {cores} host threads -- the "
                               f"ratio to the GPU arm scales with the host's core count"},
        "cpu_baseline": {"value": value, "unit": "Gbit/s", "cores": cores, "kind": "port",
                         "sample": f"{per_step} frames per step x {args.steps} steps, C port of spa_decoder.py:63-280 "
                                   f"(the reference itself is pure Python and cannot travel to the GPU box), "
                                   f"{cores} POSIX threads"},
        "e2e": {"value": value, "unit": "Gbit/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)
    return 0


GENERIC_WORKLOADS = {
    # name: (code fixture, frames, BASELINE config it stands for, two-sweep check nodes)
    "parity576": ("wimax_576_0.5", 65536, "BASELINE configs[1]: WiMAX-576 r1/2, raw ALIST H (E = 1 824)", False),
    "std576": ("wimax_576_0.5.std", 8192, "config 1' = what main.py really decodes: WiMAX-576 r1/2 on H_std (E = 41 278, check degree 96-192)", True),
    "std2304": ("wimax_2304_0.5.std", 2048, "config 2' = what main.py really decodes: WiMAX-2304 r1/2 on H_std (E = 663 172, check degree 416-632)", True),
}


def run_generic(args):
    """Auxiliary lines: the HBM-streaming generic path (fp64 parity kernels, 20 passes, early termination as the
    reference has it), reported against the measured HBM copy bandwidth."""
    import torch
    import _native
    from channel import Channel
    from settings import Settings
    from spa_decoder import SPA_Decoder
    from scipy import sparse
    torch.cuda.set_device(0)
    fixture, F, label, two_sweep = GENERIC_WORKLOADS[args.workload]
    d = np.load(os.path.join(REPO, "tests", "golden", "codes", fixture + ".npz"))
    m, n = int(d["m"]), int(d["n"])
    h = sparse.csr_matrix((np.ones(d["col_idx"].size, dtype=np.int32), d["col_idx"], d["row_ptr"]), shape=(m, n))

    class Edd:
        _h_sparse_cached = h
        _m, _n = m, n

    st = Settings(); st.set_max_iterations(MAX_ITER); st.set_precision("f64")
    dec = SPA_Decoder(Edd(), st)
    ch = Channel.create_channel(SPEED, EBN0_DB, 0.0, 1, 0.1, 1)
    ch.sigma_sq_quirk = False
    llr = ch.device_llr(F, n, seed=0x5EED, dtype="f64")
    ws = torch.empty(int(_native.lib().ldpc_workspace_bytes(dec.graph.handle, F, _native.LDPC_F64)), dtype=torch.uint8, device="cuda")
    warm = max(args.warmup, 3)
    steps = args.steps if args.workload == "parity576" else min(args.steps, 5)
    for _ in range(warm):
        out = dec.decode_batch_device(llr, workspace=ws)
    torch.cuda.synchronize()
    l0 = _native.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = dec.decode_batch_device(llr, workspace=ws)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    E = dec.graph.nnz
    per_pass = 24 * E + 26 * n + (16 * E if two_sweep else 0)
    alg = F * MAX_ITER * per_pass + F * n * (8 + 8 + 1)               # + load/store transposes
    peak = 6650.0
    pp = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(pp):
        peak = float(json.load(open(pp))["hbm_gbs"])
    line = {"metric": "decoded_info_gbit_per_s_20_spa_iters_%s_fp64_parity_kernels" % args.workload,
            "value": F * (n - m) / (ms * 1e-3) / 1e9,
            "unit": "Gbit/s", "n_gpus": 1, "steps": steps, "warmup": warm, "ms_per_step": ms,
            "higher_is_better": True, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "%s; fp64 generic kernels, %d frames per step, %d passes, working set %.1f GB (> L2); "
                                   "converged fraction %.3f" % (label, F, MAX_ITER, F * (2 * E + 2 * n) * 8 / 1e9, float(out.ok.float().mean()))},
            "gpu_launches": int(_native.launches() - l0),
            "roofline": {"bound": "hbm", "achieved": alg / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg / (ms * 1e-3) / 1e9 / peak, "traffic": None,
                         "note": "whole step (check-node + variable-node + syndrome kernels x 20 passes), algorithmic bytes "
                                 "24E+26n per pass and frame in fp64 (three message sweeps)%s; the check-node kernel is bound by "
                                 "fp64 tanh/atanh issue, not by HBM (ncu: FP64 pipe 44 %%, issue slots 69 %% busy)"
                                 % (" + 16E for the second sweep of rows that do not fit the registers" if two_sweep else "")}}
    emit(line)
    return 0


_JSON_FD = None


def emit(line):
    """The ONE line on stdout.  Everything else a library prints (e.g. the NCCL version banner) goes to stderr:
    main() points file descriptor 1 at stderr for the duration of the run and keeps the real stdout here."""
    text = json.dumps(line) + "\n"
    if _JSON_FD is None:
        sys.stdout.write(text)
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_JSON_FD, text.encode())


def main():
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--frames", type=int, default=131072, help="frames per step per GPU")
    ap.add_argument("--spin", type=float, default=1.0, help="seconds of untimed load before the timed region (after the warm-up steps)")
    ap.add_argument("--impl", type=str, default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-frames", type=int, default=0, help="cpu_baseline sample size (0 = 2048 x cores)")
    ap.add_argument("--workload", default="throughput", choices=["throughput", "parity576", "std576", "std2304"],
                    help="throughput: the headline (configs[2]); parity576 / std576 / std2304: the fp64 generic kernels on "
                         "configs[1] and on the dense H_std graphs main.py really decodes -- HBM roofline of the streaming "
                         "path (auxiliary lines, not the headline)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.workload != "throughput":
        return run_generic(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the decode path has no CPU fallback")
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    all_cpus = os.sched_getaffinity(0)
    pin_to_gpu_numa_node(local_rank)      # pinned staging buffers are first-touched near the GPU's PCIe root

    import _native
    from channel import Channel
    from settings import Settings
    from spa_decoder import SPA_Decoder
    from scipy import sparse

    m, n, rp, ci = load_code()
    k = n - m
    h = sparse.csr_matrix((np.ones(ci.size, dtype=np.int32), ci, rp), shape=(m, n))

    class Edd:
        _h_sparse_cached = h
        _m, _n = m, n

    st = Settings()
    st.set_max_iterations(MAX_ITER)
    st.set_precision("f32_fast")
    st.set_early_termination(False)
    if os.environ.get("LDPC_BENCH_ONE_FRAME"):      # A/B: the one-frame-per-thread resident kernel
        st.set_one_frame_kernel(True)
    if os.environ.get("LDPC_BENCH_ONE_GATHER"):     # A/B: one-frame gather kernel, tensor-memory messages, 4 CTAs per SM
        st.set_one_gather_kernel(True)
    if os.environ.get("LDPC_BENCH_PAIR_SCATTER"):   # A/B: pair kernel, in-place posterior accumulation, messages in TMEM
        st.set_pair_scatter_kernel(True)
    if os.environ.get("LDPC_BENCH_PAIR_REGS"):      # A/B: pair kernel with the messages in registers instead of TMEM
        st.set_pair_regs_kernel(True)
    dec = SPA_Decoder(Edd(), st)
    g = dec.graph
    assert g.is_qc and g.qc_z == 96, "quasi-cyclic fast path not detected"
    F = args.frames
    edges = g.nnz

    # synthetic input: Philox AWGN frames generated on the device (not timed), > L2 (126 MB) per batch
    ch = Channel.create_channel(SPEED, EBN0_DB, 0.0, 1, 0.1, 1)
    ch.sigma_sq_quirk = False
    llr_dev = ch.device_llr(F, n, seed=0x5EED, stream_id=rank)
    llr_host = torch.empty((F, n), dtype=torch.float32).pin_memory()
    llr_host.copy_(llr_dev)
    torch.cuda.synchronize()
    ws = torch.empty(max(256, int(_native.lib().ldpc_workspace_bytes(g.handle, F, _native.LDPC_F32_FAST))),
                     dtype=torch.uint8, device=dev)

    def step_device():
        return dec.decode_batch_device(llr_dev, workspace=ws)

    def step_host():
        return dec.decode_batch(llr_host, want_z=False, want_bits=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing -------------------------------------------------------------------
    out = None
    for _ in range(args.warmup):
        out = step_device()      # like the timed loop, keep the previous result alive while the next one is allocated
    # nvidia-smi is started BEFORE the warm-up so that its start-up (NVML attach) is over when the timed
    # region begins; only the samples taken during the timed region are kept.
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # let clocks and power state settle: keep the GPU busy for >= --spin seconds and until five consecutive steps run
    # within 2 % of the fastest step seen, at most 5 s
    t_spin = time.perf_counter()
    best, recent = float("inf"), []
    while True:
        t_step = time.perf_counter()
        out = step_device()
        torch.cuda.synchronize()
        recent = (recent + [time.perf_counter() - t_step])[-5:]
        best = min(best, recent[-1])
        spun = time.perf_counter() - t_spin
        if os.environ.get("LDPC_BENCH_TRACE") and rank == 0:
            print(f"warm-up t={spun:6.2f} s  step {recent[-1] * 1e3:7.3f} ms", file=sys.stderr)
        if spun >= max(5.0, args.spin) or (spun >= args.spin and len(recent) == 5 and max(recent) <= 1.02 * best):
            break
    barrier()
    if rank == 0:
        sampler.mark()                       # samples from here on belong to the timed region
    launches0 = _native.launches()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e_beg, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_beg.record()
    for s in range(args.steps):
        ev[s][0].record()
        out = step_device()
        ev[s][1].record()
    e_end.record()
    barrier()
    launches = _native.launches() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_total = e_beg.elapsed_time(e_end)
    step_ms = [a.elapsed_time(b) for a, b in ev]
    kernel_ms = float(np.mean(step_ms))
    if os.environ.get("LDPC_BENCH_TRACE") and rank == 0:
        print("timed steps (ms): " + " ".join(f"{t:.2f}" for t in step_ms), file=sys.stderr)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * F * k * args.steps / (ms_total * 1e-3) / 1e9
    ok_frac = float(out.ok.float().mean().item())

    # ---- end to end: pinned host LLRs in, packed decisions out ---------------------------------------
    words = (n + 31) // 32

    def timed_host(step):
        for _ in range(2):
            step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res = step()
        torch.cuda.synchronize()
        mine = time.perf_counter() - t0
        dt = torch.tensor([mine], dtype=torch.float64, device=dev)
        every = [torch.zeros_like(dt) for _ in range(world)]
        if world > 1:
            dist.all_gather(every, dt)
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        else:
            every = [dt]
        return float(dt.item()), [float(t.item()) for t in every], res

    t_e2e, t_ranks, res32 = timed_host(step_host)
    e2e_value = world * F * k * args.steps / t_e2e / 1e9
    h2d = F * n * 4
    d2h = F * (words * 4 + 4 + 1)
    e2e = {"value": e2e_value, "unit": "Gbit/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "h2d_gbs_whole_job": world * h2d * args.steps / t_e2e / 1e9,
           "per_rank_gbit_s": [F * k * args.steps / t / 1e9 for t in t_ranks],
           "note": "fp32 LLRs from pinned host memory; bound by the host->device copy (%.2f GB per step and GPU); "
                   "on the multi-GPU box all ranks share one host's memory / PCIe root complexes" % (h2d / 1e9)}

    # the same call with half precision LLRs on the host (LDPC_FLAG_LLR_F16): a separate, labelled number --
    # not the reference's input type.  Agreement of the decisions with the fp32-ingest run is measured here.
    llr_host16 = torch.empty((F, n), dtype=torch.float16).pin_memory()
    llr_host16.copy_(llr_host)

    def step_host16():
        return dec.decode_batch(llr_host16, want_z=False, want_bits=True, llr_f16=True)

    t_16, t16_ranks, res16 = timed_host(step_host16)
    e2e16 = {"value": world * F * k * args.steps / t_16 / 1e9, "unit": "Gbit/s", "h2d_bytes_per_step": F * n * 2,
             "d2h_bytes_per_step": d2h, "h2d_gbs_whole_job": world * F * n * 2 * args.steps / t_16 / 1e9,
             "per_rank_gbit_s": [F * k * args.steps / t / 1e9 for t in t16_ranks],
             "decision_bit_agreement_vs_fp32_ingest": float(1.0 - np.unpackbits(res16.zbits ^ res32.zbits).mean() * words * 32 / n),
             "syndrome_agreement_vs_fp32_ingest": float((res16.ok == res32.ok).mean()),
             "note": "LLRs rounded to IEEE half on the host, widened to fp32 on the device; this workload (reference sign "
                     "convention, 2 dB) never converges, so its decisions are those of a chaotic trajectory"}
    del llr_host16
    # int8 fixed-point LLRs (LDPC_FLAG_LLR_I8, LLR = q / 4): a quarter of the fp32 bytes; a QUANTISED input, labelled
    from spa_decoder import quantize_llr_i8
    llr_host8 = torch.empty((F, n), dtype=torch.int8).pin_memory()
    llr_host8.copy_(quantize_llr_i8(llr_dev))

    def step_host8():
        return dec.decode_batch(llr_host8, want_z=False, want_bits=True, llr_i8=True)

    t_8, t8_ranks, res8 = timed_host(step_host8)
    e2e8 = {"value": world * F * k * args.steps / t_8 / 1e9, "unit": "Gbit/s", "h2d_bytes_per_step": F * n,
            "d2h_bytes_per_step": d2h, "h2d_gbs_whole_job": world * F * n * args.steps / t_8 / 1e9,
            "per_rank_gbit_s": [F * k * args.steps / t / 1e9 for t in t8_ranks],
            "decision_bit_agreement_vs_fp32_ingest": float(1.0 - np.unpackbits(res8.zbits ^ res32.zbits).mean() * words * 32 / n),
            "syndrome_agreement_vs_fp32_ingest": float((res8.ok == res32.ok).mean()),
            "note": "LLRs quantised on the host to int8, q = round(4 LLR) clipped to +-127, widened on the device; same "
                    "non-converging workload as above, so the decision agreement compares two chaotic trajectories"}
    del llr_host8

    # ---- the Monte-Carlo path: in-kernel Philox channel + decode + counters, one all-reduce per interval ----
    from encoder_decoder_data import EncoderDecoderData
    from matrix_sparse import SparseMatrix
    from mc_driver import MonteCarloEngine, run_intervals
    edd = EncoderDecoderData(h=SparseMatrix(sparse_matrix=h))
    eng = MonteCarloEngine(edd, graph="alist", precision="f32_fast", max_iterations=MAX_ITER, early_termination=False,
                           sigma_sq_quirk=False, seed=0x5EED, device=dev)
    timers = []

    def mc_launch(frames_local, counters, frame_offset):
        eng.launch(frames_local, SPEED, EBN0_DB, counters, frame_offset=frame_offset)

    def mc_run(intervals, record):
        total, _ = run_intervals(mc_launch, device=dev, rank=rank, world=world, group=None, distributed=world > 1,
                                 frames=world * F * intervals, interval_frames=world * F, frame_cursor=0,
                                 timers=timers if record else None)
        return total

    mc_run(2, False)
    barrier()
    mc_l0 = _native.launches()
    m_beg, m_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    m_beg.record()
    mc_total = mc_run(args.steps, True)
    m_end.record()
    barrier()
    mc_wall = time.perf_counter() - t0
    mc_ms = torch.tensor([m_beg.elapsed_time(m_end)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(mc_ms, op=dist.ReduceOp.MAX)
    mc_ms = float(mc_ms.item())
    mc = {"value": world * F * k * args.steps / (mc_ms * 1e-3) / 1e9, "unit": "Gbit/s", "ms_per_interval": mc_ms / args.steps,
          "intervals": args.steps, "frames_per_interval_per_gpu": F, "gpu_launches": int(_native.launches() - mc_l0),
          "allreduce_us": float(np.median([t["allreduce_us"] for t in timers])) if timers else None,
          "host_sync_us": float(np.median([t["host_us"] for t in timers])) if timers else None,
          "kernel_ms": float(np.median([t["kernel_ms"] for t in timers])) if timers else None,
          "kernel_ms_max": float(np.max([t["kernel_ms"] for t in timers])) if timers else None,
          "counters": {"frames": int(mc_total.frames), "frame_errors": int(mc_total.frame_errors), "bit_errors": int(mc_total.bit_errors)},
          "note": "MonteCarloEngine on the bench code: ldpc_mc_run (Philox AWGN generated in the kernel prologue, %d fixed "
                  "passes, error counters folded in the epilogue), one NCCL all-reduce of the 6 int64 counters per interval "
                  "INSIDE the timed region (world %d), then the host reads the reduced counters (stopping rule); "
                  "allreduce_us = device time between the kernel and the end of the all-reduce, host_sync_us = wall time "
                  "of an interval beyond its device time (rank 0)" % (MAX_ITER, world)}

    # ---- the mode main.py actually runs: early termination, a fresh random codeword per frame (device encoder), the
    # decoder with the odd-check sign compensated (the reference's own convention never converges on this code) ----
    eng_et = MonteCarloEngine(edd, graph="alist", precision="f32_fast", max_iterations=MAX_ITER, early_termination=True,
                              fix_odd_check_sign=True, sigma_sq_quirk=False, seed=0x5EED, device=dev)
    et_frames = world * F
    eng_et.run_point(EBN0_DB, SPEED, frames=et_frames)
    barrier()
    e_beg, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_beg.record()
    et_total = None
    for _ in range(3):
        et_total = eng_et.run_point(EBN0_DB, SPEED, frames=et_frames)
    e_end.record()
    barrier()
    et_ms = torch.tensor([e_beg.elapsed_time(e_end) / 3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(et_ms, op=dist.ReduceOp.MAX)
    et_ms = float(et_ms.item())
    mc_et = {"value": et_total.frames * k / (et_ms * 1e-3) / 1e9, "unit": "Gbit/s", "frames_per_s": et_total.frames / (et_ms * 1e-3),
             "ms_per_point": et_ms, "frames_per_point": int(et_total.frames), "ebn0_db": EBN0_DB,
             "fer": et_total.fer(), "ber": et_total.ber(k), "mean_exit_pass": et_total.avg_conv(),
             "note": "NOT the headline configuration (that is 20 fixed passes): MonteCarloEngine.run_point with early "
                     "termination, random codeword per frame from the device encoder, Philox noise generated in the "
                     "kernel, counters all-reduced once per interval; fix_odd_check_sign=True so that frames converge "
                     "(mean_exit_pass = mean iteration index at which the syndrome became zero)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant (only) kernel -------------------------------------------------------
    import ctypes as C
    import hashlib
    pk = C.c_double()
    _native.check(_native.lib().ldpc_measure_mufu_peak(C.byref(pk), None))
    mufu_peak = pk.value
    alg_transc = 2.0 * edges * MAX_ITER * F                 # one tanh + one atanh per edge and pass
    # what the kernel issues per edge and pass: ex2 + lg2 + one reciprocal per two edges (an odd row's last edge has its own):
    # 192 MUFU per pass and frame for the 76 circulants of this base matrix (profiles/r2_sass_gather_wimax2304.txt)
    issued_mufu = (192.0 / 76.0) * edges * MAX_ITER * F
    sfu_achieved = alg_transc / (kernel_ms * 1e-3)
    peaks_path = os.path.join(REPO, "MEASURED_PEAKS.json")
    hbm_peak, hbm_src = 6650.0, "fallback (B200_PROFILING.md)"
    if os.path.exists(peaks_path):
        with open(peaks_path) as f:
            hbm_peak, hbm_src = float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    alg_bytes = F * (n * 4 + n + 4 + 1)                     # LLRs in, z bytes + conv_it + ok out
    # HBM traffic of the kernel comes from an ncu capture (dram__bytes_read.sum + dram__bytes_write.sum); it is only
    # quoted while the kernel sources are the ones that were profiled
    traffic, traffic_src = None, "no ncu capture on record for the current kernel sources"
    src_hash = hashlib.sha256()
    for name in ("qc_kernel.cuh", "qc_kernel_pair.cuh", "qc_kernel_gather.cuh"):
        with open(os.path.join(PKG, "csrc", name), "rb") as f:
            src_hash.update(f.read())
    cap_path = os.path.join(REPO, "profiles", "r2_ncu_traffic.json")
    if os.path.exists(cap_path):
        with open(cap_path) as f:
            cap = json.load(f)
        if cap.get("kernel_src_sha256") == src_hash.hexdigest():
            traffic = cap["dram_bytes_per_frame"] * F
            traffic_src = "ncu capture %s (%s), %.0f B per frame; algorithmic %d B per frame" % (
                cap.get("file"), cap.get("kernel"), cap["dram_bytes_per_frame"], n * 4 + n + 5)
        else:
            traffic_src = "kernel sources changed since the ncu capture %s: not quoted" % cap.get("file")
    sm_hz = (clocks.get("sm_mhz") or 1965.0) * 1e6
    # shared memory, algorithmic bytes per edge, pass and frame: posterior read 4 + previous message read 4 + message to
    # the edge buffer 4 + message read by the variable node 4 + posterior write 4 n/E
    smem_bytes = (16.0 + 4.0 * n / edges) * edges * MAX_ITER * F
    smem_peak = 128.0 * torch.cuda.get_device_properties(dev).multi_processor_count * sm_hz
    roofline = {
        "bound": "sfu", "achieved": sfu_achieved / 1e9, "peak": mufu_peak / 1e9, "unit": "Gop/s",
        "frac": sfu_achieved / mufu_peak, "traffic": traffic, "traffic_source": traffic_src,
        "note": "resident kernel (two frames per thread, messages in shared memory / registers): messages never leave the "
                "SM, HBM is not the bound; achieved = 2 algorithmic transcendentals per edge and pass / kernel time (CUDA "
                "events); peak = MUFU ex2 ops/s measured in this run by ldpc_measure_mufu_peak; the kernel issues 2.53 MUFU "
                "per edge and pass (ex2, lg2, one rcp per two edges), so pipe utilisation = 1.26 x frac.  The instruction "
                "mix itself (not the MUFU pipe) is the floor: its exact opcode mix runs at 51 cycles per edge and frame pair "
                "= 11.5 Gbit/s in isolation (tools/pipe_probe5.cu, profiles/r2_pipe_probe5.txt)",
        "mufu_pipe_utilisation": issued_mufu / (kernel_ms * 1e-3) / mufu_peak,
        "kernel_ms": kernel_ms,
        "smem": {"bound": "smem", "achieved": smem_bytes / (kernel_ms * 1e-3) / 1e9, "peak": smem_peak / 1e9, "unit": "GB/s",
                 "frac": smem_bytes / (kernel_ms * 1e-3) / smem_peak,
                 "note": "algorithmic shared-memory bytes (16 + 4 n/E per edge, pass and frame: posterior read, previous "
                         "message read, message write, message read by the variable node, posterior write) against 128 "
                         "B/clk/SM at the sampled SM clock; ncu LSU wavefront share of the same kernel: "
                         "profiles/r2_ncu_gather_summary.txt"},
        "hbm": {"bound": "hbm", "achieved": alg_bytes / (kernel_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": alg_bytes / (kernel_ms * 1e-3) / 1e9 / hbm_peak, "peak_source": hbm_src},
    }

    # measured parity of this path against the fp64 oracle (tools/parity_fast.py, committed with the kernel)
    parity = None
    par_path = os.path.join(REPO, "profiles", "r2_parity_fast.json")
    if os.path.exists(par_path):
        with open(par_path) as f:
            pr = json.load(f)["regimes"]
        keep = ("frames", "oracle_converged_fraction", "frame_agree", "frame_agree_converged", "bit_agree", "ok_agree",
                "conv_agree", "post_rel_median", "post_rel_p99", "post_frames_outside_tolerance", "flips_near_zero")
        parity = {"source": "profiles/r2_parity_fast.json (tools/parity_fast.py: LDPC_F32_FAST resident kernel vs the fp64 "
                            "oracle on identical LLRs, 20 passes)",
                  "bench_workload": {q: pr["bench"].get(q) for q in keep} if "bench" in pr else None,
                  "converging_regimes_min_frame_agree": min((v["frame_agree"] for kk, v in pr.items()
                                                             if (v.get("oracle_converged_fraction") or 0) > 0.5), default=None)}

    os.sched_setaffinity(0, all_cpus)     # the CPU baseline may use every host core again
    cores = len(all_cpus)
    cpu = None
    if world == 1:
        sample = args.cpu_frames or 2048 * cores
        rate, secs = cpu_reference_run(sample, cores)
        cpu = {"value": rate / 1e9, "unit": "Gbit/s", "cores": cores, "kind": "port", "frames": sample,
               "sample": f"{sample} frames of the same workload in {secs:.1f} s, oracle/spa_oracle.c "
                         f"(C port of spa_decoder.py:63-280), {cores} threads"}

    line = {
        "metric": METRIC, "value": value, "unit": "Gbit/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"WiMAX 802.16e n=2304 r1/2 (BASELINE configs[2]): raw ALIST H, QC z=96 resident kernel, "
                               f"SPA {MAX_ITER} fixed iterations, early termination off, {F} frames per step per GPU; "
                               f"`value` decodes LLRs that were generated beforehand (Philox AWGN, Eb/N0 {EBN0_DB} dB, "
                               f"all-zero codeword) and are resident in HBM; `mc` generates them inside the kernel",
                   "frames_per_step_per_gpu": F, "l2": "input batch larger than L2 (%.0f MB of LLRs per step)" % (h2d / 1e6),
                   "parallelism": f"frames sharded over {world} GPU(s); `value`/`e2e`: no data-path collective, "
                                  f"`mc`: one all-reduce of the counters per interval",
                   "converged_fraction": ok_frac, "host_cores": cores},
        "e2e": e2e,
        "e2e_f16_ingest": e2e16,
        "e2e_i8_ingest": e2e8,
        "mc": mc,
        "mc_early_termination": mc_et,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "parity": parity,
        "cpu_baseline": cpu,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
