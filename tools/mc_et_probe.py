"""Monte-Carlo path with early termination (what main.py runs) on the alist graph of a quasi-cyclic code: frames per
second and passes per frame against the fixed-iteration rate of the same kernel family.

    python tools/mc_et_probe.py [code] [snr dB ...]

One JSON line per Eb/N0: frames/s with early termination, mean passes per decoded frame (from the counters), the time a
frame would take at the fixed-20-pass rate scaled to those passes, and the ratio (1.0 = no per-frame overhead).
"""
import json
import os
import sys
import time

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "ldpc-simulator_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    from conftest import load_code
    from encoder_decoder_data import EncoderDecoderData
    from mc_driver import MonteCarloEngine

    import _native
    name = sys.argv[1] if len(sys.argv) > 1 else "wimax_2304_0.5"
    snrs = [float(a) for a in sys.argv[2:]] or [1.5, 2.0, 2.5, 3.0, 4.0]
    # LDPC_ET_KERNEL=one_gather | pair_gather | one_frame: force a kernel variant for the early-termination runs (A/B)
    variant = {"one_gather": _native.FLAG_ONE_GATHER, "pair_gather": _native.FLAG_PAIR_GATHER, "one_frame": _native.FLAG_ONE_FRAME}.get(os.environ.get("LDPC_ET_KERNEL", ""), 0)
    code = load_code(name)
    edd = EncoderDecoderData(h=code.sparse_matrix())
    frames = 262144
    rate = (code.n - code.m) / code.n
    fixed = MonteCarloEngine(edd, graph="alist", precision="f32_fast", max_iterations=20, early_termination=False,
                             fix_odd_check_sign=True, seed=5)
    et = MonteCarloEngine(edd, graph="alist", precision="f32_fast", max_iterations=20, early_termination=True,
                          fix_odd_check_sign=True, seed=5, kernel_flags=variant)

    def timed(eng, snr):
        best, res = 1e9, None
        for _ in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            res = eng.run_point(snr, rate, frames=frames)
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        return best, res

    t_fixed, _ = timed(fixed, snrs[0])
    per_pass = t_fixed / frames / 20.0
    print(json.dumps({"code": name, "fixed_20_passes_frames_per_s": round(frames / t_fixed), "us_per_frame_and_pass": round(per_pass * 1e6, 4)}), flush=True)
    for snr in snrs:
        t, res = timed(et, snr)
        ok = res.conv_count
        ferr = res.frame_errors
        # a converged frame ran conv_it + 2 passes (the syndrome of pass p is seen during pass p + 1), a failed one all 20
        passes = (res.conv_sum + 2 * ok + 20 * ferr) / max(res.frames, 1)
        ideal = passes * per_pass * res.frames
        print(json.dumps({"snr_db": snr, "frames_per_s": round(res.frames / t), "info_gbit_s": round(res.frames * (code.n - code.m) / t / 1e9, 2),
                          "fer": ferr / max(res.frames, 1), "mean_passes_per_frame": round(passes, 2), "ms": round(t * 1e3, 2),
                          "ms_at_fixed_rate": round(ideal * 1e3, 2), "efficiency": round(ideal / t, 3)}), flush=True)


if __name__ == "__main__":
    main()
