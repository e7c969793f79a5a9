"""Timing of a few usage scenarios around the hot path (looking for anomalies, not a benchmark)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "ldpc-simulator_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    from channel import Channel
    from conftest import load_code
    from encoder_decoder_data import EncoderDecoderData
    from mc_driver import MonteCarloEngine
    from settings import Settings
    from spa_decoder import SPA_Decoder

    class Edd:
        def __init__(self, h):
            self._h_sparse_cached, (self._m, self._n) = h, h.shape

    def decoder(code, precision, iters=20, **kw):
        st = Settings()
        st.set_max_iterations(iters)
        st.set_precision(precision)
        for k, v in kw.items():
            getattr(st, "set_" + k)(v)
        return SPA_Decoder(Edd(code.csr()), st)

    def timed(fn, reps=3):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps * 1e3

    out = []
    # 1. host call, pageable vs pinned, resident kernel
    code = load_code("wimax_2304_0.5")
    dec = decoder(code, "f32_fast")
    F = 65536
    llr = Channel.create_channel(0.5, 2.0, 0.0, 1, 0.1, 1).device_llr(F, code.n, seed=1).cpu()
    pinned = llr.pin_memory()
    for name, x in (("pageable numpy", llr.numpy()), ("pinned torch", pinned)):
        ms = timed(lambda: dec.decode_batch(x, want_z=False, want_bits=True, early_termination=False))
        out.append(dict(scenario=f"host decode_batch {F} frames wimax-2304 f32_fast, {name}", ms=round(ms, 2),
                        info_gbit_s=round(F * 1152 / ms / 1e6, 2)))
    # 2. early termination at high SNR, resident vs fixed
    hi = Channel.create_channel(0.5, 3.0, 0.0, 1, 0.1, 1)
    hi.sigma_sq_quirk = False
    x = hi.device_llr(F, code.n, seed=2)
    dfix = decoder(code, "f32_fast", fix_odd_check_sign=True)
    for early in (False, True):
        ms = timed(lambda: dfix.decode_batch_device(x, early_termination=early))
        out.append(dict(scenario=f"resident wimax-2304 3 dB sign-fixed, early_termination={early}", ms=round(ms, 2),
                        info_gbit_s=round(F * 1152 / ms / 1e6, 2)))
    # 3. generic kernels, early termination with and without compaction (wimax_2304_0.83 converges under the reference signs)
    code83 = load_code("wimax_2304_0.83")
    x83 = Channel.create_channel(0.83, 4.5, 0.0, 1, 0.1, 1)
    x83.sigma_sq_quirk = False
    l83 = x83.device_llr(16384, code83.n, seed=3)
    d83 = decoder(code83, "f32")
    for early, compact in ((False, False), (True, False), (True, True)):
        res = d83.decode_batch_device(l83, early_termination=early, compact=compact)
        ms = timed(lambda: d83.decode_batch_device(l83, early_termination=early, compact=compact))
        out.append(dict(scenario=f"generic f32 wimax-2304-0.83 4.5 dB 16384 frames early={early} compact={compact}", ms=round(ms, 2),
                        converged=round(float(res.ok.float().mean()), 3), mean_it=round(float(res.conv_it[res.conv_it >= 0].float().mean()), 2)))
    # 3b. compaction when nothing converges and the frame count is ragged: must cost nothing
    lo = Channel.create_channel(0.5, 1.0, 0.0, 1, 0.1, 1).device_llr(29127, code.n, seed=4)
    dgen = decoder(code, "f32")
    for compact in (False, True):
        ms = timed(lambda: dgen.decode_batch_device(lo, early_termination=True, compact=compact, force_generic=True))
        out.append(dict(scenario=f"generic f32 wimax-2304 1 dB 29127 frames (none converge) compact={compact}", ms=round(ms, 2)))
    # 4. Monte-Carlo on H_std in fp64 with the reference's kind of block counts
    edd = EncoderDecoderData(h=load_code("wimax_576_0.5").sparse_matrix())
    eng = MonteCarloEngine(edd, graph="std", precision="f64", max_iterations=20, seed=5)
    for blocks in (100, 1000, 4000):
        t0 = time.perf_counter()
        c = eng.run_point(3.0, 0.5, frames=blocks)
        dt = time.perf_counter() - t0
        out.append(dict(scenario=f"run_point H_std wimax-576 fp64 {blocks} blocks (fresh codeword per frame)", ms=round(dt * 1e3, 1),
                        frames_per_s=round(blocks / dt), fer=c.fer()))
    for o in out:
        print(json.dumps(o), flush=True)


if __name__ == "__main__":
    main()
