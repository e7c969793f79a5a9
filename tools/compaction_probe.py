"""Early termination on the generic kernels: masked (default) vs compacted active list (LDPC_FLAG_COMPACT)
across convergence regimes.  python tools/compaction_probe.py"""
import os
import sys
import time

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "ldpc-simulator_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    from channel import Channel
    from conftest import load_code
    from settings import Settings
    from spa_decoder import SPA_Decoder

    class Edd:
        def __init__(self, h):
            self._h_sparse_cached, (self._m, self._n) = h, h.shape

    def timed(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps * 1e3

    cases = (("ccsds_128_64", "f64", 262144, (1.0, 2.5, 4.0), 0.5, False),
             ("wimax_2304_0.83", "f32", 16384, (3.0, 3.6, 4.5), 0.83, False),
             ("wimax_576_0.5", "f64", 65536, (1.0, 2.0, 3.0), 0.5, True))
    for name, prec, frames, snrs, rate, fix in cases:
        code = load_code(name)
        st = Settings()
        st.set_max_iterations(20)
        st.set_precision(prec)
        st.set_fix_odd_check_sign(fix)
        dec = SPA_Decoder(Edd(code.csr()), st)
        for snr in snrs:
            ch = Channel.create_channel(rate, snr, 0.0, 1, 0.1, 1)
            ch.sigma_sq_quirk = False
            llr = ch.device_llr(frames, code.n, seed=3, dtype=prec)
            res = dec.decode_batch_device(llr, force_generic=True)
            conv = res.conv_it[res.conv_it >= 0]
            masked = timed(lambda: dec.decode_batch_device(llr, compact=False, force_generic=True))
            compacted = timed(lambda: dec.decode_batch_device(llr, compact=True, force_generic=True))
            print(name, prec, f"{snr} dB", "converged", round(float(res.ok.float().mean()), 3), "mean iteration",
                  round(float(conv.float().mean()), 1) if conv.numel() else None,
                  "masked", round(masked, 2), "ms, compacted", round(compacted, 2), "ms", flush=True)


if __name__ == "__main__":
    main()
