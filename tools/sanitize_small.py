"""Small run through every kernel family, meant to be executed under compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/sanitize_small.py [legacy|pair]
    compute-sanitizer --tool racecheck python tools/sanitize_small.py [legacy|pair]

``legacy`` (default): every kernel family of round 1.  ``pair``: the two-frames-per-thread resident kernels of
round 2 (packed fp32; messages in registers or in TENSOR MEMORY via tcgen05.alloc/ld/st; barrier-free check-node
phase + gather), host-fed with TMA prefetch and in Monte-Carlo mode, even and odd frame counts, early termination.

Covers: generic fp64/fp32 kernels (register and two-sweep check nodes, compaction, normalized-LLR
metric), the table-driven, registered and run-time specialised resident kernels (host-fed with TMA
prefetch, early termination queue, in-kernel Philox Monte-Carlo), the channel generator, the encoder and
the host pipeline with pageable buffers.
"""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "ldpc-simulator_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    from conftest import load_code
    from encoder_decoder_data import EncoderDecoderData
    from mc_driver import MonteCarloEngine
    from settings import Settings
    from spa_decoder import SPA_Decoder

    class Edd:
        def __init__(self, h):
            self._h_sparse_cached, (self._m, self._n) = h, h.shape

    def decoder(code, precision, iters=6):
        st = Settings()
        st.set_max_iterations(iters)
        st.set_precision(precision)
        return SPA_Decoder(Edd(code.csr()), st)

    rng = np.random.default_rng(1)
    done = []
    part = sys.argv[1] if len(sys.argv) > 1 else "legacy"
    if part == "pair":
        import _native
        for name, frames in [("wimax_2304_0.5", 701), ("wimax_576_0.5", 1190)]:
            code = load_code(name)
            llr = rng.normal(-1.0, 2.0, size=(frames, code.n)).astype(np.float32)
            for variant in ("pair_gather_kernel", "pair_scatter_kernel", "pair_regs_kernel", "one_gather_kernel"):
                st = Settings()
                st.set_max_iterations(4)
                st.set_precision("f32_fast")
                getattr(st, "set_" + variant)(True)
                d = SPA_Decoder(Edd(code.csr()), st)
                for early in (False, True):
                    d.decode_batch(llr, early_termination=early, want_posterior=True, want_bits=True)
                d.decode_batch(llr.astype(np.float16), early_termination=False, llr_f16=True, want_z=False, want_bits=True)
                d.decode_batch_device(torch.as_tensor(llr).cuda()[1:], early_termination=False)
                done.append(f"{variant} {name}")
            edd = EncoderDecoderData(h=code.sparse_matrix())
            for flags in (_native.FLAG_PAIR_GATHER, _native.FLAG_PAIR_SCATTER, _native.FLAG_PAIR_REGS):
                for early in (False, True):
                    eng = MonteCarloEngine(edd, graph="alist", precision="f32_fast", max_iterations=4, seed=3,
                                           early_termination=early, kernel_flags=flags, fix_odd_check_sign=True)
                    eng.run_point(2.0, 0.5, frames=1301, interval_frames=700, random_codewords="frame")
            done.append(f"monte-carlo pair kernels {name}")
        torch.cuda.synchronize()
        print("sanitize_small: ok --", "; ".join(done))
        return
    for name, frames in [("bch_7_4", 70), ("wimax_576_0.5", 45), ("wimax_576_0.5.std", 33), ("ccsds_128_64", 40)]:
        code = load_code(name)
        llr = rng.normal(-1.0, 2.0, size=(frames, code.n))
        for prec in ("f64", "f32"):
            d = decoder(code, prec)
            d.decode_batch(llr, want_posterior=True, normalized_llr=True, compact=True)
            d.decode_batch(llr, early_termination=False, want_bits=True)
        done.append(f"generic {name}")
    for name, frames in [("wimax_576_0.5", 300), ("wimax_2304_0.83", 70), ("wifi_648_r083", 200), ("tanner_155_64", 333),
                         ("wimax_1152_0.66B", 150)]:
        code = load_code(name)
        llr = rng.normal(-1.0, 2.0, size=(frames, code.n)).astype(np.float32)
        d = decoder(code, "f32_fast")
        kind = d.graph.prepare("f32_fast")
        for early in (True, False):
            d.decode_batch(llr, early_termination=early, want_posterior=True, want_bits=True)
            if name != "wimax_1152_0.66B":
                d.decode_batch(llr, early_termination=early, table_kernel=True, jit=False)
        d.decode_batch_device(torch.as_tensor(llr[:, :]).cuda()[1:], early_termination=True)      # rows not 16-byte aligned for n=155
        done.append(f"resident {name} ({kind})")
    for name, graph, prec in [("wimax_576_0.5", "alist", "f32_fast"), ("wifi_648_r083", "alist", "f32_fast"),
                              ("bch_7_4", "std", "f64"), ("wimax_576_0.5", "std", "f32")]:
        edd = EncoderDecoderData(h=load_code(name).sparse_matrix())
        eng = MonteCarloEngine(edd, graph=graph, precision=prec, max_iterations=5, seed=3)
        for mode in ("frame", False):
            eng.run_point(2.0, 0.5, frames=200, interval_frames=100, random_codewords=mode)
        done.append(f"monte-carlo {name} {graph} {prec} ({eng.kernel})")
    torch.cuda.synchronize()
    print("sanitize_small: ok --", "; ".join(done))


if __name__ == "__main__":
    main()
