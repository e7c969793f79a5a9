#!/usr/bin/env python3
"""SNR sweep driver: ``run_simulation`` and a CLI with the reference's flag names.

``run_simulation(encoder_decoder_data, settings, args, encoding_method, ru_data=None)`` keeps the
signature and the result semantics of python_ldpc_app/main.py:178-442 (SNR grid
``ceil((end-start)/step)+1`` points clamped to ``end``; per point ``args.blocks`` frames;
FER/BER/avg-convergence as in :357-369; ``SimulationResult`` as in :417-442), but every
frame is generated, decoded and counted on the GPU (mc_driver.MonteCarloEngine).
``args.threads`` is accepted and recorded; parallelism comes from the GPU batch and, when
``torch.distributed`` is initialised, from sharding frames over ranks.

Supported: channel modes 1 (AWGN, the north-star path), 2 and 3 (interference, channel.py:83-100; Philox
instead of the reference's Park-Miller stream), modulation 1 (BPSK) and the reference's "modulation 2"
(symbols of amplitude 0.7), standard encoding; other modes / encoders raise ``NotImplementedError``.  Interleaver settings are
recorded but are a statistical no-op on this memoryless channel.  ``--adaptive`` runs the sweep
under ``adaptive.AdaptiveController`` (main.py:620-639 of the reference).
"""
from __future__ import annotations

import argparse
import math
import os
import sys
import time
from datetime import datetime

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from encoder_decoder_data import EncoderDecoderData          # noqa: E402
from enums import EncodingMethod, InterleaverType, LDPCDecoderType, Result  # noqa: E402
from mc_driver import MonteCarloEngine                        # noqa: E402
from results import SimulationConfig, SimulationResult, SNRPointResult  # noqa: E402
from settings import Settings                                 # noqa: E402


def calculate_ber(original_data, decoded_data):
    if len(original_data) != len(decoded_data):
        return 1.0
    if not original_data:
        return 0.0
    return sum(a != b for a, b in zip(original_data, decoded_data)) / len(original_data)


def calculate_fer(decoding_result):
    return 0.0 if decoding_result == Result.OK else 1.0


def snr_grid(initial_snr, end_snr, step_snr):
    count = int(math.ceil((end_snr - initial_snr) / step_snr)) + 1
    return [min(initial_snr + i * step_snr, end_snr) for i in range(count)]


def _check_scope(settings, args, encoding_method):
    if getattr(args, "mode", 1) not in (1, 2, 3):
        raise ValueError("channel mode must be 1, 2 or 3")
    if encoding_method != EncodingMethod.STANDARD:
        raise NotImplementedError("only the standard (generator matrix) encoder is supported")
    # An interleaver setting is accepted and reported but moves no data: the channel is memoryless and
    # the reference de-interleaves with the same permutation before decoding (main.py:104-112), so no
    # statistic can depend on it.


def run_simulation(encoder_decoder_data, settings, args, encoding_method, ru_data=None):
    _check_scope(settings, args, encoding_method)
    started = time.time()
    quiet = bool(getattr(args, "quiet", False))
    say = (lambda *a, **k: None) if quiet else print

    engine = MonteCarloEngine(
        encoder_decoder_data,
        graph=getattr(args, "graph", "std"),
        precision=getattr(args, "precision", None) or getattr(settings, "get_precision", lambda: "f64")(),
        max_iterations=settings.get_max_iterations(),
        early_termination=getattr(settings, "is_early_termination", lambda: True)(),
        fix_odd_check_sign=getattr(settings, "is_fix_odd_check_sign", lambda: False)(),
        sigma_sq_quirk=not getattr(args, "no_sigma_sq_quirk", False),
        seed=getattr(args, "seed", None) if getattr(args, "seed", None) is not None else int(time.time() * 1e6) % (2 ** 63),
        normalized_llr=bool(getattr(args, "normalized_llr", False)),
        modulation=getattr(args, "modulation", 1), mode=getattr(args, "mode", 1), p=getattr(args, "p", 0.1),
        interference_snr=getattr(args, "interference_snr", 0.0) if getattr(args, "mode", 1) != 1 else 0.0,   # main.py:215
    )
    say("Processing blocks over the SNR grid...")
    say("-" * 60)
    k = encoder_decoder_data._k
    snr_points = []
    for current_snr in snr_grid(args.initial_snr, args.end_snr, args.step_snr):
        say(f"\nSNR: {current_snr:.2f} dB")
        say("-" * 60)
        cnt = engine.run_point(current_snr, args.speed, frames=args.blocks,
                               interval_frames=getattr(args, "interval_frames", None))
        avg_fer = cnt.frame_errors / args.blocks if args.fer else 0.0
        avg_ber = (cnt.bit_errors / (k * args.blocks) if k * args.blocks > 0 else 0.0) if args.ber else 0.0
        avg_norm = cnt.avg_normalized_llr(k) if getattr(args, "normalized_llr", False) else 0.0      # main.py:357
        if getattr(args, "normalized_llr", False):
            say(f"  Normalized LLR: {avg_norm:.6f}")
        if args.fer:
            say(f"  FER: {avg_fer:.6f}")
        if args.ber:
            say(f"  BER: {avg_ber:.6f}")
        ok_blocks = cnt.frames - cnt.frame_errors
        say(f"  Decoded successfully: {ok_blocks}/{args.blocks} ({100.0 * ok_blocks / args.blocks:.2f}%)")
        snr_points.append(SNRPointResult(
            snr_db=current_snr, ber=avg_ber, fer=avg_fer, avg_normalized_llr=avg_norm,
            total_blocks=args.blocks, successful_blocks=ok_blocks, failed_blocks=cnt.frame_errors,
            avg_convergence_iterations=cnt.avg_conv(), matrix_path=args.matrix,
            modulation=args.modulation, max_iterations=args.iterations, interleaver=args.interleaver,
            encoding_method=args.encoding_method))

    config = SimulationConfig(
        matrix_path=args.matrix, n=encoder_decoder_data._n, m=encoder_decoder_data._m, k=k,
        rate=encoder_decoder_data._rate, blocks=args.blocks, max_iterations=args.iterations,
        encoding_method=args.encoding_method, interleaver_type=args.interleaver, decoder_type=args.decoder,
        channel_mode=args.mode, modulation=args.modulation, speed=args.speed,
        snr_range=(args.initial_snr, args.end_snr, args.step_snr), threads=args.threads,
        timestamp=datetime.now().isoformat(), interference_snr=args.interference_snr, p=args.p)
    return SimulationResult(config=config, snr_points=snr_points, wall_clock_seconds=time.time() - started)


def build_parser():
    p = argparse.ArgumentParser(description="LDPC SPA Monte-Carlo simulation on B200 (reference-compatible flags)")
    p.add_argument("--matrix", "-m", type=str, required=True)
    p.add_argument("--blocks", "-b", type=int, default=100)
    p.add_argument("--iterations", "-i", type=int, default=5)
    p.add_argument("--interleaver", "-il", type=str, choices=["none", "regular", "random", "srandom"], default="none")
    p.add_argument("--decoder", "-d", type=str, choices=["bitflipping", "sumproduct"], default="sumproduct")
    p.add_argument("--speed", "-s", type=float, default=1.0)
    p.add_argument("--initial-snr", type=float, default=0.0)
    p.add_argument("--end-snr", type=float, default=5.0)
    p.add_argument("--step-snr", type=float, default=0.5)
    p.add_argument("--interference-snr", type=float, default=1.0)
    p.add_argument("--mode", type=int, choices=[1, 2, 3], default=1)
    p.add_argument("--p", type=float, default=0.1)
    p.add_argument("--modulation", "-mod", type=int, choices=[1, 2], default=1)
    p.add_argument("--ber", action="store_true")
    p.add_argument("--fer", action="store_true")
    p.add_argument("--normalized-llr", action="store_true")
    p.add_argument("--encoding-method", "-e", type=str, choices=["standard", "richardson-urbanke"], default="standard")
    p.add_argument("--threads", "-t", type=int, default=1)
    p.add_argument("--s-param", type=int, default=2, help="S of the S-random interleaver (recorded only)")
    p.add_argument("--ru-gap", type=int, default=None, help="accepted for compatibility; the Richardson-Urbanke encoder is out of scope")
    p.add_argument("--plot", action="store_true", help="accepted for compatibility: plot the written results with the reference's plot_results.py")
    p.add_argument("--plot-save", type=str, default=None)
    p.add_argument("--adaptive", action="store_true")
    p.add_argument("--adaptive-strategy", type=str, choices=["threshold"], default="threshold")
    p.add_argument("--matrix-dir", type=str, default=None)
    p.add_argument("--adaptive-high-ber", type=float, default=1e-2)
    p.add_argument("--adaptive-low-ber", type=float, default=1e-5)
    p.add_argument("--output-json", type=str, default=None)
    p.add_argument("--output-csv", type=str, default=None)
    # B200 additions
    p.add_argument("--precision", choices=["f64", "f32", "f32_fast"], default="f64")
    p.add_argument("--graph", choices=["std", "alist"], default="std",
                   help="std: decode on H_std as the reference does; alist: raw sparse H (quasi-cyclic fast path)")
    p.add_argument("--fix-odd-check-sign", action="store_true", help="NOT the reference behaviour, see DESIGN.md")
    p.add_argument("--no-sigma-sq-quirk", action="store_true", help="noise stddev = sigma instead of sigma^2")
    p.add_argument("--no-early-termination", action="store_true")
    p.add_argument("--seed", type=int, default=None)
    return p


def main(argv=None):
    args = build_parser().parse_args(argv)
    if not os.path.exists(args.matrix):
        print(f"Error: matrix file not found: {args.matrix}")
        return 1
    edd = EncoderDecoderData(args.matrix)
    st = Settings()
    st.set_blocks_cnt(args.blocks)
    st.set_max_iterations(args.iterations)
    st.set_decoder_type(LDPCDecoderType.SUM_PRODUCT)
    st.set_interleaver_type({"none": InterleaverType.NONE, "regular": InterleaverType.REGULAR,
                             "random": InterleaverType.RANDOM, "srandom": InterleaverType.SRANDOM}[args.interleaver])
    if args.interleaver == "srandom":
        st.set_s_param(args.s_param)
    st.set_ber_calculate(args.ber)
    st.set_fer_calculate(args.fer)
    st.set_normalized_llr_calculate(args.normalized_llr)
    st.set_precision(args.precision)
    st.set_early_termination(not args.no_early_termination)
    st.set_fix_odd_check_sign(args.fix_odd_check_sign)
    method = EncodingMethod.STANDARD if args.encoding_method == "standard" else EncodingMethod.RICHARDSON_URBANKE
    if args.adaptive:                                                      # main.py:620-639
        from adaptive import AdaptiveController, ThresholdStrategy
        from matrix_catalog import MatrixCatalog
        matrix_dir = args.matrix_dir or os.path.join(os.path.dirname(os.path.abspath(args.matrix)), "..")
        strategy = ThresholdStrategy(high_ber_threshold=args.adaptive_high_ber, low_ber_threshold=args.adaptive_low_ber)
        result = AdaptiveController(strategy, MatrixCatalog(matrix_dir)).run_adaptive_sweep(edd, st, args, method)
    else:
        result = run_simulation(edd, st, args, method)
    if args.plot or args.plot_save:
        print("note: plots are not drawn here; feed the results file to the reference's plot_results.py / visualization.py "
              "(the JSON / CSV written by --output-json / --output-csv is byte-compatible)")
    if args.output_json:
        result.to_json(args.output_json)
    if args.output_csv:
        result.to_csv(args.output_csv)
    return 0


if __name__ == "__main__":
    sys.exit(main())
