"""Adaptive parameter selection between SNR points, on top of the GPU Monte-Carlo engine.

Mirrors the public surface of python_ldpc_app/adaptive.py (``AdaptiveState``, ``AdaptiveAction``,
``AdaptiveStrategy``, ``ThresholdStrategy``, ``AdaptiveController.run_adaptive_sweep``) so that
``main.py --adaptive`` and callers of the reference keep working:

* the policy is the reference's (:61-124): BER above ``high_ber_threshold`` asks for a lower-rate code,
  BER in (0, ``low_ber_threshold``) for a higher-rate one, an average convergence iteration above
  ``convergence_ratio * max_iterations`` doubles the iteration budget (capped at 100), FER above
  ``fer_threshold`` with no interleaver switches the random interleaver on;
* rate changes are resolved against a ``MatrixCatalog`` (same family and block size, :384-402);
* the result is a ``SimulationResult`` whose ``adaptation_log`` holds the state every point ran with
  (:197-205) and whose SNR points carry the per-point matrix / iterations / interleaver (:327-341).

What differs: the frames of a point are not pushed through ``process_block`` workers (:231-317) but
through ``mc_driver.MonteCarloEngine`` (one engine per (matrix, iteration budget), kept for the sweep).
The interleaver setting is tracked and reported but does not touch the data: the channel is memoryless
and the reference de-interleaves with the same permutation before decoding (main.py:104-112), so it
cannot change any statistic.
"""
from __future__ import annotations

import os
import time
from abc import ABC, abstractmethod
from dataclasses import dataclass, field
from datetime import datetime
from typing import List, Optional

from encoder_decoder_data import EncoderDecoderData
from enums import InterleaverType, LDPCDecoderType
from matrix_catalog import MatrixCatalog, MatrixInfo
from results import SimulationConfig, SimulationResult, SNRPointResult
from settings import Settings

LOWER_RATE = "__LOWER_RATE__"      # sentinels resolved by the controller against the catalog (:88,95)
HIGHER_RATE = "__HIGHER_RATE__"
MAX_ITERATION_BUDGET = 100          # :105


@dataclass
class AdaptiveState:
    current_matrix_path: str
    current_rate: float
    current_modulation: int
    current_max_iterations: int
    current_interleaver: str
    current_encoding_method: str
    history: List[dict] = field(default_factory=list)


@dataclass
class AdaptiveAction:
    new_matrix_path: Optional[str] = None
    new_modulation: Optional[int] = None
    new_max_iterations: Optional[int] = None
    new_interleaver: Optional[str] = None
    reason: str = ""


class AdaptiveStrategy(ABC):
    @abstractmethod
    def evaluate(self, state: AdaptiveState, last_snr_result: SNRPointResult) -> Optional[AdaptiveAction]:
        """None = keep everything, otherwise what to change before the next SNR point."""

    @abstractmethod
    def get_name(self) -> str:
        ...


class ThresholdStrategy(AdaptiveStrategy):
    def __init__(self, high_ber_threshold=1e-2, low_ber_threshold=1e-5, fer_threshold=0.5, convergence_ratio=0.8):
        self.high_ber_threshold = high_ber_threshold
        self.low_ber_threshold = low_ber_threshold
        self.fer_threshold = fer_threshold
        self.convergence_ratio = convergence_ratio

    def get_name(self) -> str:
        return "threshold"

    def evaluate(self, state, last_snr_result):
        r = last_snr_result
        act = AdaptiveAction()
        why = []
        if r.ber > self.high_ber_threshold:                                   # :86-91
            act.new_matrix_path = LOWER_RATE
            why.append(f"BER={r.ber:.2e} > {self.high_ber_threshold:.2e}, switching to lower rate")
        elif 0 < r.ber < self.low_ber_threshold:                              # :94-98 (a zero BER proves nothing)
            act.new_matrix_path = HIGHER_RATE
            why.append(f"BER={r.ber:.2e} < {self.low_ber_threshold:.2e}, switching to higher rate")
        if r.avg_convergence_iterations > self.convergence_ratio * state.current_max_iterations:    # :101-110
            grown = min(2 * state.current_max_iterations, MAX_ITERATION_BUDGET)
            if grown > state.current_max_iterations:
                act.new_max_iterations = grown
                why.append(f"avg_conv={r.avg_convergence_iterations:.1f} near max={state.current_max_iterations}, "
                           f"increasing to {grown}")
        if r.fer > self.fer_threshold and state.current_interleaver == "none":                     # :113-118
            act.new_interleaver = "random"
            why.append(f"FER={r.fer:.3f} > {self.fer_threshold}, enabling random interleaver")
        if not why:
            return None
        act.reason = "; ".join(why)
        return act


_INTERLEAVERS = {"none": InterleaverType.NONE, "regular": InterleaverType.REGULAR,
                 "random": InterleaverType.RANDOM, "srandom": InterleaverType.SRANDOM}


class AdaptiveController:
    def __init__(self, strategy: AdaptiveStrategy, catalog: MatrixCatalog):
        self.strategy = strategy
        self.catalog = catalog
        self._encoder_cache = {}        # matrix path -> EncoderDecoderData
        self._engines = {}              # (matrix path, max iterations) -> MonteCarloEngine

    # ---- resources -------------------------------------------------------------------------
    def _get_encoder_decoder_data(self, matrix_path: str) -> EncoderDecoderData:
        if matrix_path not in self._encoder_cache:
            print(f"  [Adaptive] loading matrix: {os.path.basename(matrix_path)}")
            self._encoder_cache[matrix_path] = EncoderDecoderData(matrix_path)
        return self._encoder_cache[matrix_path]

    def _find_current_matrix_info(self, matrix_path: str) -> Optional[MatrixInfo]:
        want = os.path.abspath(matrix_path)
        return next((m for m in self.catalog.matrices if os.path.abspath(m.path) == want), None)

    def _engine(self, state, settings, args):
        from mc_driver import MonteCarloEngine
        key = (state.current_matrix_path, state.current_max_iterations, state.current_modulation)
        if key not in self._engines:
            seed = getattr(args, "seed", None)
            self._engines[key] = MonteCarloEngine(
                self._get_encoder_decoder_data(state.current_matrix_path),
                graph=getattr(args, "graph", "std"),
                precision=getattr(args, "precision", None) or getattr(settings, "get_precision", lambda: "f64")(),
                max_iterations=state.current_max_iterations,
                early_termination=getattr(settings, "is_early_termination", lambda: True)(),
                fix_odd_check_sign=getattr(settings, "is_fix_odd_check_sign", lambda: False)(),
                sigma_sq_quirk=not getattr(args, "no_sigma_sq_quirk", False),
                seed=seed if seed is not None else int(time.time() * 1e6) % (2 ** 63),
                normalized_llr=bool(getattr(args, "normalized_llr", False)),
                modulation=state.current_modulation, mode=getattr(args, "mode", 1), p=getattr(args, "p", 0.1),
                interference_snr=getattr(args, "interference_snr", 0.0) if getattr(args, "mode", 1) != 1 else 0.0)
        return self._engines[key]

    # ---- the sweep ---------------------------------------------------------------------------
    def run_adaptive_sweep(self, encoder_decoder_data, settings, args, encoding_method, ru_data=None):
        from main import _check_scope, snr_grid
        _check_scope(settings, args, encoding_method)
        started = time.time()
        self._encoder_cache[args.matrix] = encoder_decoder_data
        state = AdaptiveState(
            current_matrix_path=args.matrix, current_rate=encoder_decoder_data._rate,
            current_modulation=args.modulation, current_max_iterations=args.iterations,
            current_interleaver=args.interleaver, current_encoding_method=args.encoding_method)
        current_settings = settings
        snr_points, adaptation_log = [], []
        want_norm = bool(getattr(args, "normalized_llr", False))
        print("Processing blocks over the SNR grid (adaptive mode)...")
        for current_snr in snr_grid(args.initial_snr, args.end_snr, args.step_snr):
            print(f"\nSNR: {current_snr:.2f} dB  [rate={state.current_rate:.3f}, iters={state.current_max_iterations}, "
                  f"interleaver={state.current_interleaver}]")
            adaptation_log.append({                                                        # :197-205
                "snr_db": current_snr, "matrix_path": state.current_matrix_path, "rate": state.current_rate,
                "modulation": state.current_modulation, "max_iterations": state.current_max_iterations,
                "interleaver": state.current_interleaver, "encoding_method": state.current_encoding_method})
            edd = self._get_encoder_decoder_data(state.current_matrix_path)
            cnt = self._engine(state, current_settings, args).run_point(
                current_snr, args.speed, frames=args.blocks, interval_frames=getattr(args, "interval_frames", None))
            k = edd._k
            avg_fer = cnt.frame_errors / args.blocks if args.fer else 0.0                  # :320-325
            avg_ber = cnt.bit_errors / (k * args.blocks) if (args.ber and k * args.blocks > 0) else 0.0
            ok_blocks = cnt.frames - cnt.frame_errors
            point = SNRPointResult(
                snr_db=current_snr, ber=avg_ber, fer=avg_fer,
                avg_normalized_llr=cnt.avg_normalized_llr(k) if want_norm else 0.0,
                total_blocks=args.blocks, successful_blocks=ok_blocks, failed_blocks=cnt.frame_errors,
                avg_convergence_iterations=cnt.avg_conv(), matrix_path=state.current_matrix_path,
                modulation=state.current_modulation, max_iterations=state.current_max_iterations,
                interleaver=state.current_interleaver, encoding_method=state.current_encoding_method)
            snr_points.append(point)
            print(f"  FER: {avg_fer:.6f}  BER: {avg_ber:.6f}  decoded {ok_blocks}/{args.blocks}")
            action = self.strategy.evaluate(state, point)
            if action:
                print(f"  [Adaptive] {action.reason}")
                self._apply_action(action, state, edd, current_settings, args)
                current_settings = self._build_settings(state, args, like=settings)

        base = encoder_decoder_data
        config = SimulationConfig(                                                         # :419-438: the INITIAL set-up
            matrix_path=args.matrix, n=base._n, m=base._m, k=base._k, rate=base._rate, blocks=args.blocks,
            max_iterations=args.iterations, encoding_method=args.encoding_method, interleaver_type=args.interleaver,
            decoder_type=args.decoder, channel_mode=args.mode, modulation=args.modulation, speed=args.speed,
            snr_range=(args.initial_snr, args.end_snr, args.step_snr), threads=args.threads,
            timestamp=datetime.now().isoformat(), interference_snr=args.interference_snr, p=args.p)
        return SimulationResult(config=config, snr_points=snr_points, wall_clock_seconds=time.time() - started,
                                adaptation_log=adaptation_log)

    def _apply_action(self, action, state, current_encoder_data, current_settings, args):
        here = self._find_current_matrix_info(state.current_matrix_path)
        if here is not None and action.new_matrix_path in (LOWER_RATE, HIGHER_RATE):
            pick = self.catalog.get_lower_rate if action.new_matrix_path == LOWER_RATE else self.catalog.get_higher_rate
            other = pick(here)
            if other is not None:                      # at the end of the family's rate ladder nothing changes
                state.current_matrix_path, state.current_rate = other.path, other.rate
                self._get_encoder_decoder_data(other.path)
                print(f"  [Adaptive] matrix: {other.name} (rate={other.rate:.3f})")
        if action.new_max_iterations is not None:
            state.current_max_iterations = action.new_max_iterations
        if action.new_modulation is not None:
            state.current_modulation = action.new_modulation
        if action.new_interleaver is not None:
            state.current_interleaver = action.new_interleaver

    def _build_settings(self, state, args, like=None) -> Settings:
        s = Settings()
        s.set_blocks_cnt(args.blocks)
        s.set_max_iterations(state.current_max_iterations)
        s.set_interleaver_type(_INTERLEAVERS.get(state.current_interleaver, InterleaverType.NONE))
        if state.current_interleaver == "srandom":
            s.set_s_param(getattr(args, "s_param", 0))
        s.set_decoder_type(LDPCDecoderType.SUM_PRODUCT)
        s.set_ber_calculate(args.ber)
        s.set_fer_calculate(args.fer)
        s.set_normalized_llr_calculate(getattr(args, "normalized_llr", False))
        if like is not None:                            # B200 additions ride along
            for name in ("precision", "early_termination", "fix_odd_check_sign"):
                getter = getattr(like, ("get_" if name == "precision" else "is_") + name, None)
                if getter is not None:
                    getattr(s, "set_" + name)(getter())
        return s
