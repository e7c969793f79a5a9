"""Run settings object read by the decoder and the Monte-Carlo driver.

Same getters/setters and defaults as python_ldpc_app/settings.py:4-89 (the decoder
reads ``get_max_iterations`` and ``is_normalized_llr_calculate``,
spa_decoder.py:104,210); three B200-only knobs are added with defaults that keep
the reference behaviour.
"""
from enums import InterleaverType, LDPCDecoderType

_INTERLEAVER_NAMES = {InterleaverType.REGULAR: "Regular", InterleaverType.RANDOM: "Random",
                      InterleaverType.SRANDOM: "S-Random"}


class Settings:
    def __init__(self):
        self._i_blocks_cnt = 100
        self._max_iterations = 5
        self._interleaver_type = InterleaverType.NONE
        self._decoder_type = LDPCDecoderType.BIT_FLIPPING
        self._b_ber_calculate = True
        self._b_fer_calculate = False
        self._b_is_calculate_normalized_llr = False
        self._d_s_param = -1
        # B200 additions (not in the reference)
        self._precision = "f64"            # "f64" parity kernel | "f32" | "f32_fast" resident QC kernel
        self._early_termination = True     # reference behaviour (spa_decoder.py:231-241)
        self._fix_odd_check_sign = False   # reference behaviour: do NOT compensate
        self._one_frame_kernel = False     # resident path: force the one-frame-per-thread kernel (same results)
        self._pair_regs_kernel = False     # resident path: pair kernel with the messages in registers, not TMEM
        self._pair_scatter_kernel = False  # resident path: pair kernel with in-place posterior accumulation (TMEM messages)
        self._pair_gather_kernel = False   # resident path: gather pair kernel also with early termination
        self._one_gather_kernel = False    # resident path: one-frame gather kernel (tensor-memory messages, 4 CTAs per SM)

    # -- reference surface ----------------------------------------------------
    def set_blocks_cnt(self, i_num_blocks): self._i_blocks_cnt = i_num_blocks
    def get_blocks_cnt(self): return self._i_blocks_cnt
    def set_max_iterations(self, i_max_iter): self._max_iterations = i_max_iter
    def get_max_iterations(self): return self._max_iterations
    def set_interleaver_type(self, e_int_type): self._interleaver_type = e_int_type
    def get_interleaver_type(self): return self._interleaver_type
    def set_decoder_type(self, e_decoder_type): self._decoder_type = e_decoder_type
    def get_decoder_type(self): return self._decoder_type
    def set_ber_calculate(self, flag): self._b_ber_calculate = flag
    def is_ber_calculate(self): return self._b_ber_calculate
    def set_fer_calculate(self, flag): self._b_fer_calculate = flag
    def is_fer_calculate(self): return self._b_fer_calculate
    def set_normalized_llr_calculate(self, flag): self._b_is_calculate_normalized_llr = flag
    def is_normalized_llr_calculate(self): return self._b_is_calculate_normalized_llr
    def set_s_param(self, i_s_param): self._d_s_param = i_s_param
    def get_s_param(self): return self._d_s_param

    def get_interleaver_type_name(self):
        return _INTERLEAVER_NAMES.get(self._interleaver_type, "None")

    def print(self):
        print(f"Block count: {self._i_blocks_cnt}")
        print(f"Interleaver type: {self.get_interleaver_type_name()};")
        if self._decoder_type == LDPCDecoderType.BIT_FLIPPING:
            print("Decoder type: Bit-flipped algorithm;")
        elif self._decoder_type == LDPCDecoderType.SUM_PRODUCT:
            print("Decoder type: Sum-product algorithm;")

    # -- B200 additions ---------------------------------------------------------
    def set_precision(self, name):
        if name not in ("f64", "f32", "f32_fast"):
            raise ValueError("precision must be 'f64', 'f32' or 'f32_fast'")
        self._precision = name

    def get_precision(self): return self._precision
    def set_early_termination(self, flag): self._early_termination = bool(flag)
    def is_early_termination(self): return self._early_termination
    def set_fix_odd_check_sign(self, flag): self._fix_odd_check_sign = bool(flag)
    def is_fix_odd_check_sign(self): return self._fix_odd_check_sign
    def set_one_frame_kernel(self, flag): self._one_frame_kernel = bool(flag)
    def is_one_frame_kernel(self): return self._one_frame_kernel
    def set_pair_regs_kernel(self, flag): self._pair_regs_kernel = bool(flag)
    def is_pair_regs_kernel(self): return self._pair_regs_kernel
    def set_pair_scatter_kernel(self, flag): self._pair_scatter_kernel = bool(flag)
    def is_pair_scatter_kernel(self): return self._pair_scatter_kernel
    def set_pair_gather_kernel(self, flag): self._pair_gather_kernel = bool(flag)
    def is_pair_gather_kernel(self): return self._pair_gather_kernel
    def set_one_gather_kernel(self, flag): self._one_gather_kernel = bool(flag)
    def is_one_gather_kernel(self): return self._one_gather_kernel
