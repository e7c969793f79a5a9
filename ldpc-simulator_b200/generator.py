"""Bit source and the legacy LCG/Box-Muller generator.

``Generator.generate_bit_sequence`` is the reference's data source
(python_ldpc_app/generator.py:7-9).  ``ran``/``gauss`` restate the Park-Miller
minimal-standard LCG with Schrage's split and the Box-Muller step
(generator.py:15-32); the reference only uses them in channel modes 2/3 (out of
scope), but ``Channel.gen_ptr.sigma`` is part of the surface.  The Monte-Carlo
path on the GPU draws from Philox instead (csrc/awgn_philox.cuh).
"""
import math
import random

_LCG_A, _LCG_M = 16807, 2147483647
_LCG_Q, _LCG_R = 127773, 2836          # m = a*q + r


class Generator:
    def __init__(self, idum, sigma):
        self.idum = idum
        self.sigma = sigma

    @staticmethod
    def generate_bit_sequence(size):
        return [random.randint(0, 1) for _ in range(size)]

    def ran(self):
        hi, lo = divmod(self.idum, _LCG_Q)
        self.idum = _LCG_A * lo - _LCG_R * hi
        if self.idum < 0:
            self.idum += _LCG_M
        return self.idum * (1.0 / _LCG_M)

    def gauss(self, b):
        radius = self.sigma * math.sqrt(-2.0 * math.log(self.ran()))
        phase = 2.0 * math.pi * self.ran()
        return radius * (math.cos(phase) if b % 2 == 0 else math.sin(phase))
