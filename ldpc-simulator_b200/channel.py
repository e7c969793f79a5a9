"""Channel -> LLRs (the decoder's input interface).

Mirrors ``Channel.create_channel`` / ``Channel.process`` of python_ldpc_app/channel.py
(:102-125, :38-81), including the conventions the error-rate curves depend on:

* bit 0 -> symbol -1, bit 1 -> +1 (:49) -- or -+0.7 for "modulation 2" (:50-51);
  LLR = 2 y / sigma^2 (:80), so LLR < 0 means bit 0;
* mode 1 (AWGN): the noise standard deviation is sigma**2, not sigma (:68) -- kept (``sigma_sq_quirk``);
* modes 2 (partial-band interference, :83-96) and 3 (counter interference, :98-100) draw their noise from
  the two Park-Miller / Box-Muller generators (``gen_ptr``, ``gen_ptr2``) exactly as the reference does
  -- a deterministic stream, so ``process`` is bit-identical to the reference there (mode 2 also
  consumes numpy's global RandomState for the hit decision, :86).  Host only: they are in no
  benchmark configuration, the GPU generator covers mode 1.

``sigma = 1/sqrt(2 * speed * 10^(snr/10))`` (:113): ``speed`` plays the role of the code rate.

B200 additions: ``process_batch`` (host, vectorised, mode 1, for the parity mode with host-fed
LLRs) and ``device_llr`` (Philox generator on the GPU, csrc/awgn_philox.cuh).
"""
from __future__ import annotations

import math
import time

import numpy as np

import _native
from constants import IDUM1, IDUM2
from generator import Generator


class Channel:
    def __init__(self, mode, p, mod, L_c1, L_c2, L_c3):
        self.mode, self.p, self.modulation = mode, p, mod
        self.L_c1, self.L_c2, self.L_c3 = L_c1, L_c2, L_c3
        self.gen_ptr = None
        self.gen_ptr2 = None
        self.sigma_sq_quirk = True
        # the reference seeds from the clock (:30): runs are not reproducible unless reseeded
        self._rng = np.random.RandomState(int(time.time() * 1e6) % (2 ** 31))

    def seed(self, value):
        self._rng = np.random.RandomState(int(value) % (2 ** 31))

    def _require_awgn(self):
        if self.mode != 1:
            raise NotImplementedError("process_batch covers mode 1 (AWGN); modes 2 and 3 go through process() or device_llr()")

    def describe(self, speed=None, snr_db=None):
        """The ``ldpc_channel`` struct of this channel for the device generator."""
        return _native.ChannelDesc(
            int(self.mode), int(self.modulation), int(bool(self.sigma_sq_quirk)),
            float(self._speed if speed is None else speed), float(self._snr_db if snr_db is None else snr_db),
            float(getattr(self, "_snr2_db", 0.0)), float(self.p))

    @property
    def amplitude(self):
        return 0.7 if self.modulation == 2 else 1.0         # :49-51 (any other value leaves the symbol at 0, as there)

    def _noise_dev(self):
        s = self.gen_ptr.sigma
        return s * s if self.sigma_sq_quirk else s

    def process(self, data_buffer):
        """Append n LLRs to ``data_buffer._channel_data`` (a reused buffer grows, as in the reference)."""
        bits = np.asarray(data_buffer._encoded_data)
        if bits.size == 0:
            return
        if self.mode == 1:
            llr = self.process_batch(bits[None, :])[0]
            data_buffer._channel_data.extend(float(v) for v in llr)
            return
        n = int(bits.size)
        amp = self.amplitude if self.modulation in (1, 2) else 0.0
        out = data_buffer._channel_data
        for i in range(n):
            bit = -amp if bits[i] == 0 else amp
            if self.mode == 2:                                           # :83-96
                hit = np.random.randint(0, n) / n < self.p               # numpy's global stream, as the reference
                pom1 = self.gen_ptr.gauss(i)
                if hit:
                    out.append((bit + self.gen_ptr2.gauss(i) + pom1) * self.L_c2)
                else:
                    out.append((bit + pom1) * self.L_c1)
            elif self.mode == 3:                                         # :98-100
                pom1 = self.gen_ptr.gauss(i)
                pom2 = self.gen_ptr2.gauss(i)
                out.append(((bit + pom1 + pom2) * self.p + (bit + pom1) * (1 - self.p)) * self.L_c3)

    def process_batch(self, encoded):
        """encoded [F, n] bits -> LLRs [F, n] float64 (host, mode 1)."""
        self._require_awgn()
        enc = np.asarray(encoded)
        sym = np.where(enc == 0, -self.amplitude, self.amplitude)
        noise = self._rng.normal(0.0, self._noise_dev(), size=enc.shape)
        return 2.0 * (sym + noise) / (self.gen_ptr.sigma ** 2)

    def device_llr(self, frames, n, *, seed, stream_id=0, frame_offset=0, codeword=None, dtype="f32",
                   speed=None, snr_db=None):
        """LLRs [frames, n] generated on the GPU by the same Philox stream ``ldpc_mc_run`` decodes."""
        import ctypes as C
        import torch
        import _native
        desc = self.describe(speed, snr_db)
        tdt = torch.float64 if dtype == "f64" else torch.float32
        out = torch.empty((frames, n), dtype=tdt, device="cuda")
        cw, stride = None, 0
        if codeword is not None:          # [n]: one codeword for every frame; [frames, n]: one per frame
            cw = (codeword if hasattr(codeword, "data_ptr") else torch.as_tensor(np.asarray(codeword, dtype=np.uint8)))
            cw = cw.to(device="cuda", dtype=torch.uint8).contiguous()
            if cw.dim() == 2:
                if tuple(cw.shape) != (frames, n):
                    raise ValueError(f"per-frame codewords must be [{frames}, {n}]")
                stride = n
        _native.check(_native.lib().ldpc_channel_llr_ex(
            n, _native.LDPC_F64 if dtype == "f64" else _native.LDPC_F32, frames, C.byref(desc),
            int(seed), int(stream_id), int(frame_offset),
            cw.data_ptr() if cw is not None else None, stride, out.data_ptr(),
            torch.cuda.current_stream().cuda_stream))
        return out

    @staticmethod
    def sigma_for(speed, snr_db):
        return 1.0 / math.sqrt(2.0 * speed * (10.0 ** (snr_db * 0.1)))

    @staticmethod
    def create_channel(speed, sn1, sn2, mode, p, mod):
        lin1, lin2 = 10.0 ** (sn1 * 0.1), 10.0 ** (sn2 * 0.1)
        L_c1 = 4.0 * speed * lin1
        L_c2 = 4.0 * speed / (1.0 / lin1 + 1.0 / (lin2 * p)) if p else 0.0
        L_c3 = 4.0 * p * speed / (2.0 / lin2) + 4.0 * speed * (1.0 - p) * lin1
        sigma1 = Channel.sigma_for(speed, sn1) if mode in (1, 2, 3) else 0.0
        sigma2 = 0.0
        if mode == 2 and p:
            sigma2 = 1.0 / math.sqrt(2.0 * speed * lin2 * p)
        elif mode == 3:
            sigma2 = Channel.sigma_for(speed, sn2)
        ch = Channel(mode, p, mod, L_c1, L_c2, L_c3)
        ch.gen_ptr = Generator(IDUM1, sigma1)
        ch.gen_ptr2 = Generator(IDUM2, sigma2)
        ch._speed, ch._snr_db, ch._snr2_db = speed, sn1, sn2
        return ch
