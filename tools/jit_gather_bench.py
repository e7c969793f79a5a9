"""Run-time specialised kernels on WiMAX rate-1/2 codes that are NOT in the build-time registry (802.16e scales the
n = 2304 base matrix: shift = floor(p z / 96)): gather kernel (two frames per thread) against the one-frame kernel of the
same NVRTC module.

    python tools/jit_gather_bench.py [z ...]        default z = 48 80 92

20 fixed iterations, LLRs already in HBM (Philox channel at 2 dB), CUDA events around 5 launches after 2 warm-ups; one JSON
line per (z, kernel).  The first call of a code pays the NVRTC compilation (reported as compile_s; cached on disk).
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "ldpc-simulator_b200"))


def main():

    import torch
    from scipy import sparse
    import _native
    from channel import Channel
    from settings import Settings
    from spa_decoder import SPA_Decoder

    class Edd:
        def __init__(self, h):
            self._h_sparse_cached, (self._m, self._n) = h, h.shape

    base = np.array(json.load(open(os.path.join(ROOT, "ldpc-simulator_b200", "csrc", "qc_registry.json")))[0]["shift"])
    mb, nb = base.shape
    for z in [int(a) for a in sys.argv[1:]] or [48, 80, 92]:
        shift = np.where(base < 0, -1, base * z // 96)
        rows, cols = [], []
        for a in range(mb):
            for c in range(nb):
                if shift[a, c] >= 0:
                    r = np.arange(z)
                    rows.append(a * z + r)
                    cols.append(c * z + (r + shift[a, c]) % z)
        rows, cols = np.concatenate(rows), np.concatenate(cols)
        h = sparse.csr_matrix((np.ones(rows.size, dtype=np.int32), (rows, cols)), shape=(mb * z, nb * z))
        h.sort_indices()
        n, k = nb * z, (nb - mb) * z
        frames = min(262144, (512 << 20) // (4 * n)) // 1184 * 1184
        llr = Channel.create_channel(k / n, 2.0, 0.0, 1, 0.1, 1).device_llr(frames, n, seed=1)
        for label, one in (("gather (two frames per thread)", False), ("one frame per thread", True)):
            st = Settings()
            st.set_max_iterations(20)
            st.set_precision("f32_fast")
            if one:
                st.set_one_frame_kernel(True)
            dec = SPA_Decoder(Edd(h), st)
            t0 = time.time()
            family = dec.graph.prepare("f32_fast")
            compile_s = time.time() - t0
            ws = torch.empty(4096, dtype=torch.uint8, device="cuda")
            run = lambda: dec.decode_batch_device(llr, early_termination=False, workspace=ws)
            for _ in range(2):
                run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                run()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print(json.dumps({"code": f"wimax r1/2 z={z} (n={n})", "family": family, "kernel": label, "frames": frames, "ms": round(ms, 3),
                              "info_gbit_s": round(frames * k / ms / 1e6, 3), "compile_s": round(compile_s, 1)}), flush=True)


if __name__ == "__main__":
    main()
