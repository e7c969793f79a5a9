"""Device-resident decode throughput of one code on each kernel family that can run it.

    python tools/kernel_family_bench.py [code fixture names ...]      (tests/golden/codes/<name>.npz)

20 fixed iterations, LLRs already in HBM (Philox channel at 2 dB), CUDA events around 5 launches after 2
warm-ups; prints one JSON line per (code, family).  Families: the kernel specialised for the base matrix
(registered at build time or compiled with NVRTC), the table-driven resident kernel, the generic fp32
streaming kernels.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "ldpc-simulator_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import _native
    from channel import Channel
    from conftest import load_code
    from settings import Settings
    from spa_decoder import SPA_Decoder

    class Edd:
        def __init__(self, h):
            self._h_sparse_cached, (self._m, self._n) = h, h.shape

    names = sys.argv[1:] or ["wimax_2304_0.5", "wimax_1152_0.66B", "wifi_648_r083", "wimax_2304_0.83", "tanner_155_64"]
    for name in names:
        code = load_code(name)
        st = Settings()
        st.set_max_iterations(20)
        st.set_precision("f32_fast")
        dec = SPA_Decoder(Edd(code.csr()), st)
        k = code.n - code.m
        frames = max(4096, min(262144, (256 << 20) // (4 * code.n)))
        llr = Channel.create_channel(k / code.n, 2.0, 0.0, 1, 0.1, 1).device_llr(frames, code.n, seed=1)
        variants = [("specialised", dict()), ("table", dict(jit=False, table_kernel=True)),
                    ("generic_f32_fast", dict(force_generic=True)), ("generic_f32", dict(force_generic=True, precision="f32"))]
        for label, kw in variants:
            flags = (_native.FLAG_TABLE_KERNEL | _native.FLAG_NO_JIT) if label == "table" else \
                    (_native.FLAG_FORCE_GENERIC if label.startswith("generic") else 0)
            family = dec.graph.prepare("f32_fast", flags)
            if label == "table" and family != "qc_table":
                print(json.dumps({"code": name, "variant": label, "family": None, "note": "no table-driven shape for this base matrix"}))
                continue
            f = frames if not label.startswith("generic") else min(frames, 32768)
            x = llr[:f]
            ws = torch.empty(max(256, int(_native.lib().ldpc_workspace_bytes(dec.graph.handle, f, 1 if label.startswith("generic") else 2))),
                             dtype=torch.uint8, device="cuda")
            run = lambda: dec.decode_batch_device(x, early_termination=False, workspace=ws, **kw)
            for _ in range(2):
                run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5
            e0.record()
            for _ in range(reps):
                run()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            print(json.dumps({"code": name, "n": code.n, "k": k, "edges": int(code.nnz), "variant": label, "family": family,
                              "frames": f, "ms": round(ms, 3), "info_gbit_s": round(f * k / ms / 1e6, 3),
                              "edge_updates_per_s": round(f * code.nnz * 20 / ms * 1e3, 0)}), flush=True)


if __name__ == "__main__":
    main()
