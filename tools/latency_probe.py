"""Latency of the reference's per-frame entry point SPA_Decoder.decode(buf) (host buffers in and out).

    python tools/latency_probe.py

One frame per call, 20 iterations, early termination as in the reference.  Prints the median and the
launch count per call for a few (code, graph, precision) combinations.
"""
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
    import _native
    from conftest import load_code
    from settings import Settings
    from spa_decoder import SPA_Decoder

    class Edd:
        def __init__(self, h):
            self._h_sparse_cached, (self._m, self._n) = h, h.shape

    class Buf:
        pass

    rng = np.random.default_rng(0)
    for name, prec in [("bch_7_4.std", "f64"), ("wimax_576_0.5.std", "f64"), ("wimax_576_0.5", "f64"),
                       ("wimax_576_0.5", "f32_fast"), ("wimax_2304_0.5", "f32_fast"), ("wimax_2304_0.5.std", "f32")]:
        code = load_code(name)
        st = Settings()
        st.set_max_iterations(20)
        st.set_precision(prec)
        dec = SPA_Decoder(Edd(code.csr()), st)
        sig = 1.0 / np.sqrt(10 ** 0.2)
        times = []
        l0 = None
        for i in range(60):
            buf = Buf()
            buf._channel_data = list(2.0 * (-1.0 + sig * rng.standard_normal(code.n)) / sig ** 2)
            buf._decoded_data = []
            if i == 10:
                l0 = _native.launches()
            t0 = time.perf_counter()
            dec.decode(buf)
            times.append(time.perf_counter() - t0)
        launches = (_native.launches() - l0) / 50
        t = np.array(times[10:]) * 1e3
        print(json.dumps({"code": name, "precision": prec, "edges": int(code.nnz), "median_ms": round(float(np.median(t)), 3),
                          "p90_ms": round(float(np.percentile(t, 90)), 3), "launches_per_call": launches}), flush=True)


if __name__ == "__main__":
    main()
