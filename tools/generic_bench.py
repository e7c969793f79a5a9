"""Throughput of the generic (HBM-streaming) kernels on one graph.

    python tools/generic_bench.py <code fixture> <f64|f32> <frames> [reps]

20 fixed passes, LLRs in HBM.  Prints ms per decode, edge updates/s and the algorithmic HBM traffic
(per pass and frame 3 message sweeps + the n-vectors) over the elapsed time.
"""
import json
import os
import sys

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

    name, prec, frames = sys.argv[1], sys.argv[2], int(sys.argv[3])
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
    code = load_code(name)
    st = Settings()
    st.set_max_iterations(20)
    st.set_precision(prec)
    dec = SPA_Decoder(Edd(code.csr()), st)
    llr = Channel.create_channel(0.5, 2.0, 0.0, 1, 0.1, 1).device_llr(frames, code.n, seed=1, dtype=prec)
    dt = _native.LDPC_F64 if prec == "f64" else _native.LDPC_F32
    ws = torch.empty(int(_native.lib().ldpc_workspace_bytes(dec.graph.handle, frames, dt)), dtype=torch.uint8, device="cuda")
    run = lambda: dec.decode_batch_device(llr, early_termination=False, workspace=ws, force_generic=True)
    run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    esz = 8 if prec == "f64" else 4
    E, n = int(code.nnz), code.n
    bytes_alg = frames * 20 * (3 * E * esz + 3 * n * esz + 2 * n)
    print(json.dumps({"code": name, "precision": prec, "frames": frames, "edges": E, "ms": round(ms, 3),
                      "edge_updates_per_s": round(frames * E * 20 / ms * 1e3), "algorithmic_GBps": round(bytes_alg / ms / 1e6, 1),
                      "workspace_GB": round(ws.numel() / 1e9, 2)}))


if __name__ == "__main__":
    main()
