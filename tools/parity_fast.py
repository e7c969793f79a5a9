"""Measured parity of the throughput path (LDPC_F32_FAST, SM-resident kernel) against the fp64 oracle.

    python tools/parity_fast.py [--out profiles/r2_parity_fast.json] [--bench-frames 65536] [--frames 16384]
                                [--trace-frames 512] [--regimes bench,r083,...] [--one-frame]

The oracle (oracle/spa_oracle.c, pinned to the unmodified reference by tests/golden) is the checker; this
tool is measurement infrastructure like tests/ and may import it.  Both decoders get the same
float32-representable LLRs.  Per regime it records, over all frames:
    frame_agree            hard decisions of all n bits AND syndrome result AND iteration-at-convergence equal
    frame_agree_converged  the same, over the frames the ORACLE converges on
    bit_agree, ok_agree, conv_agree
    post_*                 |posterior - oracle| / max(|oracle|, 1): median / p99 / max, and the share of frames
                           with an entry outside the north-star tolerance (1e-4 relative or 1e-5 absolute)
    flips_near_zero        share of the disagreeing bits whose oracle posterior is within tolerance of zero
and, on a subset, the same figures after every pass (per-pass divergence of the two trajectories).
"""
import argparse
import json
import os
import sys
import time
import zlib

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "ldpc-simulator_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)

REGIMES = {
    # name: (code fixture, rate, Eb/N0 dB, fix_odd_check_sign, note)
    "bench": ("wimax_2304_0.5", 0.5, 2.0, False, "the bench.py workload: reference sign convention, nothing converges"),
    "r083_3.5dB": ("wimax_2304_0.83", 5.0 / 6.0, 3.5, False, "reference signs; check degrees 20/21"),
    "r083_4.5dB": ("wimax_2304_0.83", 5.0 / 6.0, 4.5, False, "reference signs"),
    "fix_1.5dB": ("wimax_2304_0.5", 0.5, 1.5, True, "odd-check sign compensated, waterfall"),
    "fix_2.0dB": ("wimax_2304_0.5", 0.5, 2.0, True, "odd-check sign compensated"),
    "fix_2.5dB": ("wimax_2304_0.5", 0.5, 2.5, True, "odd-check sign compensated"),
    "fix_6dB": ("wimax_2304_0.5", 0.5, 6.0, True, "saturated, converging"),
    "sat_6dB": ("wimax_2304_0.5", 0.5, 6.0, False, "saturated, reference signs"),
}


def llr_batch(seed, frames, n, ebn0_db, rate):
    rng = np.random.default_rng(seed)
    sig = 1.0 / np.sqrt(2.0 * rate * 10 ** (ebn0_db / 10.0))
    out = np.empty((frames, n), dtype=np.float32)
    step = 4096
    for a in range(0, frames, step):
        b = min(frames, a + step)
        out[a:b] = (2.0 * (-1.0 + sig * rng.standard_normal((b - a, n))) / sig ** 2).astype(np.float32)
    return out


def compare(res, ref, n):
    zdiff = res.z != ref["z"]
    frame_bad = zdiff.any(axis=1) | (res.ok != ref["ok"]) | (res.conv_it != ref["conv_it"])
    conv = ref["ok"] == 1
    out = {
        "frames": int(res.z.shape[0]),
        "oracle_converged_fraction": float(conv.mean()),
        "gpu_converged_fraction": float((res.ok == 1).mean()),
        "frame_agree": float(1.0 - frame_bad.mean()),
        "frame_agree_converged": float(1.0 - frame_bad[conv].mean()) if conv.any() else None,
        "frames_disagreeing": int(frame_bad.sum()),
        "frames_disagreeing_converged": int(frame_bad[conv].sum()),
        "bit_agree": float(1.0 - zdiff.mean()),
        "ok_agree": float((res.ok == ref["ok"]).mean()),
        "conv_agree": float((res.conv_it == ref["conv_it"]).mean()),
    }
    if res.post is not None and ref.get("post") is not None:
        # posteriors are comparable where both decoders stopped after the same pass
        same_exit = (res.conv_it == ref["conv_it"])
        got = res.post[same_exit].astype(np.float64)
        want = ref["post"][same_exit]
        err = np.abs(got - want)
        rel = err / np.maximum(np.abs(want), 1.0)
        viol = (err > 1e-5) & (err > 1e-4 * np.abs(want))
        out.update({
            "post_rel_median": float(np.median(rel)) if rel.size else None,
            "post_rel_p99": float(np.quantile(rel, 0.99)) if rel.size else None,
            "post_rel_max": float(rel.max()) if rel.size else None,
            "post_frames_outside_tolerance": float(viol.any(axis=1).mean()) if rel.size else None,
            "post_entries_outside_tolerance": float(viol.mean()) if rel.size else None,
        })
        flips = zdiff[same_exit]
        if flips.any():
            w = np.abs(want[flips])
            out["flips_near_zero"] = float((w <= 1e-5).mean())
            out["flip_abs_posterior_median"] = float(np.median(w))
            out["flip_abs_posterior_max"] = float(w.max())
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r2_parity_fast.json"))
    ap.add_argument("--bench-frames", type=int, default=65536)
    ap.add_argument("--frames", type=int, default=16384)
    ap.add_argument("--trace-frames", type=int, default=512)
    ap.add_argument("--regimes", default=",".join(REGIMES))
    ap.add_argument("--precision", default="f32_fast")
    ap.add_argument("--one-frame", action="store_true", help="force the one-frame-per-thread resident kernel")
    ap.add_argument("--threads", type=int, default=0)
    a = ap.parse_args()

    from conftest import load_code
    from oracle import spa_oracle as so
    from settings import Settings
    from spa_decoder import SPA_Decoder

    class Edd:
        def __init__(self, h):
            self._h_sparse_cached, (self._m, self._n) = h, h.shape

    def decoder(code, fix, iters=20):
        st = Settings()
        st.set_max_iterations(iters)
        st.set_precision(a.precision)
        st.set_fix_odd_check_sign(fix)
        if a.one_frame and hasattr(st, "set_one_frame_kernel"):
            st.set_one_frame_kernel(True)
        return SPA_Decoder(Edd(code.csr()), st)

    report = {"precision": a.precision, "one_frame_kernel": bool(a.one_frame), "max_iter": 20, "regimes": {}}
    for name in a.regimes.split(","):
        fixture, rate, ebn0, fix, note = REGIMES[name]
        code = load_code(fixture)
        frames = a.bench_frames if name == "bench" else a.frames
        t0 = time.time()
        llr = llr_batch(zlib.crc32(name.encode()) % 65521 + 20261018, frames, code.n, ebn0, rate)
        dec = decoder(code, fix)
        res = dec.decode_batch(llr, want_posterior=True, early_termination=True)
        t1 = time.time()
        ref = so.decode_batch(code.row_ptr, code.col_idx, code.n, llr.astype(np.float64), 20,
                              fix_odd_check_sign=fix, nthreads=a.threads)
        t2 = time.time()
        r = compare(res, ref, code.n)
        r.update({"code": fixture, "ebn0_db": ebn0, "fix_odd_check_sign": fix, "note": note,
                  "gpu_s": round(t1 - t0, 2), "oracle_s": round(t2 - t1, 2)})
        if name == "bench":
            # the timed bench configuration runs 20 fixed passes without early termination
            fixed = dec.decode_batch(llr, want_posterior=False, early_termination=False)
            never = ref["ok"] == 0
            r["fixed_iterations_equal_early_termination_on_unconverged"] = float(
                ((fixed.z == res.z).all(axis=1) & (fixed.ok == res.ok))[never].mean()) if never.any() else None
        # ---- per-pass divergence on a subset ----
        tf = min(a.trace_frames, frames)
        if tf > 0:
            sub = llr[:tf]
            tr = so.decode_batch(code.row_ptr, code.col_idx, code.n, sub.astype(np.float64), 20,
                                 fix_odd_check_sign=fix, want_trace=True, nthreads=a.threads)
            per_pass = []
            for p in range(1, 21):
                g = decoder(code, fix, p).decode_batch(sub, want_posterior=True, early_termination=True)
                # frames for which pass p-1 is executed by the oracle
                alive = ~np.isnan(tr["post_trace"][:, p - 1, 0])
                # and is the exit pass of the GPU run (not converged earlier there)
                alive &= (g.conv_it < 0) | (g.conv_it == p - 1)
                if not alive.any():
                    per_pass.append({"pass": p - 1, "frames": 0})
                    continue
                want = tr["post_trace"][alive, p - 1, :]
                got = g.post[alive].astype(np.float64)
                err = np.abs(got - want)
                rel = err / np.maximum(np.abs(want), 1.0)
                dz = (got < 0) != (want < 0)
                per_pass.append({"pass": p - 1, "frames": int(alive.sum()),
                                 "bit_agree": float(1.0 - dz.mean()),
                                 "frames_with_flip": float(dz.any(axis=1).mean()),
                                 "post_rel_median": float(np.median(rel)), "post_rel_p99": float(np.quantile(rel, 0.99)),
                                 "post_rel_max": float(rel.max())})
            r["per_pass"] = per_pass
        report["regimes"][name] = r
        print(name, json.dumps({k: v for k, v in r.items() if k != "per_pass"}), flush=True)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(report, f, indent=1)
    print("wrote", a.out)


if __name__ == "__main__":
    main()
