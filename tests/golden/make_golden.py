#!/usr/bin/env python3
"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Run in the authoring container only (needs /root/reference, which does not
exist on the GPU box):

    python tests/golden/make_golden.py [--only NAME ...] [--jobs 8]

The reference (omkuprin7/ldpc-simulator, python_ldpc_app/) ships no golden
vectors for the SPA path, so parity is pinned by executing it here on seeded
inputs and committing its outputs:

* ``SPA_Decoder.decode`` is driven unmodified.  To decode on an arbitrary
  graph (raw ALIST H as well as H_std) it is handed a 3-attribute stand-in for
  ``EncoderDecoderData`` (it only touches ``_h_sparse_cached``, ``_m``, ``_n``;
  spa_decoder.py:28-31,66-67).
* The posterior LLRs are never exposed by the reference
  (``_arr_aposteriori_llrs`` stays empty, spa_decoder.py:23), so the
  ``Settings`` object passed in is a subclass whose
  ``is_normalized_llr_calculate()`` looks at the caller's local
  ``arr_aposteriori_llrs`` -- it is consulted right after the posterior of
  each iteration is formed (spa_decoder.py:210).
* Code definitions are stored as CSR index arrays (tests/golden/codes/*.npz)
  as returned by the reference's own ALIST reader (utils.py:21-113).

Nothing here is imported by the product or by the tests; the tests only read
the .npz / .json files this script writes.
"""
from __future__ import annotations

import argparse
import hashlib
import io
import json
import os
import sys
import time
from concurrent.futures import ProcessPoolExecutor

sys.dont_write_bytecode = True
REF_ROOT = "/root/reference"
REF_APP = os.path.join(REF_ROOT, "python_ldpc_app")
DB = os.path.join(REF_ROOT, "Channel_Codes_Database")
sys.path.insert(0, REF_APP)

import numpy as np  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CODES_DIR = os.path.join(HERE, "codes")

CODE_FILES = {
    "bch_7_4": "BCH_7_4_1_strip.alist.txt",
    "ccsds_128_64": "Standardized LDPC Codes/CCSDS_ldpc_n128_k64.alist.txt",
    "tanner_155_64": "Custom LDPC Codes/Tanner_155_64.alist.txt",
    "wifi_648_r083": "Standardized LDPC Codes/wifi_648_r083.alist.txt",
    "wimax_576_0.5": "Wimax LDPC Codes/wimax_576_0.5.alist.txt",
    "wimax_1152_0.66B": "Wimax LDPC Codes/wimax_1152_0.66B.alist.txt",
    "wimax_2304_0.5": "Wimax LDPC Codes/wimax_2304_0.5.alist.txt",
    "wimax_2304_0.75B": "Wimax LDPC Codes/wimax_2304_0.75B.alist.txt",
    "wimax_2304_0.83": "Wimax LDPC Codes/wimax_2304_0.83.alist.txt",
}


# --------------------------------------------------------------------------- #
# reference harness
# --------------------------------------------------------------------------- #
class _GraphOnly:
    """What SPA_Decoder reads from its first argument (spa_decoder.py:28-31,66-67)."""

    def __init__(self, h_csr):
        self._h_sparse_cached = h_csr
        self._m, self._n = h_csr.shape


class _Frame:
    """What SPA_Decoder.decode reads/writes on its data buffer (:88,233,245)."""

    def __init__(self, llr):
        self._channel_data = [float(x) for x in llr]
        self._decoded_data = []


def _make_spy_settings(max_iter, calc_norm):
    from settings import Settings

    class Spy(Settings):
        def __init__(self):
            super().__init__()
            self.posteriors = []

        def is_normalized_llr_calculate(self):
            fr = sys._getframe(1)
            loc = fr.f_locals
            if "arr_aposteriori_llrs" in loc and "i_cur_iter" in loc:
                it = loc["i_cur_iter"]
                if len(self.posteriors) == it:        # first consultation in this pass
                    self.posteriors.append(np.array(loc["arr_aposteriori_llrs"], dtype=np.float64))
            return calc_norm

    s = Spy()
    s.set_max_iterations(max_iter)
    return s


def ref_decode(h_csr, llr, max_iter, calc_norm=False, edd=None):
    """Run the unmodified reference decoder on one frame."""
    from spa_decoder import SPA_Decoder
    from enums import Result

    settings = _make_spy_settings(max_iter, calc_norm)
    dec = SPA_Decoder(edd if edd is not None else _GraphOnly(h_csr), settings)
    buf = _Frame(llr)
    res = dec.decode(buf)
    return dict(
        z=np.array(buf._decoded_data, dtype=np.uint8),
        ok=(res == Result.OK),
        conv_it=int(dec.convergence_iteration),
        posts=settings.posteriors,
        norm=float(dec._d_summarize_normalized_llr) if calc_norm else 0.0,
    )


_WORKER_H = {}


def _worker_decode(task):
    key, path_or_none, llr, max_iter, calc_norm = task
    if key not in _WORKER_H:
        from scipy import sparse
        d = np.load(path_or_none)
        _WORKER_H[key] = sparse.csr_matrix(
            (np.ones(d["col_idx"].size, dtype=np.int32), d["col_idx"], d["row_ptr"]),
            shape=(int(d["m"]), int(d["n"])))
    r = ref_decode(_WORKER_H[key], llr, max_iter, calc_norm)
    return r["z"], r["ok"], r["conv_it"], r["posts"][-1], r["norm"], len(r["posts"])


def decode_many(graph_npz, llrs, max_iter, calc_norm, jobs):
    tasks = [(graph_npz, graph_npz, llrs[f], max_iter, calc_norm) for f in range(llrs.shape[0])]
    with ProcessPoolExecutor(max_workers=jobs) as ex:
        out = list(ex.map(_worker_decode, tasks, chunksize=max(1, len(tasks) // (jobs * 4))))
    z = np.stack([o[0] for o in out])
    ok = np.array([o[1] for o in out], dtype=np.uint8)
    conv = np.array([o[2] for o in out], dtype=np.int32)
    post = np.stack([o[3] for o in out])
    norm = np.array([o[4] for o in out], dtype=np.float64)
    passes = np.array([o[5] for o in out], dtype=np.int32)
    return z, ok, conv, post, norm, passes


# --------------------------------------------------------------------------- #
# helpers
# --------------------------------------------------------------------------- #
def save_graph(path, csr):
    csr = csr.tocsr()
    csr.sort_indices()
    np.savez_compressed(path, m=np.int32(csr.shape[0]), n=np.int32(csr.shape[1]),
                        row_ptr=csr.indptr.astype(np.int32), col_idx=csr.indices.astype(np.int32))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def sigma_of(speed, snr_db):
    import math
    return 1.0 / math.sqrt(2.0 * speed * (10.0 ** (snr_db * 0.1)))      # channel.py:113


def llr_from_bits(bits, g, sig, quirk):
    """channel.py:49,68,76,80 with unit normals g supplied by us (seeded)."""
    sym = np.where(np.asarray(bits) == 0, -1.0, 1.0)
    dev = sig ** 2 if quirk else sig
    return 2.0 * (sym + dev * g) / (sig ** 2)


def log(msg):
    print(f"[make_golden {time.strftime('%H:%M:%S')}] {msg}", flush=True)


# --------------------------------------------------------------------------- #
# fixture builders
# --------------------------------------------------------------------------- #
def build_codes():
    from utils import read_parity_check_matrix
    os.makedirs(CODES_DIR, exist_ok=True)
    meta = {}
    for name, rel in CODE_FILES.items():
        path = os.path.join(DB, rel)
        h = read_parity_check_matrix(path).get_sparse_matrix()
        save_graph(os.path.join(CODES_DIR, name + ".npz"), h)
        h = h.tocsr(); h.sort_indices()
        meta[name] = dict(file=rel, n=int(h.shape[1]), m=int(h.shape[0]), nnz=int(h.nnz),
                          sha256_col_idx=sha(h.indices.astype(np.int32)),
                          sha256_row_ptr=sha(h.indptr.astype(np.int32)))
        log(f"code {name}: n={h.shape[1]} m={h.shape[0]} nnz={h.nnz}")
    with open(os.path.join(CODES_DIR, "index.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)


def build_stdform():
    """EncoderDecoderData (encoder_decoder_data.py:187-219) fingerprints + H_std graphs."""
    from encoder_decoder_data import EncoderDecoderData
    out = {}
    for name in ["bch_7_4", "ccsds_128_64", "tanner_155_64", "wimax_576_0.5", "wimax_2304_0.5"]:
        t0 = time.time()
        with _quiet():
            edd = EncoderDecoderData(os.path.join(DB, CODE_FILES[name]))
        hs = edd._h_std.get_sparse_matrix().tocsr(); hs.sort_indices()
        g = edd._g.get_sparse_matrix().tocsr(); g.sort_indices()
        save_graph(os.path.join(CODES_DIR, name + ".std.npz"), hs)
        np.savez_compressed(os.path.join(CODES_DIR, name + ".stdmeta.npz"),
                            permutation=np.array(edd._permutation, dtype=np.int32),
                            m=np.int32(edd._m), n=np.int32(edd._n), k=np.int32(edd._k),
                            rate=np.float64(edd._rate),
                            g_row_ptr=g.indptr.astype(np.int32), g_col_idx=g.indices.astype(np.int32))
        out[name] = dict(m=int(edd._m), n=int(edd._n), k=int(edd._k), nnz_std=int(hs.nnz),
                         perm_head=[int(x) for x in edd._permutation[:6]], seconds=round(time.time() - t0, 2))
        log(f"stdform {name}: {out[name]}")
    with open(os.path.join(CODES_DIR, "stdform_index.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


class _quiet:
    def __enter__(self):
        self._o = sys.stdout
        sys.stdout = io.StringIO()

    def __exit__(self, *a):
        sys.stdout = self._o


def build_bch_kat():
    """Deterministic known-answer vectors on BCH(7,4), decoded on H_std, 50 passes max."""
    d = np.load(os.path.join(CODES_DIR, "bch_7_4.std.npz"))
    from scipy import sparse
    h = sparse.csr_matrix((np.ones(d["col_idx"].size, dtype=np.int32), d["col_idx"], d["row_ptr"]),
                          shape=(int(d["m"]), int(d["n"])))
    llrs = np.array([
        [-1.5, 0.3, -2.0, 0.8, -0.2, -1.1, 0.6],
        [-4.0, -3.0, 2.5, -5.0, -1.0, 0.5, -2.0],
        [0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0],
        [40.0, -40.0, 40.0, 40.0, -40.0, 40.0, -40.0],
        [1e-12, -1e-12, 3.0, -3.0, 1e-11, 0.5, -0.5],        # exercises the |tanh|<=1e-10 branch
        [-0.0, 0.0, -0.0, 2.0, -2.0, 0.0, 1.0],              # signed zeros / ties
        [35.5, 34.9, -35.1, 36.0, -34.0, 35.0, -35.0],       # straddles the 17.5 clip
    ])
    max_iter = 50
    trace = np.full((llrs.shape[0], max_iter, 7), np.nan)
    z = np.zeros((llrs.shape[0], 7), np.uint8); ok = np.zeros(llrs.shape[0], np.uint8)
    conv = np.zeros(llrs.shape[0], np.int32); passes = np.zeros(llrs.shape[0], np.int32)
    for f in range(llrs.shape[0]):
        r = ref_decode(h, llrs[f], max_iter)
        z[f] = r["z"]; ok[f] = r["ok"]; conv[f] = r["conv_it"]; passes[f] = len(r["posts"])
        for it, p in enumerate(r["posts"]):
            trace[f, it] = p
    np.savez_compressed(os.path.join(HERE, "bch74_kat.npz"), llr=llrs, z=z, ok=ok, conv_it=conv,
                        passes=passes, post_trace=trace, max_iter=np.int32(max_iter))
    log(f"bch kat: ok={ok.tolist()} conv={conv.tolist()} passes={passes.tolist()}")


def _mixed_llrs(rng, n, frames, snrs, speed, quirk, codewords=None):
    llr = np.zeros((frames, n)); snr_of = np.zeros(frames)
    for f in range(frames):
        snr = snrs[f % len(snrs)]
        sig = sigma_of(speed, snr)
        bits = codewords[f] if codewords is not None else np.zeros(n, dtype=np.uint8)
        llr[f] = llr_from_bits(bits, rng.standard_normal(n), sig, quirk)
        snr_of[f] = snr
    return llr, snr_of


def build_decode_set(tag, graph_name, frames, snrs, speed, quirk, max_iter, seed, jobs,
                     calc_norm=False, random_codewords=False, store_f32=False):
    """Seeded frames through the unmodified reference decoder on the named graph.

    ``store_f32`` (the large sets): the LLRs are rounded to float32 BEFORE the reference decodes them, so the
    file can hold them -- and the posteriors -- in single precision without changing what was decoded."""
    gpath = os.path.join(CODES_DIR, graph_name + ".npz")
    d = np.load(gpath)
    n, m = int(d["n"]), int(d["m"])
    rng = np.random.default_rng(seed)
    codewords = None; data = None
    if random_codewords:
        # needs the reference's G (H_std column order); map back to ALIST order for raw graphs
        base = graph_name.replace(".std", "")
        meta = np.load(os.path.join(CODES_DIR, base + ".stdmeta.npz"))
        from scipy import sparse
        k = int(meta["k"])
        g = sparse.csr_matrix((np.ones(meta["g_col_idx"].size, dtype=np.int32), meta["g_col_idx"],
                               meta["g_row_ptr"]), shape=(k, int(meta["n"])))
        data = rng.integers(0, 2, size=(frames, k), dtype=np.uint8)
        cw_std = (g.T.dot(data.T.astype(np.int32)) % 2).T.astype(np.uint8)   # data_buffer.py:47-82
        if graph_name.endswith(".std"):
            codewords = cw_std
        else:
            perm = meta["permutation"]
            codewords = np.zeros_like(cw_std)
            codewords[:, perm] = cw_std            # H_std[:, j] = H[:, perm[j]]
    llr, snr_of = _mixed_llrs(rng, n, frames, snrs, speed, quirk, codewords)
    if store_f32:
        llr = llr.astype(np.float32).astype(np.float64)
    t0 = time.time()
    z, ok, conv, post, norm, passes = decode_many(gpath, llr, max_iter, calc_norm, jobs)
    if store_f32:
        llr, post = llr.astype(np.float32), post.astype(np.float32)
    extra = {}
    if data is not None:
        extra = dict(data=data, codeword=codewords)
    np.savez_compressed(os.path.join(HERE, tag + ".npz"), graph=graph_name, llr=llr, snr_db=snr_of,
                        speed=np.float64(speed), sigma_sq_quirk=np.int32(quirk),
                        max_iter=np.int32(max_iter), seed=np.int64(seed), z=np.packbits(z, axis=1),
                        n=np.int32(n), m=np.int32(m), ok=ok, conv_it=conv, post=post, norm=norm,
                        passes=passes, calc_norm=np.int32(calc_norm), **extra)
    log(f"{tag}: {frames} frames on {graph_name} in {time.time() - t0:.1f}s; ok={int(ok.sum())} "
        f"conv hist={np.bincount(conv[conv >= 0], minlength=1)[:8].tolist()}")


def _mc_worker(task):
    """main.py:295-339 single-process loop, verbatim semantics, seeded."""
    (matrix_path, speed, snr, blocks, max_iter, seed) = task
    import random
    from encoder_decoder_data import EncoderDecoderData
    from data_buffer import DataBuffer
    from channel import Channel
    from spa_decoder import SPA_Decoder
    from settings import Settings
    from enums import Result
    with _quiet():
        edd = EncoderDecoderData(matrix_path)
    st = Settings(); st.set_max_iterations(max_iter)
    random.seed(seed)
    ch = Channel.create_channel(speed, snr, 0.0, 1, 0.1, 1)
    ch._rng = np.random.RandomState(seed % (2 ** 31))     # the reference seeds from the clock (channel.py:30)
    dec = SPA_Decoder(edd, st)
    okc = fail = biterr = convsum = convcnt = 0
    for _ in range(blocks):
        buf = DataBuffer(edd._k)
        buf.encode(edd._g_transpose)
        ch.process(buf)
        res = dec.decode(buf)
        if res == Result.OK:
            okc += 1
        else:
            fail += 1
            info = buf._data[:edd._k]; decd = buf._decoded_data[:edd._k]
            biterr += sum(1 for a, b in zip(info, decd) if a != (b ^ 1))
        if dec.convergence_iteration >= 0:
            convsum += dec.convergence_iteration; convcnt += 1
    return snr, blocks, fail, biterr, convsum, convcnt


def build_bch_mc_anchor(jobs, blocks_per_point=40000):
    """Reference Monte-Carlo anchors on BCH(7,4) (config 0), its own conventions."""
    path = os.path.join(DB, CODE_FILES["bch_7_4"])
    out = {}
    for label, speed in (("speed_4_7", 4.0 / 7.0), ("speed_1", 1.0)):
        tasks = []
        per = blocks_per_point // jobs
        for snr in range(0, 7):
            for w in range(jobs):
                tasks.append((path, speed, float(snr), per, 50, 1000 * snr + w + (7 if speed == 1.0 else 0)))
        with ProcessPoolExecutor(max_workers=jobs) as ex:
            res = list(ex.map(_mc_worker, tasks))
        pts = {}
        for snr, blocks, fail, biterr, convsum, convcnt in res:
            p = pts.setdefault(str(snr), dict(frames=0, frame_err=0, bit_err=0, conv_sum=0, conv_cnt=0))
            p["frames"] += blocks; p["frame_err"] += fail; p["bit_err"] += biterr
            p["conv_sum"] += convsum; p["conv_cnt"] += convcnt
        out[label] = dict(speed=speed, max_iter=50, k=4, n=7, points=pts)
        log(f"bch mc {label}: " + ", ".join(f"{s}dB FER={p['frame_err']/p['frames']:.3e}" for s, p in sorted(pts.items())))
    with open(os.path.join(HERE, "bch74_mc_anchor.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


def build_w576_mc_anchor(jobs):
    """Reference Monte-Carlo anchors for the graph main.py really decodes on (H_std of WiMAX-576 r1/2),
    its own conventions (sigma^2 quirk, speed 0.5, 20 passes).  Slow: ~17 s per frame at 0 dB."""
    path = os.path.join(DB, CODE_FILES["wimax_576_0.5"])
    plan = {0.0: 40, 3.0: 40, 4.0: 60, 5.0: 120}          # frames per worker
    pts = {}
    for snr, per in plan.items():
        tasks = [(path, 0.5, snr, per, 20, 57600 + int(snr) * 100 + w) for w in range(jobs)]
        with ProcessPoolExecutor(max_workers=jobs) as ex:
            res = list(ex.map(_mc_worker, tasks))
        p = dict(frames=0, frame_err=0, bit_err=0, conv_sum=0, conv_cnt=0, per_worker_bit_err=[], per_worker_frames=per)
        for _snr, blocks, fail, biterr, convsum, convcnt in res:
            p["frames"] += blocks; p["frame_err"] += fail; p["bit_err"] += biterr
            p["conv_sum"] += convsum; p["conv_cnt"] += convcnt; p["per_worker_bit_err"].append(biterr)
        pts[str(snr)] = p
        log(f"w576 std mc {snr} dB: frames {p['frames']} FER {p['frame_err']/p['frames']:.4f} "
            f"BER {p['bit_err']/(288*p['frames']):.5f}")
        with open(os.path.join(HERE, "wimax576_std_mc_anchor.json"), "w") as f:
            json.dump(dict(speed=0.5, max_iter=20, k=288, n=576, points=pts), f, indent=1, sort_keys=True)


def build_norm_history():
    """Per-pass history of the "normalized LLR" metric (spa_decoder.py:210-228): the lists SPA_Decoder leaves in
    _arr_changed_by_iterations / _normalized_llr_by_iterations after one decode(), for a few seeded frames."""
    from scipy import sparse
    from spa_decoder import SPA_Decoder
    out = {}
    for graph_name, frames, snrs, max_iter, seed in [("ccsds_128_64", 12, [1.0, 2.5, 4.0], 20, 901), ("bch_7_4.std", 8, [0.0, 3.0], 50, 902)]:
        d = np.load(os.path.join(CODES_DIR, graph_name + ".npz"))
        n, m = int(d["n"]), int(d["m"])
        h = sparse.csr_matrix((np.ones(d["col_idx"].size, dtype=np.int32), d["col_idx"], d["row_ptr"]), shape=(m, n))
        rng = np.random.default_rng(seed)
        llr, snr_of = _mixed_llrs(rng, n, frames, snrs, 0.5, False)
        rows = []
        for f in range(frames):
            settings = _make_spy_settings(max_iter, True)
            dec = SPA_Decoder(_GraphOnly(h), settings)
            buf = _Frame(llr[f])
            dec.decode(buf)
            rows.append(dict(llr=[float(x) for x in llr[f]], conv_it=int(dec.convergence_iteration),
                             changed=[int(x) for x in dec._arr_changed_by_iterations],
                             normalized=[float(x) for x in dec._normalized_llr_by_iterations],
                             summary=float(dec._d_summarize_normalized_llr)))
        out[graph_name] = dict(max_iter=max_iter, frames=rows)
        log(f"norm history {graph_name}: passes {[len(r['changed']) for r in rows]}")
    with open(os.path.join(HERE, "norm_history.json"), "w") as fh:
        json.dump(out, fh)


def build_results_sample():
    """results.py writers on fixed inputs -> byte-exact expected JSON / CSV text."""
    from results import SimulationResult, SimulationConfig, SNRPointResult
    cfg = SimulationConfig(matrix_path="codes/wimax_576_0.5.alist.txt", n=576, m=288, k=288, rate=0.5,
                           blocks=1000, max_iterations=20, encoding_method="standard",
                           interleaver_type="none", decoder_type="sumproduct", channel_mode=1,
                           modulation=1, speed=0.5, snr_range=(0.0, 2.0, 1.0), threads=1,
                           timestamp="2026-10-18T12:00:00", interference_snr=1.0, p=0.1)
    pts = [SNRPointResult(snr_db=float(s), ber=b, fer=f, avg_normalized_llr=0.0, total_blocks=1000,
                          successful_blocks=1000 - int(f * 1000), failed_blocks=int(f * 1000),
                          avg_convergence_iterations=c, matrix_path="codes/wimax_576_0.5.alist.txt",
                          modulation=1, max_iterations=20, interleaver="none", encoding_method="standard")
           for s, b, f, c in ((0, 0.0917, 1.0, 0.0), (1, 0.031415, 0.5, 7.25), (2, 1e-7, 0.001, 3.0))]
    res = SimulationResult(config=cfg, snr_points=pts, wall_clock_seconds=12.5,
                           adaptation_log=[{"snr_db": 1.0, "action": "Увеличение итераций"}])
    jp = os.path.join(HERE, "results_sample.json"); cp = os.path.join(HERE, "results_sample.csv")
    res.to_json(jp); res.to_csv(cp)
    back = SimulationResult.from_json(jp)
    assert back.config.snr_range == (0.0, 2.0, 1.0)
    log("results sample written")


def build_catalog_listing():
    """matrix_catalog.py:24-125 over the real database -> expected (name, n, k, m, rate, family)."""
    from matrix_catalog import MatrixCatalog
    cat = MatrixCatalog(DB)
    rows = [dict(rel=os.path.relpath(mi.path, DB), name=mi.name, n=mi.n, k=mi.k, m=mi.m, rate=mi.rate,
                 family=mi.family) for mi in cat.matrices]
    heads = {}
    for r in rows:                       # header line needed by the fallback parser (:127-142)
        with open(os.path.join(DB, r["rel"])) as f:
            heads[r["rel"]] = f.readline().strip()
    with open(os.path.join(HERE, "catalog_listing.json"), "w") as f:
        json.dump(dict(repr=repr(cat), entries=rows, first_lines=heads), f, indent=1)
    log(repr(cat))


def build_adaptive_strategy():
    """adaptive.py:61-124 (ThresholdStrategy.evaluate) and :384-412 (_apply_action against the real catalog)
    on a grid of point results -> expected actions / state transitions."""
    from adaptive import AdaptiveController, AdaptiveState, ThresholdStrategy
    from matrix_catalog import MatrixCatalog
    from results import SNRPointResult
    cases = []
    strategies = [dict(), dict(high_ber_threshold=5e-2, low_ber_threshold=1e-4, fer_threshold=0.3, convergence_ratio=0.5)]
    for si, kw in enumerate(strategies):
        st = ThresholdStrategy(**kw)
        for ber in (0.0, 1e-7, 9.99e-6, 1e-5, 5e-5, 1e-3, 1e-2, 1.0001e-2, 0.05, 0.2):
            for fer in (0.0, 0.3, 0.5, 0.5001, 1.0):
                for conv, max_it in ((0.0, 5), (4.0, 5), (4.01, 5), (3.0, 50), (45.0, 50), (90.0, 100), (60.0, 64)):
                    for il in ("none", "random"):
                        state = AdaptiveState("x/wimax_576_0.5.alist.txt", 0.5, 1, max_it, il, "standard")
                        pt = SNRPointResult(snr_db=1.0, ber=ber, fer=fer, avg_normalized_llr=0.0, total_blocks=100,
                                            successful_blocks=50, failed_blocks=50, avg_convergence_iterations=conv)
                        act = st.evaluate(state, pt)
                        cases.append(dict(strategy=si, ber=ber, fer=fer, conv=conv, max_it=max_it, interleaver=il,
                                          action=None if act is None else dict(
                                              matrix=act.new_matrix_path, modulation=act.new_modulation,
                                              max_iterations=act.new_max_iterations, interleaver=act.new_interleaver,
                                              reason=act.reason)))
    # rate ladder walks over the real catalog
    cat = MatrixCatalog(DB)
    ctl = AdaptiveController(ThresholdStrategy(), cat)
    ctl._get_encoder_decoder_data = lambda path: None          # no matrix loading, only the state transition
    walks = []
    for start in ("Wimax LDPC Codes/wimax_576_0.5.alist.txt", "Wimax LDPC Codes/wimax_2304_0.83.alist.txt",
                  "Standardized LDPC Codes/wigig_R063_N672_K420.alist.txt", "BCH_7_4_1_strip.alist.txt"):
        for direction in ("__LOWER_RATE__", "__HIGHER_RATE__"):
            path = os.path.join(DB, start)
            info = ctl._find_current_matrix_info(path)
            state = AdaptiveState(path, info.rate if info else 0.0, 1, 5, "none", "standard")
            trail = []
            for _ in range(6):
                from adaptive import AdaptiveAction
                with _quiet():
                    ctl._apply_action(AdaptiveAction(new_matrix_path=direction), state, None, None, None)
                trail.append([os.path.relpath(state.current_matrix_path, DB), state.current_rate])
            walks.append(dict(start=start, direction=direction, trail=trail))
    with open(os.path.join(HERE, "adaptive_strategy.json"), "w") as f:
        json.dump(dict(strategies=strategies, cases=cases, walks=walks), f, separators=(",", ":"))
    log(f"adaptive: {len(cases)} strategy cases, {len(walks)} catalog walks")


def build_channel_modes():
    """channel.py:38-100 for modes 2 and 3 (and modulation 2): the noise comes from the Park-Miller generators
    seeded with IDUM1 / IDUM2 (deterministic), the hit decision of mode 2 from numpy's global RandomState."""
    from channel import Channel

    class Buf:
        def __init__(self, bits):
            self._encoded_data = list(bits)
            self._channel_data = []

    rng = np.random.default_rng(99)
    cases = []
    for mode, mod, p, speed, sn1, sn2 in ((2, 1, 0.3, 0.5, 3.0, 1.0), (2, 2, 0.7, 0.5, 1.0, -2.0), (3, 1, 0.25, 0.75, 2.0, 0.0),
                                          (3, 2, 0.9, 0.5, 4.0, 4.0), (2, 1, 1.0, 1.0, 0.0, 0.0)):
        bits = [int(b) for b in rng.integers(0, 2, size=96)]
        ch = Channel.create_channel(speed, sn1, sn2, mode, p, mod)
        np.random.seed(4242)
        out = []
        for _frame in range(2):                 # the generators keep their state from frame to frame
            buf = Buf(bits)
            ch.process(buf)
            out.append(buf._channel_data)
        cases.append(dict(mode=mode, modulation=mod, p=p, speed=speed, sn1=sn1, sn2=sn2, bits=bits, numpy_seed=4242,
                          L_c=[ch.L_c1, ch.L_c2, ch.L_c3], sigma=[ch.gen_ptr.sigma, ch.gen_ptr2.sigma], llr=out))
    with open(os.path.join(HERE, "channel_modes.json"), "w") as f:
        json.dump(dict(cases=cases), f)
    log(f"channel modes: {len(cases)} cases")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", nargs="*", default=None)
    ap.add_argument("--jobs", type=int, default=os.cpu_count() or 1)
    a = ap.parse_args()
    J = a.jobs
    steps = {
        "codes": build_codes,
        "stdform": build_stdform,
        "bch_kat": build_bch_kat,
        "results": build_results_sample,
        "norm_history": build_norm_history,
        "adaptive": build_adaptive_strategy,
        "channel_modes": build_channel_modes,
        "catalog": build_catalog_listing,
        # BCH(7,4) on H_std, reference channel conventions (sigma^2 quirk, speed 4/7), random codewords
        "bch_random": lambda: build_decode_set("bch74_std_random", "bch_7_4.std", 4096, [0, 1, 2, 3, 4, 5, 6],
                                               4.0 / 7.0, True, 50, 7401, J, calc_norm=True, random_codewords=True),
        # all-even-degree code: converges, exercises conv_it > 0
        "ccsds_alist": lambda: build_decode_set("ccsds128_alist", "ccsds_128_64", 512, [1, 2, 3, 4, 5],
                                                0.5, False, 20, 12801, J, calc_norm=True),
        "ccsds_std": lambda: build_decode_set("ccsds128_std", "ccsds_128_64.std", 96, [2, 3, 4, 5, 6],
                                              0.5, True, 20, 12802, J, random_codewords=True),
        "tanner_std": lambda: build_decode_set("tanner155_std", "tanner_155_64.std", 48, [2, 4, 6],
                                               64.0 / 155.0, True, 10, 15501, J, random_codewords=True),
        "wifi_alist": lambda: build_decode_set("wifi648_alist", "wifi_648_r083", 32, [3, 5, 7],
                                               0.83, False, 20, 64801, J),
        # config 1: WiMAX-576 r1/2, mix of Eb/N0, 20 passes, both graphs
        "w576_alist": lambda: build_decode_set("wimax576_alist", "wimax_576_0.5", 128, [1, 2, 3, 4, 6],
                                               0.5, False, 20, 20261018, J),
        "w576_alist_cw": lambda: build_decode_set("wimax576_alist_cw", "wimax_576_0.5", 32, [2, 4, 6],
                                                  0.5, False, 20, 57602, J, random_codewords=True),
        "w576_std": lambda: build_decode_set("wimax576_std", "wimax_576_0.5.std", 32, [3, 4, 5, 6],
                                             0.5, True, 20, 57603, J, random_codewords=True),
        # config 2 / 3 codes
        "w2304_alist": lambda: build_decode_set("wimax2304_alist", "wimax_2304_0.5", 24, [1, 2, 3],
                                                0.5, False, 20, 230401, J),
        "w2304_075B": lambda: build_decode_set("wimax2304_075B_alist", "wimax_2304_0.75B", 16, [2, 3, 4],
                                               0.75, False, 20, 230402, J),
        # all check degrees even (20): converges under the reference's sign convention -> early termination
        "w2304_083": lambda: build_decode_set("wimax2304_083_alist", "wimax_2304_0.83", 48, [3.5, 4.0, 4.5, 5.0],
                                              0.83, False, 20, 230404, J),
        "w2304_std": lambda: build_decode_set("wimax2304_std", "wimax_2304_0.5.std", 8, [5, 6],
                                              0.5, True, 2, 230403, J),
        # round 2: a thicker live-reference pin
        "w576_alist_2k": lambda: build_decode_set("wimax576_alist_2k", "wimax_576_0.5", 2048, [1, 2, 3, 4, 6],
                                                  0.5, False, 20, 20261019, J, store_f32=True),
        "w2304_alist_256": lambda: build_decode_set("wimax2304_alist_256", "wimax_2304_0.5", 256, [1, 2, 3, 6],
                                                    0.5, False, 20, 230411, J, store_f32=True),
        "w2304_std_20": lambda: build_decode_set("wimax2304_std_20it", "wimax_2304_0.5.std", 8, [4, 5, 6, 7],
                                                 0.5, True, 20, 230413, J, random_codewords=True),
        "bch_mc": lambda: build_bch_mc_anchor(J),
        "w576_mc": lambda: build_w576_mc_anchor(J),
    }
    for name, fn in steps.items():
        if a.only and name not in a.only:
            continue
        t0 = time.time()
        fn()
        log(f"step {name} done in {time.time() - t0:.1f}s")


if __name__ == "__main__":
    main()
