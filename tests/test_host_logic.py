"""Host-side logic: ALIST reader, edge index, QC detection, standard form, catalog, results, channel."""
import json
import os

import numpy as np
import pytest
from scipy import sparse

from conftest import GOLDEN, REPO, load_code, load_golden

import utils
from channel import Channel
from data_buffer import DataBuffer
from encoder_decoder_data import EncoderDecoderData, standard_form
from enums import InterleaverType, LDPCDecoderType, Result
from generator import Generator
from matrix import Matrix
from matrix_catalog import MatrixCatalog
from matrix_sparse import SparseMatrix
from results import SimulationConfig, SimulationResult, SNRPointResult
from settings import Settings

CODES = ["bch_7_4", "ccsds_128_64", "tanner_155_64", "wifi_648_r083", "wimax_576_0.5", "wimax_2304_0.5",
         "wimax_2304_0.75B", "wimax_2304_0.83"]


# ---- ALIST ------------------------------------------------------------------------------
@pytest.mark.parametrize("name", CODES)
def test_alist_write_read_round_trip(name, tmp_path):
    code = load_code(name)
    path = tmp_path / (name + ".alist.txt")
    utils.write_alist(str(path), code.csr())
    h = utils.read_parity_check_matrix(str(path))
    assert (h.get_rows(), h.get_cols()) == (code.m, code.n)
    rp, ci = h.csr_pattern()
    assert np.array_equal(rp, code.row_ptr) and np.array_equal(ci, code.col_idx)
    assert h.get_sparse_matrix().dtype == np.int32 and set(h.get_sparse_matrix().data.tolist()) == {1}


def test_alist_bch_text_as_in_the_database(tmp_path):
    # the BCH(7,4) file of the reference database, 14 lines (format: utils.py:26-98)
    text = ("7 3\n3 4\n1 1 2 2 3 2 1 \n4 4 4 \n1 0 0 \n2 0 0 \n1 3 0 \n1 2 0 \n1 2 3 \n2 3 0 \n3 0 0 \n"
            "1 3 4 5 \n2 4 5 6 \n3 5 6 7 \n")
    p = tmp_path / "BCH_7_4_1_strip.alist.txt"
    p.write_text(text)
    h = utils.read_parity_check_matrix(str(p))
    assert h.get_data() == [[1, 0, 1, 1, 1, 0, 0], [0, 1, 0, 1, 1, 1, 0], [0, 0, 1, 0, 1, 1, 1]]
    assert h.get_message_bit_length() == 4


@pytest.mark.parametrize("text", ["", "7\n", "0 3\n", "7 3\n3 4\n1 1\n4 4 4\n", "2 1\n1 2\n1 1\n2\n1\n1\n1 9\n"])
def test_alist_errors_give_an_empty_matrix(text, tmp_path, capsys):
    p = tmp_path / "bad.alist.txt"
    p.write_text(text)
    h = utils.read_parity_check_matrix(str(p))
    assert h.get_rows() == 0 and h.get_cols() == 0          # utils.py:109-113
    assert "Error" in capsys.readouterr().out
    with pytest.raises(ValueError):
        EncoderDecoderData(str(p))                            # encoder_decoder_data.py:194-195


def test_alist_missing_file_gives_empty_matrix(tmp_path):
    assert utils.read_parity_check_matrix(str(tmp_path / "nope.txt")).get_cols() == 0


@pytest.mark.skipif(not os.path.isdir("/root/reference/Channel_Codes_Database"), reason="reference DB not present")
def test_reader_on_the_real_database_files():
    with open(os.path.join(GOLDEN, "codes", "index.json")) as f:
        index = json.load(f)
    for name, meta in index.items():
        h = utils.read_parity_check_matrix(os.path.join("/root/reference/Channel_Codes_Database", meta["file"]))
        code = load_code(name)
        rp, ci = h.csr_pattern()
        assert np.array_equal(rp, code.row_ptr) and np.array_equal(ci, code.col_idx)


# ---- edge index / QC --------------------------------------------------------------------
@pytest.mark.parametrize("name", CODES + ["wimax_576_0.5.std"])
def test_edge_index_is_csr_and_csc_permutation(name):
    code = load_code(name)
    ei = code.sparse_matrix().edge_index()
    nnz = code.nnz
    assert np.array_equal(np.sort(ei["csc_edge"]), np.arange(nnz))
    csc = code.csr().tocsc()
    csc.sort_indices()
    assert np.array_equal(ei["col_ptr"], csc.indptr)
    rows_of = ei["edge_row"][ei["csc_edge"]]
    assert np.array_equal(rows_of, csc.indices)                # ascending row inside each column
    assert np.array_equal(ei["col_idx"][ei["csc_edge"]], np.repeat(np.arange(code.n), np.diff(csc.indptr)))


@pytest.mark.parametrize("name,expect", [("wimax_576_0.5", (24, 12, 24, 76)), ("wimax_2304_0.5", (96, 12, 24, 76)),
                                         ("wimax_2304_0.75B", (96, 6, 24, 88)), ("wimax_2304_0.83", (96, 4, 24, 80)), ("wifi_648_r083", (27, 4, 24, 88)),
                                         ("tanner_155_64", (31, 3, 5, 15)), ("bch_7_4", None), ("ccsds_128_64", None)])
def test_qc_detection(name, expect):
    code = load_code(name)
    got = code.sparse_matrix().detect_qc()
    if expect is None:
        assert got is None
        return
    z, sh = got
    assert (z, sh.shape[0], sh.shape[1], int((sh >= 0).sum())) == expect
    # rebuilding H from the shift table gives the original matrix
    rows, cols = [], []
    for br in range(sh.shape[0]):
        for bc in range(sh.shape[1]):
            if sh[br, bc] >= 0:
                r = np.arange(z)
                rows.append(br * z + r)
                cols.append(bc * z + (r + sh[br, bc]) % z)
    h = sparse.csr_matrix((np.ones(z * expect[3], dtype=np.int32), (np.concatenate(rows), np.concatenate(cols))),
                          shape=(code.m, code.n))
    assert (h != code.csr()).nnz == 0


def test_wimax_2304_shift_table_is_the_802_16e_rate_half_matrix():
    z, sh = load_code("wimax_2304_0.5").sparse_matrix().detect_qc()
    assert sh[0].tolist() == [-1, 94, 73, -1, -1, -1, -1, -1, 55, 83, -1, -1, 7, 0] + [-1] * 10
    assert sh[11].tolist() == [43, -1, -1, -1, -1, 66, -1, 41, -1, -1, -1, 26, 7] + [-1] * 10 + [0]


# ---- standard form / generator ------------------------------------------------------------
@pytest.mark.parametrize("name", ["bch_7_4", "ccsds_128_64", "tanner_155_64", "wimax_576_0.5", "wimax_2304_0.5"])
def test_standard_form_equals_reference(name, stdform_index):
    code = load_code(name)
    std = load_code(name + ".std")
    meta = np.load(os.path.join(GOLDEN, "codes", name + ".stdmeta.npz"))
    h_std, perm, rank = standard_form(code.sparse_matrix())
    assert rank == int(meta["m"]) == stdform_index[name]["m"]
    assert perm == meta["permutation"].tolist()
    assert perm[:6] == stdform_index[name]["perm_head"]
    h_std.sort_indices()
    assert h_std.nnz == stdform_index[name]["nnz_std"]
    assert np.array_equal(h_std.indptr, std.row_ptr) and np.array_equal(h_std.indices, std.col_idx)


def test_encoder_decoder_data_surface(tmp_path, capsys):
    code = load_code("tanner_155_64")
    path = tmp_path / "Tanner_155_64.alist.txt"
    utils.write_alist(str(path), code.csr())
    edd = EncoderDecoderData(str(path))
    assert "Warning: Matrix rank is 91, expected 93" in capsys.readouterr().out
    meta = np.load(os.path.join(GOLDEN, "codes", "tanner_155_64.stdmeta.npz"))
    assert (edd._n, edd._m, edd._k) == (155, 91, 64) and edd._rate == pytest.approx(64 / 155)
    g = edd._g.get_sparse_matrix().tocsr()
    g.sort_indices()
    assert np.array_equal(g.indptr, meta["g_row_ptr"]) and np.array_equal(g.indices, meta["g_col_idx"])
    assert edd._g_transpose.get_rows() == 155 and edd._h_sparse_cached is edd._h_std.get_sparse_matrix()
    coo, v2c, c2v = edd.get_decoder_structures()
    assert edd._decoder_structures_initialized and sum(len(v) for v in c2v.values()) == coo.nnz
    # H_std = [A | I]
    hs = edd._h_std.get_sparse_matrix().toarray()
    assert np.array_equal(hs[:, edd._k:], np.eye(edd._m, dtype=hs.dtype))
    # every codeword satisfies both H_std and (after un-permuting) the raw H
    u = np.random.default_rng(1).integers(0, 2, size=(16, edd._k), dtype=np.uint8)
    cw = edd.encode_batch(u)
    assert not ((hs @ cw.T) % 2).any()
    assert not ((code.csr() @ edd.to_alist_order(cw).T) % 2).any()
    assert edd.info_mask("std").sum() == edd._k == edd.info_mask("alist").sum()


def test_bch_fingerprint_from_survey():
    edd = EncoderDecoderData(h=load_code("bch_7_4").sparse_matrix())
    assert edd._permutation == [3, 4, 5, 6, 0, 1, 2]
    assert edd._h_std.get_data() == [[1, 0, 1, 1, 1, 0, 0], [1, 1, 1, 0, 0, 1, 0], [0, 1, 1, 1, 0, 0, 1]]
    assert edd._g.get_data() == [[1, 0, 0, 0, 1, 1, 0], [0, 1, 0, 0, 0, 1, 1], [0, 0, 1, 0, 1, 1, 1],
                                 [0, 0, 0, 1, 1, 0, 1]]


def test_data_buffer_encode_matches_oracle():
    from oracle import spa_oracle as so
    edd = EncoderDecoderData(h=load_code("ccsds_128_64").sparse_matrix())
    hs = edd._h_std.get_sparse_matrix().toarray().astype(np.uint8)
    buf = DataBuffer(edd._k)
    assert len(buf._data) == edd._k and buf.get_size() == edd._k
    buf.encode(edd._g_transpose)
    assert buf._encoded_data == so.encode(hs, np.array(buf._data, dtype=np.uint8)).tolist()
    twin = DataBuffer(0)
    twin._data = list(buf._data)
    twin.encode(edd._g)                                         # G instead of G^T
    assert twin._encoded_data == buf._encoded_data
    with pytest.raises(ValueError):
        DataBuffer(3).encode(edd._g)


# ---- sparse / dense matrix carriers ---------------------------------------------------------
def test_sparse_matrix_algebra():
    a = SparseMatrix(values=[[1, 0, 1], [0, 1, 1]])
    assert (a.get_rows(), a.get_cols(), a.get_message_bit_length()) == (2, 3, 1)
    assert a.multiply(a.transpose()).get_data() == [[0, 1], [1, 0]]
    a.permute_columns([2, 0, 1])
    assert a.get_data() == [[1, 1, 0], [1, 0, 1]]
    a.swap_rows(0, 1)
    assert a.get_data() == [[1, 0, 1], [1, 1, 0]]
    a.permute_rows([1, 0])
    assert a.get_data() == [[1, 1, 0], [1, 0, 1]]
    assert a.extract_sub_matrix(0, 1, 2, 2).get_data() == [[1, 0], [0, 1]]
    i2 = SparseMatrix.create_identity_matrix(2)
    assert i2.concatenate_horizontally(a).get_cols() == 5 and a.concatenate_vertically(a).get_rows() == 4
    a.set_element(0, 2, 1)
    assert a.get_element(0, 2) == 1
    a.add_row([0, 0, 1])
    assert a.get_rows() == 3
    with pytest.raises(IndexError):
        a.get_element(5, 0)
    with pytest.raises(ValueError):
        a.permute_columns([0, 1])
    empty = SparseMatrix()
    assert empty.get_rows() == 0 and empty.get_cols() == 0


def test_dense_matrix_twin_agrees_with_sparse():
    vals = [[1, 0, 1, 1], [0, 1, 1, 0], [1, 1, 0, 1]]
    d, s = Matrix(values=vals), SparseMatrix(values=vals)
    assert d.multiply(d.transpose()).get_data() == s.multiply(s.transpose()).get_data()
    d.permute_columns([3, 2, 1, 0]); s.permute_columns([3, 2, 1, 0])
    assert d.get_data() == s.get_data()
    assert d.extract_sub_matrix(1, 1, 2, 2).get_data() == s.extract_sub_matrix(1, 1, 2, 2).get_data()
    assert d.to_sparse().get_data() == s.get_data()
    assert Matrix.create_identity_matrix(2).get_data() == [[1, 0], [0, 1]]


# ---- catalog ------------------------------------------------------------------------------
def test_catalog_reproduces_reference_listing(tmp_path):
    with open(os.path.join(GOLDEN, "catalog_listing.json")) as f:
        listing = json.load(f)
    for e in listing["entries"]:
        p = tmp_path / e["rel"]
        p.parent.mkdir(parents=True, exist_ok=True)
        p.write_text(listing["first_lines"][e["rel"]] + "\n")
    cat = MatrixCatalog(str(tmp_path))
    assert repr(cat) == listing["repr"] and len(cat) == len(listing["entries"]) == 119
    key = lambda d: (d["family"], d["rate"], d["n"], d["name"])
    got = sorted((dict(name=m.name, n=m.n, k=m.k, m=m.m, rate=m.rate, family=m.family) for m in cat.matrices), key=key)
    ref = sorted((dict(name=e["name"], n=e["n"], k=e["k"], m=e["m"], rate=e["rate"], family=e["family"])
                  for e in listing["entries"]), key=key)
    assert got == ref
    assert [(m.family, m.rate, m.n) for m in cat.matrices] == sorted((m.family, m.rate, m.n) for m in cat.matrices)
    bch = cat.get_by_family("bch")[0]
    assert (bch.n, bch.k) == (7, 4)
    w = cat.get_nearest_rate(0.5, family="wimax", block_size=2304)
    assert w.name == "wimax_2304_0.5.alist.txt"
    assert cat.get_higher_rate(w).rate == pytest.approx(0.66) and cat.get_lower_rate(w) is None
    assert all(0.7 <= m.rate <= 0.8 for m in cat.get_by_rate_range(0.7, 0.8))


# ---- results ------------------------------------------------------------------------------
def _sample_result():
    path = "codes/wimax_576_0.5.alist.txt"
    cfg = SimulationConfig(matrix_path=path, n=576, m=288, k=288, rate=0.5, blocks=1000, max_iterations=20,
                           encoding_method="standard", interleaver_type="none", decoder_type="sumproduct",
                           channel_mode=1, modulation=1, speed=0.5, snr_range=(0.0, 2.0, 1.0), threads=1,
                           timestamp="2026-10-18T12:00:00", interference_snr=1.0, p=0.1)
    pts = [SNRPointResult(snr_db=float(s), ber=b, fer=f, avg_normalized_llr=0.0, total_blocks=1000,
                          successful_blocks=1000 - int(f * 1000), failed_blocks=int(f * 1000),
                          avg_convergence_iterations=c, matrix_path=path, modulation=1, max_iterations=20,
                          interleaver="none", encoding_method="standard")
           for s, b, f, c in ((0, 0.0917, 1.0, 0.0), (1, 0.031415, 0.5, 7.25), (2, 1e-7, 0.001, 3.0))]
    return SimulationResult(config=cfg, snr_points=pts, wall_clock_seconds=12.5,
                            adaptation_log=[{"snr_db": 1.0, "action": "Увеличение итераций"}])


def test_results_writers_are_byte_identical_to_the_reference(tmp_path):
    res = _sample_result()
    res.to_json(str(tmp_path / "r.json"))
    res.to_csv(str(tmp_path / "r.csv"))
    assert (tmp_path / "r.json").read_bytes() == open(os.path.join(GOLDEN, "results_sample.json"), "rb").read()
    assert (tmp_path / "r.csv").read_bytes() == open(os.path.join(GOLDEN, "results_sample.csv"), "rb").read()
    back = SimulationResult.from_json(os.path.join(GOLDEN, "results_sample.json"))
    assert back == res and back.config.snr_range == (0.0, 2.0, 1.0)
    empty = SimulationResult(config=res.config, snr_points=[], wall_clock_seconds=0.0)
    empty.to_csv(str(tmp_path / "none.csv"))
    assert not (tmp_path / "none.csv").exists()                  # results.py:84-85


# ---- small carriers ---------------------------------------------------------------------------
def test_settings_defaults_and_enums():
    s = Settings()
    assert (s.get_blocks_cnt(), s.get_max_iterations(), s.get_s_param()) == (100, 5, -1)
    assert s.get_interleaver_type() == InterleaverType.NONE and s.get_decoder_type() == LDPCDecoderType.BIT_FLIPPING
    assert s.is_ber_calculate() and not s.is_fer_calculate() and not s.is_normalized_llr_calculate()
    assert s.get_interleaver_type_name() == "None"
    s.set_interleaver_type(InterleaverType.SRANDOM)
    assert s.get_interleaver_type_name() == "S-Random"
    assert s.get_precision() == "f64" and s.is_early_termination() and not s.is_fix_odd_check_sign()
    with pytest.raises(ValueError):
        s.set_precision("bf16")
    assert Result.OK.value == "eOk" and Result.DATA_TRANSFER_NOT_OK.value == "eDataTransferNotOk"
    assert Result.from_flag(1) is Result.OK and Result.from_flag(0) is Result.DATA_TRANSFER_NOT_OK


def test_channel_matches_oracle_formula():
    from oracle import spa_oracle as so
    ch = Channel.create_channel(0.5, 2.0, 0.0, 1, 0.1, 1)
    assert ch.gen_ptr.sigma == pytest.approx(so.sigma(0.5, 2.0))
    assert ch.L_c1 == pytest.approx(4 * 0.5 * 10 ** 0.2)
    bits = np.random.default_rng(3).integers(0, 2, size=(5, 64))
    ch.seed(1234)
    got = ch.process_batch(bits)
    g = np.random.RandomState(1234).normal(0.0, 1.0, size=bits.shape)
    np.testing.assert_allclose(got, so.channel_llr(bits.astype(np.uint8), g, ch.gen_ptr.sigma, True), rtol=1e-12)
    ch.sigma_sq_quirk = False
    ch.seed(1234)
    np.testing.assert_allclose(ch.process_batch(bits), so.channel_llr(bits.astype(np.uint8), g, ch.gen_ptr.sigma, False),
                               rtol=1e-12)
    buf = DataBuffer(0)
    buf._encoded_data = [0, 1, 1, 0]
    ch.process(buf); ch.process(buf)
    assert len(buf._channel_data) == 8                         # appends (channel.py:81)
    with pytest.raises(NotImplementedError):                   # modes 2/3 exist per frame on the host only
        Channel.create_channel(0.5, 2.0, 1.0, 2, 0.1, 1).process_batch(bits)


def test_lcg_generator_is_park_miller():
    g = Generator(1, 1.0)
    g.ran()
    assert g.idum == 16807
    for _ in range(9999):
        g.ran()
    assert g.idum == 1043618065                                  # the classic minimal-standard check value
    assert all(b in (0, 1) for b in Generator.generate_bit_sequence(50))


def test_snr_grid_and_frame_split():
    from main import snr_grid
    from mc_driver import split_frames, wilson_interval
    assert snr_grid(0.0, 2.0, 1.0) == [0.0, 1.0, 2.0]
    assert snr_grid(0.0, 5.0, 0.5)[-1] == 5.0 and len(snr_grid(0.0, 5.0, 0.5)) == 11
    assert snr_grid(0.0, 1.0, 0.4) == [0.0, 0.4, 0.8, 1.0]       # clamped last point (main.py:207-209)
    for total in (0, 1, 7, 64, 1000):
        for world in (1, 2, 3, 8):
            parts = [split_frames(total, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == total
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
    lo, hi = wilson_interval(10243, 80000)
    assert lo == pytest.approx(0.12574, abs=2e-4) and hi == pytest.approx(0.13037, abs=2e-4)


# ---- adaptive controller (policy + catalog walks; the sweep itself runs on the GPU) ------------
def test_adaptive_strategy_equals_the_reference_on_a_grid_of_points():
    """ThresholdStrategy.evaluate (adaptive.py:61-124): 1400 (BER, FER, convergence, budget, interleaver)
    points x 2 threshold sets evaluated by the unmodified reference (make_golden.py: adaptive)."""
    from adaptive import AdaptiveState, ThresholdStrategy
    from results import SNRPointResult
    with open(os.path.join(GOLDEN, "adaptive_strategy.json")) as f:
        gold = json.load(f)
    strategies = [ThresholdStrategy(**kw) for kw in gold["strategies"]]
    assert strategies[0].get_name() == "threshold"
    for c in gold["cases"]:
        state = AdaptiveState("x/wimax_576_0.5.alist.txt", 0.5, 1, c["max_it"], c["interleaver"], "standard")
        pt = SNRPointResult(snr_db=1.0, ber=c["ber"], fer=c["fer"], avg_normalized_llr=0.0, total_blocks=100,
                            successful_blocks=50, failed_blocks=50, avg_convergence_iterations=c["conv"])
        act = strategies[c["strategy"]].evaluate(state, pt)
        want = c["action"]
        if want is None:
            assert act is None, c
        else:
            got = dict(matrix=act.new_matrix_path, modulation=act.new_modulation, max_iterations=act.new_max_iterations,
                       interleaver=act.new_interleaver, reason=act.reason)
            assert got == want, c


def test_adaptive_rate_ladder_walks_equal_the_reference(tmp_path, capsys):
    """AdaptiveController._apply_action (:384-412) against the catalog of the reference's database."""
    from adaptive import AdaptiveAction, AdaptiveController, AdaptiveState, ThresholdStrategy
    with open(os.path.join(GOLDEN, "catalog_listing.json")) as f:
        listing = json.load(f)
    for e in listing["entries"]:
        p = tmp_path / e["rel"]
        p.parent.mkdir(parents=True, exist_ok=True)
        p.write_text(listing["first_lines"][e["rel"]] + "\n")
    ctl = AdaptiveController(ThresholdStrategy(), MatrixCatalog(str(tmp_path)))
    ctl._get_encoder_decoder_data = lambda path: None
    with open(os.path.join(GOLDEN, "adaptive_strategy.json")) as f:
        walks = json.load(f)["walks"]
    assert len(walks) == 8
    for w in walks:
        path = str(tmp_path / w["start"])
        info = ctl._find_current_matrix_info(path)
        state = AdaptiveState(path, info.rate if info else 0.0, 1, 5, "none", "standard")
        for rel, rate in w["trail"]:
            ctl._apply_action(AdaptiveAction(new_matrix_path=w["direction"]), state, None, None, None)
            assert (os.path.relpath(state.current_matrix_path, tmp_path), state.current_rate) == (rel, rate), w
    # the other fields of an action are plain assignments (:404-412)
    ctl._apply_action(AdaptiveAction(new_max_iterations=40, new_interleaver="random", new_modulation=2), state, None, None, None)
    assert (state.current_max_iterations, state.current_interleaver, state.current_modulation) == (40, "random", 2)
    capsys.readouterr()


def test_channel_modes_2_and_3_are_bit_identical_to_the_reference():
    """channel.py:83-100: interference modes on the Park-Miller generators (+ numpy's global stream for the
    hit decision of mode 2), both modulations; two consecutive frames so the generator state carries over."""
    from channel import Channel
    with open(os.path.join(GOLDEN, "channel_modes.json")) as f:
        cases = json.load(f)["cases"]
    assert {c["mode"] for c in cases} == {2, 3}

    class Buf:
        def __init__(self, bits):
            self._encoded_data, self._channel_data = list(bits), []

    for c in cases:
        ch = Channel.create_channel(c["speed"], c["sn1"], c["sn2"], c["mode"], c["p"], c["modulation"])
        assert [ch.L_c1, ch.L_c2, ch.L_c3] == c["L_c"] and [ch.gen_ptr.sigma, ch.gen_ptr2.sigma] == c["sigma"]
        np.random.seed(c["numpy_seed"])
        for want in c["llr"]:
            buf = Buf(c["bits"])
            ch.process(buf)
            assert buf._channel_data == want, (c["mode"], c["modulation"])


def test_modulation_2_uses_symbols_of_amplitude_07():
    from channel import Channel
    a = Channel.create_channel(0.5, 2.0, 0.0, 1, 0.1, 1); a.seed(3)
    b = Channel.create_channel(0.5, 2.0, 0.0, 1, 0.1, 2); b.seed(3)
    bits = np.array([[0, 1, 1, 0, 1]])
    la, lb = a.process_batch(bits), b.process_batch(bits)
    s2 = a.gen_ptr.sigma ** 2
    np.testing.assert_allclose((la - lb) * s2 / 2.0, np.where(bits == 0, -0.3, 0.3), atol=1e-12)
    from oracle import spa_oracle as oracle
    want = oracle.channel_llr(bits.ravel(), (lb.ravel() * s2 / 2.0 - np.where(bits.ravel() == 0, -0.7, 0.7)) / s2,
                              a.gen_ptr.sigma, True, amp=0.7)
    np.testing.assert_allclose(lb.ravel(), want, rtol=1e-12)
