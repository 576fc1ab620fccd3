"""Accuracy harness (simulator/trueReadLevels.cpp:18-196 in spirit): the generator records the true first / last PRG level of
every read; the alignments the path selects must put the read ends there. Run on the oracle here (the GPU tier proves product == oracle)."""
import numpy as np
import pytest

import harness as H


@pytest.mark.parametrize("name,floor", [("S", 0.93), ("genes", 0.95)])
def test_read_ends_land_on_their_true_levels(dataset, name, floor):
    d, b, mu, sd = dataset(name)
    o = H.Oracle(d).pairs(b, mu, sd, 1024)
    n = o["n_cols"]
    first = np.array([next((l for l in o["level"][r, :n[r]] if l != -1), -1) for r in range(len(n))])
    last = np.array([next((l for l in o["level"][r, :n[r]][::-1] if l != -1), -1) for r in range(len(n))])
    ok = (first == b["truth_first"]) & (last == b["truth_last"])
    assert ok.mean() >= floor, "only %.3f of the reads have both ends on their true levels" % ok.mean()
    assert (last == b["truth_last"]).mean() >= 0.99


# ---- the reference's own harness: `.levels` files + trueReadLevels::evaluateAlignment

import os          # noqa: E402
import subprocess  # noqa: E402
import sys         # noqa: E402

COMPARE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "truth_ref_compare.py")


@pytest.mark.skipif(not H.have_ref(), reason="compiled reference (oracle/_ref) not built on this box")
def test_truth_evaluation_matches_the_references_trueReadLevels():
    """hlala_truth_load / hlala_truth_evaluate against the unmodified trueReadLevels constructor + evaluateAlignment inside the reference's own
    alignOneReadPair (own process), per pair, on the compiled reference's alignments; all pairs once more against a plain Python count"""
    r = subprocess.run([sys.executable, COMPARE], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0 and "ok:" in r.stdout, r.stdout[-3000:]


@pytest.mark.gpu
@pytest.mark.skipif(not H.have_ref(), reason="compiled reference (oracle/_ref) not built on this box")
def test_truth_evaluation_of_the_gpu_alignments_matches_the_reference():
    """the same comparison with the alignments hlala_align_pairs produces on the GPU: the accuracy figure of the product equals the reference's"""
    r = subprocess.run([sys.executable, COMPARE, "--gpu"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0 and "ok:" in r.stdout and "hlala_align_pairs (GPU)" in r.stdout, r.stdout[-3000:]


def test_levels_files_follow_the_references_layout(dataset, tmp_path):
    """hlala-synth --levels-prefix: six lines per read in sequencing orientation (a reverse-strand read reverse-complemented in the FASTQ, its levels and
    true-alignment characters reversed only, readSimulator.cpp:1549,1618-1630), consistent with the seed batch"""
    d, _b, _mu, _sd = dataset("small")
    pre = str(tmp_path / "R")
    _prg_kw, rd_kw, _m, _s = __import__("conftest").DATASETS["small"]
    b = H.synth_reads(d, str(tmp_path / "seeds.bin"), levels_prefix=pre, **rd_kw)
    comp = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}
    for m in (0, 1):
        lv = H.read_levels_file("%s_%d.levels" % (pre, m + 1))
        fq = open("%s_%d.fq" % (pre, m + 1)).read().split("\n")
        assert len(lv) == rd_kw["pairs"] and len(fq) == 4 * rd_kw["pairs"] + 1
        for p_ in range(rd_kw["pairs"]):
            r = 2 * p_ + m; s, e = b["read_off"][r], b["read_off"][r + 1]
            bases = bytes(b["bases"][s:e]).decode(); truth = list(b["truth_levels"][s:e])
            prim = [c for c in range(b["chain_off"][r], b["chain_off"][r + 1]) if not b["chain_flag"][c] & 0x100]
            rev = bool(b["chain_flag"][prim[0]] & 0x10)
            levels, labels, fa_levels, fa_labels, fa_seq = lv["r%d" % p_]
            assert fq[4 * p_] == "@r%d" % p_ and fq[4 * p_ + 2] == "+"
            assert fq[4 * p_ + 1] == ("".join(comp[c] for c in reversed(bases)) if rev else bases)
            assert fq[4 * p_ + 3] == bytes(b["quals"][s:e][::-1] if rev else b["quals"][s:e]).decode()
            assert levels == (truth[::-1] if rev else truth) and fa_levels == levels and fa_seq == (bases[::-1] if rev else bases)
            assert len(labels) == len(levels) == len(fa_labels) and all((l == -1) == (c == "_") for l, c in zip(levels, labels))
            assert truth[0] == b["truth_first"][r] and [x for x in truth if x != -1][-1] == b["truth_last"][r]


def test_levels_parser_edge_cases(tmp_path):
    """blank lines and CR LF are skipped like the reference's eraseNL / empty-line rule; a record whose true-alignment lines differ in length, a line that does
    not start with @ and a missing file are errors (asserts in the reference, trueReadLevels.cpp:207,219,311-312)"""
    import ctypes as C
    L = C.CDLL(H.LIB_PRODUCT); L.hlala_last_error.restype = C.c_char_p; L.hlala_truth_n_reads.restype = C.c_int64
    good = tmp_path / "a_1.levels"; good.write_text("\n@q1\r\n5 6 -1 7\r\nACGT\r\n5 6 -1 7\r\nAC_T\r\nACGT\r\n\n@q2\n9\nA\n9\nA\nA\n")
    t = C.c_void_p()
    assert L.hlala_truth_load(str(good).encode(), None, C.byref(t)) == 0 and L.hlala_truth_n_reads(t) == 2
    # one unpaired "pair": mate 2 has no levels and no aligned bases
    import numpy as np
    aln = dict(n_cols=np.array([5, 0], np.int32), level=np.array([[5, 6, -1, 7, 8], [0] * 5], np.int32), schar=np.frombuffer(b"ACG_T\0\0\0\0\0", np.uint8).reshape(2, 5).copy(),
               read_reverse=np.zeros(2, np.uint8))
    L.hlala_truth_free(t)
    per, tot, n = H.truth_evaluate(aln, str(good), None, names=["q1"])
    assert per.tolist() == [[4, 3]] and tot.tolist() == [4, 3, 1] and n == 2      # the aligned G sits on level -1 like the truth, T on 8 instead of 7
    bad = tmp_path / "b_1.levels"; bad.write_text("@q1\n5 6\nAC\n5 6\nAC\nACG\n")
    assert L.hlala_truth_load(str(bad).encode(), None, C.byref(t)) != 0 and b"different lengths" in L.hlala_last_error()
    bad.write_text("q1\n5\nA\n5\nA\nA\n")
    assert L.hlala_truth_load(str(bad).encode(), None, C.byref(t)) != 0 and b"starting with @" in L.hlala_last_error()
    assert L.hlala_truth_load(str(tmp_path / "none.levels").encode(), None, C.byref(t)) != 0 and b"cannot open" in L.hlala_last_error()
    with pytest.raises(RuntimeError, match="aligned bases"):
        H.truth_evaluate(dict(aln, n_cols=np.array([3, 0], np.int32)), str(good), None, names=["q1"])
