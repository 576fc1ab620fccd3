"""Read simulator of the reference (simulator/readSimulator.cpp:58-334, 1194-1693) behind hlala_simulate_read_pairs: host only."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

import harness as H

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import simulator_ref_compare as S  # noqa: E402


@pytest.mark.skipif(not H.have_ref(), reason="compiled reference (oracle/_ref) not built on this box")
def test_simulated_reads_match_the_references_readSimulator():
    """R_1.fq / R_2.fq / R_1.levels / R_2.levels byte for byte and the average error rates exactly, against the unmodified readSimulator (own process): exact and
    interpolated read lengths, removal of upper quality classes for either read, insertions / deletions from quality-0 rows, includeDeletions, perfect reads,
    appending a second haplotype; on synthetic matrices and, where /root/reference is present, on the matrix the reference ships. Also hlala_simulate_individual against
    the unmodified HLATyper::simulateOneIndividual (all seven files of four individuals) and hlala_simulate_from_graph against simulator::simulateFromGraph with its random
    diploid walks (five files of three runs), byte for byte"""
    r = subprocess.run([sys.executable, os.path.join(HERE, "simulator_ref_compare.py")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0 and "ok:" in r.stdout, r.stdout[-3000:]


def _lib():
    L = C.CDLL(H.LIB_PRODUCT); L.hlala_simulate_read_pairs.restype = C.c_int64; L.hlala_last_error.restype = C.c_char_p
    return L


def _run(L, mat, rl, path, out, ip=0, d1=1, d2=0, cov=20.0, mu=250.0, sd=30.0, perfect=0, incdel=0, prefix="s_", app=0):
    er = np.zeros(2)
    n = L.hlala_simulate_read_pairs(mat.encode(), C.c_int(rl), C.c_int(ip), C.c_int(d1), C.c_int(d2), H.p(path), C.c_int64(len(path)), C.c_double(cov), C.c_double(mu), C.c_double(sd),
                                    C.c_int(perfect), C.c_int(incdel), prefix.encode(), out.encode(), C.c_int(app), H.p(er))
    return n, er


def test_simulated_reads_are_consistent_with_their_path(tmp_path):
    """perfect reads copy the path: every base equals the path character at its level, levels skip the path's gaps, the second read is the reverse complement with its
    levels in sequencing order, names carry r<n>|||<start>|||<jump>, both mates share the name; the call is deterministic (default-seeded generator as in the reference)"""
    L = _lib(); mat = str(tmp_path / "m.txt"); S.synthetic_matrix(mat)
    path = S.random_path(6000, 11); out = str(tmp_path / "R")
    n, er = _run(L, mat, 101, path, out, perfect=1)
    assert n > 300 and 0 < er[0] < 0.5
    comp = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}
    fq = [open(out + "_%d.fq" % m).read().split("\n") for m in (1, 2)]
    lv = [H.read_levels_file(out + "_%d.levels" % m) for m in (1, 2)]
    assert len(fq[0]) == 4 * n + 1 and len(lv[0]) == n
    n_rev_first = 0
    for i in range(n):
        name = fq[0][4 * i][1:]
        assert fq[1][4 * i][1:] == name and name.startswith("s_r%d|||" % (i + 1)) and len(name.split("|||")) == 3
        seqs = [fq[m][4 * i + 1] for m in (0, 1)]
        for m in (0, 1):
            levels, labels, fa_levels, fa_labels, fa_seq = lv[m][name]
            assert len(levels) == 101 == len(seqs[m]) and -1 not in levels and fa_levels == levels and fa_labels == labels
            rev = levels[0] > levels[-1]
            if m == 0:
                n_rev_first += rev
            want = "".join(chr(path[l]) for l in levels)
            assert labels == want and fa_seq == want
            assert seqs[m] == ("".join(comp[c] for c in want) if rev else want)
            assert all(path[l] != ord("_") for l in levels) and sorted(levels, reverse=rev) == levels
        assert (lv[0][name][0][0] > lv[0][name][0][-1]) != (lv[1][name][0][0] > lv[1][name][0][-1])     # one mate per strand
    assert 0.3 * n < n_rev_first < 0.7 * n                                                             # the mates swap places with probability 1/2
    out2 = str(tmp_path / "R2"); n2, _ = _run(L, mat, 101, path, out2, perfect=1)
    assert n2 == n and open(out2 + "_1.levels").read() == open(out + "_1.levels").read() and open(out2 + "_2.fq").read() == open(out + "_2.fq").read()
    n3, _ = _run(L, mat, 101, path, out2, perfect=1, app=1)
    assert open(out2 + "_2.fq").read() == 2 * open(out + "_2.fq").read()


def test_simulator_errors_are_reported(tmp_path):
    """what the reference answers with a throw or an assert: missing matrix, wrong header, read length without a table (interpolation off), a table of longer reads,
    a path shorter than a read, more quality classes to remove than a position has"""
    L = _lib(); mat = str(tmp_path / "m.txt"); S.synthetic_matrix(mat)
    path = S.random_path(3000, 12); out = str(tmp_path / "R")
    assert _run(L, str(tmp_path / "none.txt"), 101, path, out)[0] < 0 and b"cannot open the quality matrix" in L.hlala_last_error()
    bad = str(tmp_path / "bad.txt"); open(bad, "w").write("readLength\tquality\tpositionInRead\tN\tExpectedCorrect\tEmpiricalCorrect\n")
    assert _run(L, bad, 101, path, out)[0] < 0 and b"is not qualityScore" in L.hlala_last_error()
    assert _run(L, mat, 120, path, out, ip=0)[0] < 0 and b"holds no reads of length 120" in L.hlala_last_error()
    assert _run(L, mat, 80, path, out, ip=1)[0] < 0 and b"beyond the read length" in L.hlala_last_error()
    assert _run(L, mat, 101, path[:50].copy(), out)[0] < 0 and b"shorter than one read" in L.hlala_last_error()
    assert _run(L, mat, 101, path, out, d1=4)[0] < 0 and b"cannot remove 4 quality classes" in L.hlala_last_error()
    assert _run(L, mat, 101, path, out, d1=21)[0] < 0 and b"0..20" in L.hlala_last_error()
    assert _run(L, mat, 120, path, out, ip=1)[0] > 0


def test_simulated_individual_closes_the_loop_with_the_type_evaluation(dataset, tmp_path):
    """hlala_simulate_individual (HLATyper::simulateOneIndividual; pinned byte for byte in simulator_ref_compare.py): two types per gene, reads of 101 bases from the
    gene haplotypes, and a truth file that hlala_evaluate_types (--trueHLA) reads back: calls equal to the truth score 2 of 2 at every locus, a wrong call 1 of 2"""
    from test_evaluate_types import product_eval
    L = _lib(); L.hlala_simulate_individual.restype = C.c_int64
    d, _b, _mu, _sd = dataset("small")
    mat = str(tmp_path / "m.txt"); S.synthetic_matrix(mat)
    out = str(tmp_path / "ind"); os.makedirs(out)
    types = C.create_string_buffer(4096)
    n = L.hlala_simulate_individual(d.encode(), mat.encode(), out.encode(), C.c_double(250.0), C.c_double(25.0), C.c_int(0), C.c_int(1), C.c_uint(5), types, C.c_int64(4096))
    assert n > 0, L.hlala_last_error()
    chosen = dict(x.split(":", 1) for x in types.value.decode().split(";"))
    assert list(chosen) == ["A"] and all(t.startswith("A*") for t in chosen["A"].split("/"))
    for f in S.FILES:
        assert os.path.getsize(os.path.join(out, f)) > 0
    assert open(os.path.join(out, "R_1.fq")).read().count("\n") == 4 * n
    assert open(os.path.join(out, "R_1.fq")).readline().startswith("@PRG_A_HAPLO_0r1|||")
    t1, t2 = chosen["A"].split("/")
    best = str(tmp_path / "best.txt")
    open(best, "w").write("Locus\tChromosome\tAllele\tQ1\tQ2\nA\t1\t%s\t1\t1\nA\t2\t%s\t1\t1\n" % (t2, t1))
    got, _ = product_eval(out, best, os.path.join(out, "HLAtypes.txt"))      # the individual's ID is the output directory (HLATyper.cpp:899)
    assert got == {"A": (2, 2)}
    open(best, "w").write("Locus\tChromosome\tAllele\tQ1\tQ2\nA\t1\t%s\t1\t1\nA\t2\tA*99:99\t1\t1\n" % t1)
    got, _ = product_eval(out, best, os.path.join(out, "HLAtypes.txt"))
    assert got == {"A": (2, 2 if t1 == t2 else 1)}
    # same seed, same individual; no matrix, no simulation (assert(rS != 0) in the reference)
    out2 = str(tmp_path / "ind2"); os.makedirs(out2)
    assert L.hlala_simulate_individual(d.encode(), mat.encode(), out2.encode(), C.c_double(250.0), C.c_double(25.0), C.c_int(0), C.c_int(1), C.c_uint(5), None, C.c_int64(0)) == n
    assert open(os.path.join(out2, "R_2.levels")).read() == open(os.path.join(out, "R_2.levels")).read()
    assert L.hlala_simulate_individual(d.encode(), b"", out2.encode(), C.c_double(250.0), C.c_double(25.0), C.c_int(0), C.c_int(1), C.c_uint(5), None, C.c_int64(0)) < 0
    assert L.hlala_simulate_individual((d + "_nowhere").encode(), mat.encode(), out2.encode(), C.c_double(250.0), C.c_double(25.0), C.c_int(0), C.c_int(1), C.c_uint(5), None, C.c_int64(0)) < 0


def test_reads_simulated_from_the_graph_feed_the_truth_harness(dataset, tmp_path):
    """hlala_simulate_from_graph (simulator::simulateFromGraph; pinned byte for byte in simulator_ref_compare.py): random diploid walks through the loaded graph, reads
    with graph levels; hlala_truth_load takes the `.levels` files, and an alignment that puts every base on its simulated level scores 100 %"""
    d, _b, _mu, _sd = dataset("small")
    G = H.Product(d); L = G.lib; L.hlala_simulate_from_graph.restype = C.c_int64; L.hlala_truth_n_reads.restype = C.c_int64
    mat = str(tmp_path / "m.txt"); S.synthetic_matrix(mat)
    out = str(tmp_path / "g"); os.makedirs(out)
    n = L.hlala_simulate_from_graph(G.g, b"label", mat.encode(), C.c_int(101), C.c_double(200.0), C.c_double(20.0), C.c_int(1), out.encode(), C.c_double(4.0), C.c_int(1), C.c_uint(2))
    assert n > 50, L.hlala_last_error()
    assert open(os.path.join(out, "parameters.txt")).read().startswith("Graph: label\nsimulatedGraphGenomes: 1\n")
    r1, r2 = os.path.join(out, "R_1.levels"), os.path.join(out, "R_2.levels")
    lv = [H.read_levels_file(r1), H.read_levels_file(r2)]
    names = list(lv[0]); assert len(names) == n and names[0].startswith("PRG_0h1_r1|||") and any(x.startswith("PRG_0h2_") for x in names)
    nl = G.dims()["n_levels"]
    assert all(-1 <= l < nl - 1 for x in lv[0].values() for l in x[0])
    # a "perfect aligner": columns = the simulated bases on their simulated levels, in graph direction
    cap = 128; aln = dict(n_cols=np.zeros(2 * n, np.int32), level=np.zeros((2 * n, cap), np.int32), schar=np.zeros((2 * n, cap), np.uint8), read_reverse=np.zeros(2 * n, np.uint8))
    for p_, nm in enumerate(names):
        for m in (0, 1):
            levels = lv[m][nm][0]; known = [l for l in levels if l != -1]; rev = known[0] > known[-1]
            aln["read_reverse"][2 * p_ + m] = rev; aln["n_cols"][2 * p_ + m] = len(levels)
            aln["level"][2 * p_ + m, :len(levels)] = levels[::-1] if rev else levels; aln["schar"][2 * p_ + m, :len(levels)] = ord("A")
    per, tot, n_ids = H.truth_evaluate(aln, r1, r2, names=names)
    assert n_ids == n and tot[0] == tot[1] == 2 * 101 * n and tot[2] == 0
    aln["level"][0, 3] += 1
    per, tot, _ = H.truth_evaluate(aln, r1, r2, names=names)
    assert tot[0] - tot[1] == 1 and per[0].tolist() == [202, 201]
    # a graph handle is needed, seeds differ, errors are reported
    n2 = L.hlala_simulate_from_graph(G.g, b"label", mat.encode(), C.c_int(101), C.c_double(200.0), C.c_double(20.0), C.c_int(1), out.encode(), C.c_double(4.0), C.c_int(1), C.c_uint(3))
    assert n2 > 0 and H.read_levels_file(r1) != lv[0]
    assert L.hlala_simulate_from_graph(None, b"", mat.encode(), C.c_int(101), C.c_double(200.0), C.c_double(20.0), C.c_int(1), out.encode(), C.c_double(4.0), C.c_int(1), C.c_uint(3)) == -1
    assert L.hlala_simulate_from_graph(G.g, b"", mat.encode(), C.c_int(101), C.c_double(200.0), C.c_double(20.0), C.c_int(1), (out + "/missing").encode(), C.c_double(4.0), C.c_int(1), C.c_uint(3)) < 0
    G.close()
