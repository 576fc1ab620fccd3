"""Run by tests/test_read_simulator.py in a process of its own: the product's read simulator (hlala_simulate_read_pairs, host only) against the
UNMODIFIED simulator::readSimulator (oracle/_ref: constructor + simulate_paired_reads_from_edgePath, simulator/readSimulator.cpp:58-334, 1194-1693)
on the same quality matrices and paths: the four output files (R_1.fq, R_2.fq, R_1.levels, R_2.levels) byte for byte, the average error rates exactly.
Boost.Random is absent here; the reference is compiled against the stand-in of oracle/shim (C++ standard library distributions), which is what the
product uses as well: the order and number of draws, the tables and all arithmetic are what this pins."""
import ctypes as C
import json
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import harness as H  # noqa: E402

REF_MATRIX = "/root/reference/simulator/predefinedQualityMatrices/I101_NA12878.txt"


def synthetic_matrix(path, lengths=(101,), seed=3, accurate=False):
    """quality classes with position-dependent counts and, unlike the file the reference ships, rows of quality character 0 (insertion / deletion counts);
    accurate: mostly high qualities (about 1 % errors, like a sequencing run) instead of about 15 %"""
    rng = np.random.RandomState(seed)
    with open(path, "wb") as f:
        f.write(b"readLength\tqualityScore\tpositionInRead\tN\tExpectedCorrect\tEmpiricalCorrect\n")
        for L in lengths:
            for pos in range(L):
                for qi, (q, acc) in enumerate(((b"#", 0.55), (b"+", 0.9), (b"5", 0.99), (b"?", 0.999), (b"I", 0.9999))):
                    n = int(rng.randint(0 if q == b"+" else 50, 5000)) * ((1, 2, 20, 200, 400)[qi] if accurate else 1)
                    f.write(b"%d\t%s\t%d\t%d\t%.6f\t%.9f\n" % (L, q, pos, n, acc, acc - rng.rand() * 0.01))
                if pos % 3 != 1 and pos > 0:   # no counts at position 0: with includeDeletions a read starting in a deletion trips an assert of the reference (:1654)
                    f.write(b"%d\t\x00\t%d\t%d\t0\t0\n" % (L, pos, int(rng.randint(2, 40))))
            f.write(b"\r\n")


def random_path(n, seed):
    rng = np.random.RandomState(seed)
    a = np.frombuffer(b"ACGT", np.uint8)[rng.randint(0, 4, n)].copy()
    for s in rng.randint(0, n - 60, 12):
        a[s:s + rng.randint(1, 50)] = ord("_")
    a[rng.randint(0, n, 8)] = ord("N")
    return a


FILES = ("R_1.fq", "R_2.fq", "R_1.levels", "R_2.levels", "HLAtypes.txt", "haplotypes.txt", "parameters.txt")


def individual(R, P, d, matrix):
    """hlala_simulate_individual against the unmodified HLATyper::simulateOneIndividual (hla/HLATyper.cpp:690-930) on a synthetic PRG with 3 genes x 12 alleles:
    all seven files byte for byte, for several seeds, with and without novel intron / exon recombinants, with and without sequencing errors"""
    prg = os.path.join(d, "prg"); H.synth_prg(prg, levels=14000, haps=4, genes=3, alleles=12, seed=3)
    ref = H.quiet(H.Ref, prg)
    R.hlala_ref_simulate_individual.restype = C.c_int; P.hlala_simulate_individual.restype = C.c_int64
    out = os.path.join(d, "ind"); os.makedirs(out)
    seen = set(); npairs = 0
    for seed, novel, err in ((1, 0, 1), (2, 1, 1), (7, 0, 0), (11, 1, 1)):
        rc = H.quiet(R.hlala_ref_simulate_individual, ref.h, prg.encode(), matrix.encode(), out.encode(), C.c_double(300.0), C.c_double(30.0), C.c_int(novel), C.c_int(err), C.c_uint(seed))
        assert rc == 0, R.hlala_ref_last_error()
        want = {f: open(os.path.join(out, f), "rb").read() for f in FILES}
        for f in FILES:
            os.remove(os.path.join(out, f))
        types = C.create_string_buffer(4096)
        n = P.hlala_simulate_individual(prg.encode(), matrix.encode(), out.encode(), C.c_double(300.0), C.c_double(30.0), C.c_int(novel), C.c_int(err), C.c_uint(seed), types, C.c_int64(4096))
        assert n > 0, P.hlala_last_error()
        for f in FILES:
            got = open(os.path.join(out, f), "rb").read()
            if got != want[f]:
                la, lb = want[f].split(b"\n"), got.split(b"\n")
                k = next((i for i in range(min(len(la), len(lb))) if la[i] != lb[i]), min(len(la), len(lb)))
                raise AssertionError("individual seed %d: %s differs at line %d:\n  ref %r\n  got %r" % (seed, f, k + 1, la[k][:300] if k < len(la) else None, lb[k][:300] if k < len(lb) else None))
        assert n == want["R_1.fq"].count(b"\n") // 4
        row = want["HLAtypes.txt"].split(b"\n")[1].split(b"\t")[1:]
        assert types.value.decode().split(";") == ["%s:%s" % (g, t.decode()) for g, t in zip("ABC", row)], (types.value, row)
        seen.add(tuple(row)); npairs += n
    assert len(seen) >= 3, "the seeds should lead to different individuals"
    # hlala_simulate_from_graph against the unmodified simulator::simulateFromGraph (random diploid walks through the same graph): five files byte for byte
    G = H.Product(prg); R.hlala_ref_simulate_from_graph.restype = C.c_int; P.hlala_simulate_from_graph.restype = C.c_int64
    gfiles = ("parameters.txt", "R_1.fq", "R_2.fq", "R_1.levels", "R_2.levels"); gpairs = 0
    for seed, genomes, cov, err in ((3, 1, 5.0, 1), (4, 2, 2.0, 1), (9, 1, 3.0, 0)):
        rc = H.quiet(R.hlala_ref_simulate_from_graph, ref.h, matrix.encode(), C.c_int(101), C.c_double(280.0), C.c_double(20.0), C.c_int(genomes), out.encode(), C.c_double(cov), C.c_int(err), C.c_uint(seed))
        assert rc == 0, R.hlala_ref_last_error()
        want = {f: open(os.path.join(out, f), "rb").read() for f in gfiles}
        for f in gfiles:
            os.remove(os.path.join(out, f))
        n = G.lib.hlala_simulate_from_graph(G.g, (prg + "/PRG/graph.txt").encode(), matrix.encode(), C.c_int(101), C.c_double(280.0), C.c_double(20.0), C.c_int(genomes), out.encode(),
                                            C.c_double(cov), C.c_int(err), C.c_uint(seed))
        assert n > 0, G.lib.hlala_last_error()
        for f in gfiles:
            got = open(os.path.join(out, f), "rb").read()
            if got != want[f]:
                la, lb = want[f].split(b"\n"), got.split(b"\n")
                k = next((i for i in range(min(len(la), len(lb))) if la[i] != lb[i]), min(len(la), len(lb)))
                raise AssertionError("from graph, seed %d: %s differs at line %d:\n  ref %r\n  got %r" % (seed, f, k + 1, la[k][:300] if k < len(la) else None, lb[k][:300] if k < len(lb) else None))
        assert n == want["R_1.fq"].count(b"\n") // 4
        gpairs += n
    # the levels in these files are graph levels: every simulated base names a level whose edges can emit it (perfect reads of the last run)
    lv = H.read_levels_file(os.path.join(out, "R_1.levels")); lo = G.array("level_edge_off"); em = G.array("edge_emis")
    for levels, labels, _a, _b, _c in list(lv.values())[:200]:
        for l, ch in zip(levels, labels):
            assert l == -1 or ord(ch) in em[lo[l]:lo[l + 1]]
    return dict(individuals=4, distinct_type_sets=len(seen), pairs=int(npairs), graph_runs=3, graph_pairs=int(gpairs))


def main():
    R = C.CDLL(H.LIB_REF); R.hlala_ref_simulate_pairs.restype = C.c_longlong; R.hlala_ref_last_error.restype = C.c_char_p
    P = C.CDLL(H.LIB_PRODUCT); P.hlala_simulate_read_pairs.restype = C.c_int64; P.hlala_last_error.restype = C.c_char_p
    d = tempfile.mkdtemp(prefix="sim_ref_")
    syn = os.path.join(d, "syn.txt"); synthetic_matrix(syn)
    syn2 = os.path.join(d, "syn2.txt"); synthetic_matrix(syn2, lengths=(60, 90), seed=8)
    cases = [  # matrix, read length, interpolate, drop, drop 2nd, path levels, path seed, coverage, mean, sd, perfectly, include deletions, prefix
        (syn, 101, 0, 1, 0, 7000, 1, 20.0, 250.0, 30.0, 0, 0, "PRG_0h1_"),
        (syn, 101, 0, 0, 0, 5000, 2, 15.0, 120.0, 40.0, 0, 1, ""),
        (syn, 130, 1, 2, 1, 6000, 3, 25.0, 200.0, 15.0, 0, 0, "x"),            # a table of longer reads cannot be used: its positions trip assert(positionInRead < readLength), :127
        (syn2, 100, 1, 1, 1, 6000, 4, 20.0, 300.0, 25.0, 0, 0, "h2_"),       # two tables (60 and 90): the nearer one is scaled
        (syn2, 150, 1, 1, 0, 8000, 5, 12.0, 400.0, 50.0, 1, 0, "perfect_"),
        (syn, 101, 0, 1, 0, 3000, 9, 60.0, 250.0, 30.0, 0, 0, "dense_"),
    ]
    if os.path.exists(REF_MATRIX):
        cases += [(REF_MATRIX, 101, 0, 1, 0, 9000, 6, 30.0, 350.0, 40.0, 0, 0, "PRG_1h2_"), (REF_MATRIX, 150, 1, 1, 2, 9000, 7, 30.0, 450.0, 60.0, 0, 0, "")]
    report = []
    for ci, (mat, L, ip, d1, d2, n, ps, cov, mu, sd, perf, incdel, pre) in enumerate(cases):
        path = random_path(n, ps)
        outs = []
        for which, lib, fn in (("ref", R, R.hlala_ref_simulate_pairs), ("got", P, P.hlala_simulate_read_pairs)):
            o = os.path.join(d, "%s%d" % (which, ci)); er = np.zeros(2, np.float64); tot = 0
            for app in (0, 1):   # the second call appends, as simulateFromGraph does for the second haplotype
                k = (fn if "--verbose" in sys.argv else lambda *a: H.quiet(fn, *a))(mat.encode(), C.c_int(L), C.c_int(ip), C.c_int(d1), C.c_int(d2), H.p(path), C.c_longlong(n), C.c_double(cov), C.c_double(mu), C.c_double(sd),
                            C.c_int(perf), C.c_int(incdel), pre.encode(), o.encode(), C.c_int(app), H.p(er))
                if k <= 0:   # the reference throws when a deletion is drawn exactly at the end of the path (std::string::at, readSimulator.cpp:1375): so must the product
                    tot = -1
                    break
                tot += k
            outs.append((tot, er.copy(), [open(o + s, "rb").read() for s in ("_1.fq", "_2.fq", "_1.levels", "_2.levels")] if tot > 0 else None))
        (nr, er_r, fr), (ng, er_g, fg) = outs
        if nr < 0 or ng < 0:
            assert nr == ng, "case %d: one side failed: ref %d, product %d (%s / %s)" % (ci, nr, ng, R.hlala_ref_last_error(), P.hlala_last_error())
            report.append(dict(case=ci, read_length=L, both_failed=P.hlala_last_error().decode()[:80]))
            continue
        assert nr == ng, (ci, nr, ng)
        assert (er_r == er_g).all(), (ci, er_r, er_g)
        for a, b, s in zip(fr, fg, ("R_1.fq", "R_2.fq", "R_1.levels", "R_2.levels")):
            if a != b:
                la, lb = a.split(b"\n"), b.split(b"\n")
                k = next(i for i in range(min(len(la), len(lb))) if la[i] != lb[i])
                raise AssertionError("case %d: %s differs at line %d:\n  ref %r\n  got %r" % (ci, s, k + 1, la[k][:200], lb[k][:200]))
        lv = fr[2].split(b"\n")
        report.append(dict(case=ci, read_length=L, pairs=int(nr), error_rates=[round(float(x), 6) for x in er_r], inserted_bases=int(sum(x.split(b" ").count(b"-1") for x in lv[1::6])),
                           deletion_columns=int(sum(x.count(b"_") for x in lv[5::6]))))
    done = [r for r in report if "pairs" in r]
    assert len(done) >= len(cases) - 2 and any(r["inserted_bases"] > 0 for r in done) and any(r["deletion_columns"] > 0 for r in done), report
    # the files feed the truth harness: the reference's own parser reads what the product wrote (checked through the product's parser, pinned elsewhere)
    t = C.c_void_p(); P.hlala_truth_n_reads.restype = C.c_int64
    assert P.hlala_truth_load(os.path.join(d, "got0_1.levels").encode(), os.path.join(d, "got0_2.levels").encode(), C.byref(t)) == 0
    assert "pairs" in report[0] and P.hlala_truth_n_reads(t) == report[0]["pairs"] // 2   # the appended half repeats the first half's names
    P.hlala_truth_free(t)
    report.append(individual(R, P, d, syn))
    print("ok: " + json.dumps(report))


if __name__ == "__main__":
    main()
