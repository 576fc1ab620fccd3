"""Run by tests/test_truth_levels.py in a process of its own (one compiled-reference graph per process): the product's accuracy harness
(hlala_truth_load / hlala_truth_evaluate, host only) against the UNMODIFIED simulator::trueReadLevels (oracle/_ref): the reference parses
the same `.levels` files with its own constructor (simulator/trueReadLevels.cpp:203-320) and counts inside its own alignOneReadPair
(mapper/processBAM.cpp:3555-3560 -> evaluateAlignment, trueReadLevels.cpp:18-196); the product counts on the reference's alignments of
the same pairs. Equal per pair, not only in total. With --gpu the product's own alignments (hlala_align_pairs) are counted instead.
Pairs with a read below 90 % cannot go through the reference (its diagnostic print of such reads is undefined behaviour, see
oracle/ref_driver.cpp run_truth_pairs); for those, and for all others once more, the count is checked against plain Python."""
import json
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import harness as H  # noqa: E402


def python_count(aln, truth, npairs):
    per = np.zeros((npairs, 2), np.int64); low = np.zeros(npairs, bool)
    for p in range(npairs):
        for m in (0, 1):
            r = 2 * p + m; n = aln["n_cols"][r]
            lv = aln["level"][r, :n][aln["schar"][r, :n] != ord("_")]
            t = np.array(truth[m]["r%d" % p][0]); t = t[::-1] if aln["read_reverse"][r] else t
            assert len(t) == len(lv)
            ok = int((lv == t).sum()); per[p] += (len(t), ok); low[p] |= ok / len(t) < 0.9
    return per, low


def main():
    gpu = "--gpu" in sys.argv
    d = tempfile.mkdtemp(prefix="truth_ref_")
    H.synth_prg(d, levels=25000, haps=4, genes=2, alleles=64, seed=5)
    pre = os.path.join(d, "R"); N = 1500
    b = H.synth_reads(d, os.path.join(d, "seeds.bin"), pairs=N, len=100, clip_frac=0.2, indel_rate=0.004, gene_frac=0.3, seed=5, levels_prefix=pre)
    r1, r2 = pre + "_1.levels", pre + "_2.levels"
    R = H.quiet(H.Ref, d)
    if gpu:
        P = H.Product(d); P.to_gpu(0); aln = P.pairs(b, 100.0, 10.0, 1024)
    else:
        aln = H.quiet(R.pairs, b, 100.0, 10.0, 1024)
    got_per, got_tot, n_ids = H.truth_evaluate(aln, r1, r2)
    assert n_ids == N, n_ids
    py_per, low = python_count(aln, (H.read_levels_file(r1), H.read_levels_file(r2)), N)
    assert (got_per == py_per).all() and got_tot[0] == py_per[:, 0].sum() == int(np.diff(b["read_off"]).sum()) and got_tot[1] == py_per[:, 1].sum()
    good = np.nonzero(~low)[0]
    assert len(good) > 0.8 * N
    run = (R.truth_pairs if "--verbose" in sys.argv else lambda *a: H.quiet(R.truth_pairs, *a))
    want_per, want_tot = run(H.subset_pairs(b, good), 100.0, 10.0, r1, r2, good)
    assert (got_per[good] == want_per).all(), "per-pair counts differ at pairs %s" % good[np.nonzero((got_per[good] != want_per).any(axis=1))[0][:10]]
    assert (want_tot == want_per.sum(axis=0)).all()
    frac = want_tot[1] / want_tot[0]
    assert frac < 1.0, "the pairs the reference counted should hold some misplaced bases, else the comparison proves little"
    # explicit names, and the reference's assert on unknown IDs
    got2, _t, _n = H.truth_evaluate(aln, r1, r2, names=["r%d" % i for i in range(N)])
    assert (got2 == got_per).all()
    try:
        H.truth_evaluate(aln, r1, r2, names=["x%d" % i for i in range(N)])
        raise AssertionError("unknown read IDs must fail")
    except RuntimeError as e:
        assert "hlala error -5" in str(e) and "no true levels for read x0" in str(e), e
    print("ok: " + json.dumps(dict(pairs=N, pairs_through_the_reference=int(len(good)), bases=int(want_tot[0]), on_true_level=int(want_tot[1]), misplaced=int(want_tot[0] - want_tot[1]),
                                   all_pairs_fraction=round(float(got_tot[1] / got_tot[0]), 5), reads_below_90=int(got_tot[2]),
                                   alignments="hlala_align_pairs (GPU)" if gpu else "compiled reference")))


if __name__ == "__main__":
    main()
