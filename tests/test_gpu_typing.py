"""Typing stage on the GPU through the C ABI (gene filter + compaction, per-read x cluster log-likelihoods, allele-pair sums, host
calls/QC/writers) against the typing oracle, which tests/test_typing_oracle.py pins byte-for-byte to the unmodified reference.
Per-read log-likelihoods and mismatch counts: bit-exact (north_star allows 1e-6). Allele-pair sums: 1e-12 relative (device exp/log
instead of glibc's; north_star 1e-6). Calls and every output file: identical text."""
import ctypes as C
import filecmp
import os

import numpy as np
import pytest

import harness as H

pytestmark = pytest.mark.gpu


def run_product(d, b, mu, sd, out_dir, world=1):
    P = H.Product(d); P.to_gpu(0)
    T = H.ProductTyping(P, d)
    if world == 1:
        sess = H.session_align(P, b, mu, sd, 512)
        blob, nsel = T.extract(sess)
        P.lib.hlala_session_free(sess)
        blobs = [blob]
    else:
        import sys
        sys.path.insert(0, H.PKG)
        import hlala_dist
        blobs = []; base = 0
        for r in range(world):
            sb = hlala_dist.shard_batch(b, r, world)
            sess = H.session_align(P, sb, mu, sd, 512)
            blob, nsel = T.extract(sess, base=base); blobs.append(blob)
            base += (len(sb["read_off"]) - 1) // 2
            P.lib.hlala_session_free(sess)
    return P, T, blobs


@pytest.mark.parametrize("name", ["typing", "typing1k"])
def test_typing_matches_oracle(dataset, tmp_path, name):
    d, b, mu, sd = dataset(name)
    P, T, blobs = run_product(d, b, mu, sd, None)
    out = str(tmp_path / "gpu" / "hla")
    T.infer(blobs, mu, sd, out)
    aln = P.pairs(b, mu, sd, 512, want_levels=False)       # the product's own alignments (parity-tested elsewhere) feed the oracle
    or_dir = str(tmp_path / "oracle" / "hla")
    O = H.OracleTyping(d, b, aln, mu, sd, or_dir)
    assert T.n_loci == O.n_loci == 17
    worst = 0.0
    for i in range(17):
        g = T.locus(i); o = O.locus(i)
        assert (g["C"], g["R"]) == (o["C"], o["R"])
        assert np.array_equal(g["LL"], o["LL"]), "locus %d: per-read log-likelihoods must be bit-identical" % i
        assert np.array_equal(g["mism"], o["mism"])
        assert np.array_equal(g["pair_mavg"], o["pair_mavg"]) and np.array_equal(g["pair_mmin"], o["pair_mmin"])
        rel = np.abs(g["pair_ll"] - o["pair_ll"]) / np.maximum(1.0, np.abs(o["pair_ll"]))
        worst = max(worst, float(rel.max()))
        assert rel.max() < 1e-12, "locus %d: allele-pair log-likelihood sums differ by %g (relative)" % (i, rel.max())
    fo = sorted(os.listdir(or_dir)); assert sorted(os.listdir(out)) == fo and len(fo) == 73
    exact = ["R1_bestguess.txt", "R1_bestguess_G.txt", "summaryStatistics.txt", "histogram_matchesPerRead.txt", "R1_parameters.txt"]
    exact += [f for f in fo if f.startswith(("R1_pileup_", "R1_readIDs_", "R1_columnIncompatibilities_"))]
    bad = [f for f in exact if not filecmp.cmp(os.path.join(or_dir, f), os.path.join(out, f), shallow=False)]
    assert not bad, bad
    # the pair tables print 6 significant digits of sums that may differ in the last bits: compare them field by field
    for f in fo:
        if not f.startswith("R1_PP_"):
            continue
        if filecmp.cmp(os.path.join(or_dir, f), os.path.join(out, f), shallow=False):
            continue
        a = open(os.path.join(or_dir, f)).read().split("\n"); g_ = open(os.path.join(out, f)).read().split("\n")
        assert len(a) == len(g_) and a[0] == g_[0]
        # Same rows; the order may differ only among rows whose LL agree to the printed precision: the reference sorts by the exact doubles, and sums
        # that differ in the last bits (device exp/log, max-shifted products) order near-ties differently
        ra = {l.split("\t")[0]: [float(x) for x in l.split("\t")[1:]] for l in a[1:] if l}; rg = {l.split("\t")[0]: [float(x) for x in l.split("\t")[1:]] for l in g_[1:] if l}
        assert ra.keys() == rg.keys()
        ka = [l.split("\t")[0] for l in a[1:] if l]; kg = [l.split("\t")[0] for l in g_[1:] if l]
        va = np.array([ra[k] for k in ka]); vg = np.array([rg[k] for k in ka])
        assert np.allclose(va, vg, rtol=1e-5, atol=0), f
        lg = np.array([rg[k][1] for k in kg]); assert np.all(np.diff(lg) <= 1e-6 * np.abs(lg[1:])), f     # our file is sorted by LL as well
        moved = [i for i, (x, y) in enumerate(zip(ka, kg)) if x != y]
        for i in moved:
            assert abs(ra[ka[i]][1] - ra[kg[i]][1]) <= 1e-5 * abs(ra[ka[i]][1]), (f, ka[i], kg[i])
    O.close(); T.close(); P.close()


def test_typing_two_rank_split_gives_same_calls(dataset, tmp_path):
    """Pairs extracted by two 'ranks' and gathered, reads split across two ranks for the kernels with a host-side sum standing in for
    the all-reduce: same calls, pair sums within 1e-12 of the single-rank run."""
    d, b, mu, sd = dataset("typing")
    P, T, blobs1 = run_product(d, b, mu, sd, None)
    T.infer(blobs1, mu, sd, None)
    single = [T.locus(i, read_ll=False) for i in range(17)]
    P2, T2, blobs2 = run_product(d, b, mu, sd, None, world=2)
    assert b"".join(blobs1)[64:200] != b"" and sum(len(x) for x in blobs2) > len(blobs1[0]) * 0.9
    # run rank 0 and rank 1 one after the other; the callback records each rank's partial sums instead of reducing them
    partial = [[], []]
    for rank in (0, 1):
        def cb(ctx, ptr, count, stream, rank=rank):
            import torch
            torch.cuda.synchronize()
            partial[rank].append(H.dev_f64_tensor(ptr, count).cpu().numpy().copy())
            return 0
        T2.infer(blobs2, mu, sd, None, rank=rank, world=2, allreduce=cb, keep_read_ll=False)
    for i in range(17):
        n = len(single[i]["pair_ll"])
        tot = partial[0][i] + partial[1][i]
        assert np.allclose(tot[:n], single[i]["pair_ll"], rtol=1e-12, atol=1e-9)
        assert np.array_equal(tot[n:2 * n], single[i]["pair_mavg"]) and np.array_equal(tot[2 * n:], single[i]["pair_mmin"])
    T.close(); P.close(); T2.close(); P2.close()


@pytest.mark.gpu
def test_allele_pair_product_kernel_matches_termwise_and_numpy():
    """the max-shifted allele-pair kernel (one exp per (cluster, read)) against the term-by-term kernel and a numpy statement of
    HLATyper.cpp:2280-2364 / Utilities::logAvg, including reads on which some clusters are > 700 nats below the best one"""
    import ctypes as C
    lib = C.CDLL(H.LIB_PRODUCT); lib.hlala_last_error.restype = C.c_char_p
    rng = np.random.default_rng(3)
    Cn, R = 157, 211
    ll = -rng.gamma(2.0, 20.0, size=(Cn, R)); ll[rng.random((Cn, R)) < 0.05] -= 1500.0; ll[:, 7] -= 900.0; ll[3, 7] += 900.0     # read 7: every cluster but one vanishes
    ll[5] = ll[4]                                                                                                              # two identical clusters (d == 0 branch of logAvg)
    mm = rng.integers(0, 9, size=(Cn, R)).astype(np.int32)
    ll = np.ascontiguousarray(ll); npair = Cn * (Cn + 1) // 2
    out = {}
    for termwise in (1, 0):
        pl = np.zeros(npair); pa = np.zeros(npair); pm = np.zeros(npair); ms = C.c_double(0)
        rc = lib.hlala_typing_pair_probe(0, Cn, R, H.p(ll), H.p(mm), termwise, H.p(pl), H.p(pa), H.p(pm), C.byref(ms))
        assert rc == 0, lib.hlala_last_error().decode()
        out[termwise] = (pl, pa, pm)
    iu = np.triu_indices(Cn)
    want_ll = np.array([np.sum(np.log(0.5) + np.logaddexp(ll[a], ll[b])) for a, b in zip(*iu)])
    want_avg = np.array([np.sum((mm[a] + mm[b]) / 2.0) for a, b in zip(*iu)]); want_min = np.array([np.sum(np.minimum(mm[a], mm[b])) for a, b in zip(*iu)], np.float64)
    for termwise in (1, 0):
        pl, pa, pm = out[termwise]
        assert np.max(np.abs(pl - want_ll) / np.abs(want_ll)) < 1e-12, termwise
        assert np.array_equal(pa, want_avg) and np.array_equal(pm, want_min), termwise
    assert np.max(np.abs(out[0][0] - out[1][0]) / np.abs(out[1][0])) < 1e-12
