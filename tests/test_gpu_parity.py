"""Parity tests proper: the CUDA path, through the C ABI, against the oracle (the compiled reference when oracle/_ref is present,
else the restatement) on seeded inputs; golden fixtures; size-independent properties at larger sizes. Integer/byte/index outputs
are compared bit-exactly; log-likelihoods are compared bit-exactly too (north_star tolerance: 1e-6); mapping qualities within 1e-12."""
import ctypes as C
import os

import numpy as np
import pytest

import harness as H

pytestmark = pytest.mark.gpu

_products = {}


def product(d):
    if d not in _products:
        P = H.Product(d); P.to_gpu(0); _products[d] = P
    return _products[d]


def assert_pairs_equal(got, want):
    n = want["n_cols"]
    assert np.array_equal(got["n_cols"], n)
    for r in range(len(n)):
        for k in ("level", "edge", "gchar", "schar", "from_seed", "mapq"):
            assert np.array_equal(got[k][r, :n[r]], want[k][r, :n[r]]), "read %d: %s" % (r, k)
    assert np.array_equal(got["read_reverse"], want["read_reverse"])
    assert np.allclose(got["pair_mapq"], want["pair_mapq"], rtol=0, atol=1e-12) and np.allclose(got["read_mapq"], want["read_mapq"], rtol=0, atol=1e-12)


@pytest.mark.parametrize("name", ["S", "genes"])
def test_chains_match_oracle(dataset, name):
    d, b, mu, sd = dataset(name)
    want = H.oracle_chains(d, b, 1024); got = product(d).chains(b, 1024)
    for k in ("chain_order", "status", "n_cols", "seed_begin", "seed_end"):
        assert np.array_equal(got[k], want[k]), k
    assert np.array_equal(got["ll"], want["ll"]), "log-likelihoods differ (tolerance 1e-6 allowed by north_star; we require 0)"
    n = want["n_cols"]
    for i in range(len(n)):
        for k in ("level", "edge", "gchar", "schar", "from_seed"):
            assert np.array_equal(got[k][i, :n[i]], want[k][i, :n[i]]), "slot %d: %s" % (i, k)


@pytest.mark.parametrize("name", ["S", "genes"])
def test_pairs_match_oracle(dataset, name):
    d, b, mu, sd = dataset(name)
    want = H.oracle_pairs(d, b, mu, sd, 1024); got = product(d).pairs(b, mu, sd, 1024)
    assert_pairs_equal(got, want)
    # per-level coverage (processBAM.cpp:2411-2426) recomputed from the oracle's alignments
    cov = np.zeros_like(got["bases_per_level"])
    for r in range(len(want["n_cols"])):
        n = want["n_cols"][r]; lv = want["level"][r, :n]; g = want["gchar"][r, :n]; m = (lv != -1) & (g != ord("_"))
        np.add.at(cov, lv[m], 1)
    assert np.array_equal(cov, got["bases_per_level"])


def test_golden_fixture(dataset):
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "alignment_small.npz"))
    d, b, mu, sd = dataset("small")
    got = product(d).pairs(b, mu, sd, 512)
    pack = lambda cols, n: np.concatenate([cols[r, :n[r]] for r in range(len(n))])
    assert np.array_equal(got["n_cols"], gold["n_cols"])
    for k in ("level", "edge", "gchar", "schar", "from_seed", "mapq"):
        assert np.array_equal(pack(got[k], got["n_cols"]), gold[k]), k
    assert np.allclose(got["pair_mapq"], gold["pair_mapq"], rtol=0, atol=1e-12)
    ch = product(d).chains(b, 512)
    assert np.array_equal(ch["ll"], gold["chain_ll"]) and np.array_equal(pack(ch["edge"], ch["n_cols"]), gold["chain_edge"])


def test_cascade_regression_fixture(tmp_path):
    """19 pairs found by tools/scale_parity.py at 1 M pairs where a revisited cell (after a gap-path jump) decided a tie in the
    extension DP; expected outputs come from the compiled reference (tests/golden/make_cascade_regress.py)."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_cascade_regress import PRG_KW
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cascade_regress_pairs.npz"))
    d = str(tmp_path / "prg"); H.synth_prg(d, **PRG_KW)
    b = {k[3:]: gold[k] for k in gold.files if k.startswith("in_")}
    P = H.Product(d); P.to_gpu(0)
    got = P.pairs(b, 100.0, 10.0, 640, want_levels=False)
    pack = lambda cols, n: np.concatenate([cols[r, :n[r]] for r in range(len(n))])
    assert np.array_equal(got["n_cols"], gold["n_cols"])
    for k in ("level", "edge", "gchar", "schar", "from_seed", "mapq"):
        assert np.array_equal(pack(got[k], got["n_cols"]), gold[k]), k
    assert np.allclose(got["pair_mapq"], gold["pair_mapq"], rtol=0, atol=1e-12) and np.allclose(got["read_mapq"], gold["read_mapq"], rtol=0, atol=1e-12)
    P.close()


def test_waves_and_scalar_dp_give_identical_results(dataset, monkeypatch):
    d, b, mu, sd = dataset("S")
    base = product(d).pairs(b, mu, sd, 1024)
    monkeypatch.setenv("HLALA_WAVE_BYTES", "2000000")      # many waves, alternating between three lanes (streams)
    waves = product(d).pairs(b, mu, sd, 1024)
    monkeypatch.setenv("HLALA_LANES", "1")                 # many waves, one after the other on one lane
    one_lane = product(d).pairs(b, mu, sd, 1024)
    monkeypatch.delenv("HLALA_WAVE_BYTES"); monkeypatch.delenv("HLALA_LANES")
    monkeypatch.setenv("HLALA_SCALAR_DP", "1")             # every extension through the scalar DP kernel
    scalar = product(d).pairs(b, mu, sd, 1024)
    for other in (waves, one_lane, scalar):
        for k in base:
            if base[k] is not None:
                assert np.array_equal(base[k], other[k]), k


def test_capacity_error_is_reported(dataset):
    d, b, mu, sd = dataset("small")
    with pytest.raises(RuntimeError, match="capacity"):
        product(d).pairs(b, mu, sd, 64)         # 100-base reads cannot fit 64 columns


def test_session_digest_idempotence_and_shard_additivity(dataset, tmp_path):
    """Size-independent properties on a batch too large for the oracle to be worth running."""
    import sys
    sys.path.insert(0, H.PKG)
    import hlala_dist
    d, _, mu, sd = dataset("S")
    b = H.synth_reads(d, str(tmp_path / "big.bin"), pairs=20000, len=100, clip_frac=0.15, seed=99)
    P = product(d)
    full = P.pairs(b, mu, sd, 512)
    again = P.pairs(b, mu, sd, 512)
    for k in full:
        if full[k] is not None:
            assert np.array_equal(full[k], again[k]), "run-to-run difference in " + k
    halves = [P.pairs(hlala_dist.shard_batch(b, r, 2), mu, sd, 512) for r in range(2)]
    assert np.array_equal(full["bases_per_level"], halves[0]["bases_per_level"] + halves[1]["bases_per_level"])
    assert np.array_equal(full["pair_mapq"], np.concatenate([h["pair_mapq"] for h in halves]))
    assert np.array_equal(full["n_cols"], np.concatenate([h["n_cols"] for h in halves]))
    assert (full["n_cols"] >= 100).all() and ((full["pair_mapq"] >= 0) & (full["pair_mapq"] <= 1)).all()
    # every read base appears exactly once in its alignment
    nb = np.array([(full["schar"][r, :full["n_cols"][r]] != ord("_")).sum() for r in range(0, len(full["n_cols"]), 97)])
    assert (nb == 100).all()
    # session API digest agrees with the fetched results
    L = P.lib
    sb = H.make_batch_struct(b); sess = C.c_void_p()
    P._chk(L.hlala_session_create(P.g, C.byref(sb), C.c_int32(512), C.byref(sess)))
    P._chk(L.hlala_session_run(sess, C.c_double(mu), C.c_double(sd), C.c_uint64(0), None))
    dig = (C.c_int64 * 4)(); sll = C.c_double(0)
    P._chk(L.hlala_session_digest(sess, dig, C.byref(sll)))
    assert dig[0] == int(full["n_cols"].sum()) and dig[2] == int((full["pair_mapq"] < 1).sum()) and dig[3] == 0
    edges = sum(int((full["edge"][r, :full["n_cols"][r]].astype(np.int64) + 1).sum()) for r in range(len(full["n_cols"])))
    assert dig[1] == edges
    assert abs(sll.value - full["pair_ll"].sum()) < 1e-6 * abs(sll.value)
    L.hlala_session_free(sess)


def test_pairs_match_oracle_250bp_wide_insert(dataset):
    """BASELINE.json config sweep: 2x250 bp reads, 30 % soft-clipped, more indels, insert-size model N(250, 35)"""
    d, b, mu, sd = dataset("L250")
    want = H.oracle_pairs(d, b, mu, sd, 1024); got = product(d).pairs(b, mu, sd, 1024)
    assert_pairs_equal(got, want)


def test_empty_and_single_pair_batches(dataset):
    """no pairs at all: nothing is aligned, nothing is counted, no error; a batch of one pair"""
    d, b, mu, sd = dataset("small")
    P = product(d)
    empty = {k: (np.zeros(1, b[k].dtype) if k in ("read_off", "chain_off", "cigar_off") else np.zeros(1, b[k].dtype)[:0].copy()) for k in H.BATCH_KEYS}
    for k in ("bases", "quals", "chain_contig", "chain_pos", "chain_flag", "chain_as", "cigar"):
        empty[k] = np.zeros(1, b[k].dtype)          # non-null pointers, zero logical length
    got = P.pairs(empty, mu, sd, 256)
    assert len(got["pair_mapq"]) == 0 and int(got["bases_per_level"].sum()) == 0
    # a single pair on its own gives what it gives inside the batch
    import sys
    sys.path.insert(0, H.PKG)
    import hlala_dist
    full = P.pairs(b, mu, sd, 512)
    npairs = (len(b["read_off"]) - 1) // 2
    one = P.pairs(hlala_dist.shard_batch(b, npairs - 1, npairs), mu, sd, 512)
    assert np.array_equal(one["n_cols"], full["n_cols"][-2:]) and np.array_equal(one["pair_mapq"], full["pair_mapq"][-1:])
    for k in ("level", "edge", "schar", "mapq"):
        assert np.array_equal(one[k], full[k][-2:]), k


def _with_many_chains(b, contig_len, n_extra, pairs, seed=1):
    """first `pairs` pairs: every read gets n_extra extra secondary records (the primary's CIGAR at other places of its contig, same strand,
    lower score): more kept chains per read and more combinations per pair than the pair kernel's small tier holds"""
    rng = np.random.RandomState(seed); nr = len(b["read_off"]) - 1
    out = {k: [] for k in ("chain_contig", "chain_pos", "chain_flag", "chain_as")}; cig = []; cig_off = [0]; chain_off = [0]
    for r in range(nr):
        c0, c1 = int(b["chain_off"][r]), int(b["chain_off"][r + 1])
        recs = [(int(b["chain_contig"][c]), int(b["chain_pos"][c]), int(b["chain_flag"][c]), int(b["chain_as"][c]), list(b["cigar"][b["cigar_off"][c]:b["cigar_off"][c + 1]])) for c in range(c0, c1)]
        if recs and r < 2 * pairs:
            prim = [x for x in recs if not (x[2] & 0x100)][0]; ctg, pos, flag, as_, cg = prim
            reflen = sum(int(x) >> 4 for x in cg if (int(x) & 15) in (0, 2, 3, 7, 8))
            for k in range(n_extra):
                p2 = int(rng.randint(0, contig_len[ctg] - reflen - 1))
                recs.append((ctg, p2, flag | 0x100, as_ - 1 - k, cg))
        for ctg, pos, flag, as_, cg in recs:
            out["chain_contig"].append(ctg); out["chain_pos"].append(pos); out["chain_flag"].append(flag); out["chain_as"].append(as_)
            cig += [int(x) for x in cg]; cig_off.append(len(cig))
        chain_off.append(len(out["chain_contig"]))
    nb = dict(b)
    nb["chain_off"] = np.array(chain_off, b["chain_off"].dtype); nb["cigar_off"] = np.array(cig_off, b["cigar_off"].dtype); nb["cigar"] = np.array(cig, b["cigar"].dtype)
    for k in out:
        nb[k] = np.array(out[k], b[k].dtype)
    return nb


def test_pairs_beyond_the_small_pair_tier(dataset):
    """reads with 44 kept chains (> 32) and pairs with > 512 combinations: served by the pair kernel's large tier, equal to the oracle (which, like the reference, has no limit)"""
    d, b, mu, sd = dataset("small")
    P = product(d); contig_len = np.diff(P.array("contig_off"))
    nb = _with_many_chains(b, contig_len, 43, 6)
    sub = {k: v for k, v in nb.items()}
    want = H.oracle_pairs(d, nb, mu, sd, 1024); got = P.pairs(nb, mu, sd, 1024)
    assert_pairs_equal(got, want)
    # beyond the large tier (> 64 kept chains) the pair is reported, not silently dropped
    nb2 = _with_many_chains(b, contig_len, 70, 1)
    with pytest.raises(RuntimeError, match="capacity|invariant"):
        P.pairs(nb2, mu, sd, 1024)
