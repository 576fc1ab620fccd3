"""k-mer seeding (GraphAndEdgeIndex::Index / queryIndex / findChains).
CPU: the oracle restatement and the product's host index builder against the compiled reference (when oracle/_ref is present) and
against the committed golden fixture (generated from the compiled reference by tests/golden/make_golden_kmer.py).
GPU: hlala_seed_kmers through the C ABI against the oracle, bit-exact including the order of the chains."""
import hashlib
import os

import numpy as np
import pytest

import harness as H
from conftest import DATASETS

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kmer_small.npz")
KEYS = ("chain_off", "begin", "end", "edge_off", "edges")


def index_digest(ix):
    return np.frombuffer(hashlib.sha1(ix["kmers"].tobytes() + ix["pos_off"].tobytes() + ix["edge_off"].tobytes() + ix["edges"].tobytes()).digest(), np.uint8)


def same_index(a, b):
    return all(a[k].shape == b[k].shape and (a[k] == b[k]).all() for k in ("kmers", "pos_off", "edge_off", "edges"))


@pytest.fixture(scope="module")
def small(dataset):
    d, _b, _mu, _sd = dataset("small")
    return d, H.Oracle(d).graph()


@pytest.mark.parametrize("k", [25, 11])
def test_oracle_matches_golden(small, k):
    d, G = small
    gold = np.load(GOLD)
    O = H.OracleKmer(G, k)
    ix = O.kmer_dump()
    assert np.array_equal(index_digest(ix), gold["k%d_index_sha1" % k]), "oracle index differs from the compiled reference's"
    ch = O.find_chains(gold["read_off"], gold["bases"])
    for key in KEYS:
        assert np.array_equal(ch[key], gold["k%d_%s" % (k, key)]), key
    O.close()


@pytest.mark.parametrize("k", [25, 11, 4])
def test_product_index_matches_oracle(small, k):
    """host side of the product (no GPU needed): enumeration + ordering of hla-la_b200/host/kmer_index.cpp"""
    d, G = small
    P = H.Product(d); P.kmer_index(k)
    O = H.OracleKmer(G, k)
    a = P.kmer_dump(); b = O.kmer_dump()
    assert same_index(a, b)
    if k in (25, 11):
        gold = np.load(GOLD)
        assert np.array_equal(index_digest(a), gold["k%d_index_sha1" % k])
        assert list(gold["k%d_index_dims" % k]) == [len(a["kmers"]), len(a["edge_off"]) - 1, len(a["edges"])]
    O.close(); P.close()


def test_index_argument_errors(small):
    d, _G = small
    P = H.Product(d)
    with pytest.raises(RuntimeError):
        P.kmer_index(1)
    with pytest.raises(RuntimeError):
        P.kmer_dump()          # no index yet
    P.kmer_index(8)
    off, bases = np.zeros(2, np.int64), np.zeros(1, np.uint8)
    with pytest.raises(RuntimeError, match="no CPU fallback|not on a GPU"):
        P.find_chains(off, bases)      # the graph is not on a GPU: the product refuses instead of falling back
    P.close()


@pytest.mark.skipif(not H.have_ref(), reason="compiled reference (oracle/_ref) not built on this box")
def test_oracle_and_product_match_compiled_reference():
    """runs in its own process: one compiled-reference graph per process"""
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "kmer_ref_compare.py")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout[-3000:]


@pytest.mark.skipif(not H.have_ref(), reason="compiled reference (oracle/_ref) not built on this box")
@pytest.mark.parametrize("seed", [1, 7, 19])
def test_fuzz_random_graphs_against_compiled_reference(seed):
    """random levelled graphs with gap bubbles and shuffled node / edge order (tests/kmer_fuzz.py, own process per graph)"""
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "kmer_fuzz.py"), str(seed), "3", "6"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0 and r.stdout.count(" ok") == 2, r.stdout[-3000:]


# ------------------------------------------------------------------------------------------------------------------ GPU
_gpu = {}


def gpu_product(d):
    if d not in _gpu:
        P = H.Product(d); P.to_gpu(0); _gpu[d] = P
    return _gpu[d]


def per_read(ch, r):
    c0, c1 = ch["chain_off"][r], ch["chain_off"][r + 1]
    return (ch["begin"][c0:c1], ch["end"][c0:c1], np.diff(ch["edge_off"][c0:c1 + 1]), ch["edges"][ch["edge_off"][c0]:ch["edge_off"][c1]])


def assert_reads_equal(got, want, allow_capacity=False):
    """bit-exact, order included; a read the kernel flagged HLALA_E_CAPACITY (-4) must have no chains"""
    nr = len(want["chain_off"]) - 1
    assert len(got["chain_off"]) - 1 == nr
    if not allow_capacity:
        assert got["n_failed"] == 0 and (got["status"] == 0).all()
        for key in KEYS:
            assert np.array_equal(got[key], want[key]), key
        return 0
    assert set(np.unique(got["status"])) <= {0, -4}
    for r in range(nr):
        if got["status"][r] != 0:
            assert got["chain_off"][r + 1] == got["chain_off"][r]
            continue
        for a, b in zip(per_read(got, r), per_read(want, r)):
            assert np.array_equal(a, b), "read %d" % r
    assert got["n_failed"] == int((got["status"] != 0).sum())
    return got["n_failed"]


@pytest.mark.gpu
@pytest.mark.parametrize("name,k,n,length", [("small", 25, 600, 150), ("small", 11, 400, 100), ("S", 25, 1500, 150), ("typing", 25, 600, 250), ("typing", 14, 300, 150), ("small", 5, 60, 40)])
def test_gpu_chains_match_oracle(dataset, name, k, n, length):
    d, _b, _mu, _sd = dataset(name)
    G = H.Oracle(d).graph()
    O = H.OracleKmer(G, k)
    off, bases = H.walk_reads(G, n, length, 1000 + k)
    want = O.find_chains(off, bases)
    P = gpu_product(d); P.kmer_index(k)
    got = P.find_chains(off, bases)
    assert_reads_equal(got, want)
    O.close()


@pytest.mark.gpu
def test_gpu_chains_second_tier_and_capacity(dataset):
    """a short k in a 14-node-wide gene region keeps hundreds of chains alive. Reads beyond the first tier's 128 running chains are re-run by the second
    tier (1024 running chains) and are exact; what exceeds that as well is flagged HLALA_E_CAPACITY and carries no chains, never a silent truncation"""
    import os
    d, _b, _mu, _sd = dataset("genes")
    G = H.Oracle(d).graph()
    P = gpu_product(d)
    for k, seed in ((12, 1012), (9, 77)):
        O = H.OracleKmer(G, k)
        off, bases = H.walk_reads(G, 300, 150, seed)
        want = O.find_chains(off, bases)
        P.kmer_index(k)
        os.environ["HLALA_SEED_NO_BIG_TIER"] = "1"
        try:
            one = P.find_chains(off, bases)
        finally:
            os.environ.pop("HLALA_SEED_NO_BIG_TIER")
        n_failed_one_tier = assert_reads_equal(one, want, allow_capacity=True)
        got = P.find_chains(off, bases)
        n_failed = assert_reads_equal(got, want, allow_capacity=True)
        assert one["n_second_tier"] == 0 and got["n_second_tier"] == n_failed_one_tier
        assert 0 < n_failed_one_tier < (300 if k == 9 else 100) and n_failed < n_failed_one_tier
        if k == 12:
            assert n_failed == 0
        O.close()


@pytest.mark.gpu
def test_gpu_chains_match_golden(dataset):
    d, _b, _mu, _sd = dataset("small")
    gold = np.load(GOLD)
    P = gpu_product(d)
    for k in (25, 11):
        P.kmer_index(k)
        got = P.find_chains(gold["read_off"], gold["bases"])
        for key in KEYS:
            assert np.array_equal(got[key], gold["k%d_%s" % (k, key)]), "%s (k=%d)" % (key, k)


@pytest.mark.gpu
def test_gpu_chains_edge_cases(dataset):
    """empty batch, reads shorter than k, a read of exactly k bases, reads without any indexed k-mer; and the size-independent property that
    every chain spells its stretch of the read along connected edges"""
    d, _b, _mu, _sd = dataset("small")
    G = H.Oracle(d).graph(); k = 25
    P = gpu_product(d); P.kmer_index(k)
    got = P.find_chains(np.zeros(1, np.int64), np.zeros(1, np.uint8))
    assert len(got["begin"]) == 0 and list(got["chain_off"]) == [0]
    off, bases = H.walk_reads(G, 200, 150, 5, err=0.0, random_frac=0.0, short_frac=0.0)
    # truncate a few reads to k-1, k and k+1 bases and make one read all 'N'
    lens = np.diff(off).copy(); lens[0] = k - 1; lens[1] = k; lens[2] = k + 1
    seqs = [bases[off[i]:off[i] + lens[i]].copy() for i in range(len(lens))]; seqs[3][:] = ord("N")
    off2 = np.zeros(len(lens) + 1, np.int64); off2[1:] = np.cumsum([len(s) for s in seqs]); bases2 = np.ascontiguousarray(np.concatenate(seqs))
    O = H.OracleKmer(G, k); want = O.find_chains(off2, bases2); got = P.find_chains(off2, bases2)
    for key in KEYS:
        assert np.array_equal(got[key], want[key]), key
    assert got["chain_off"][1] == 0 and got["chain_off"][4] == got["chain_off"][3]
    ef, et, em = G["edge_from"], G["edge_to"], G["edge_emis"]
    for r in range(len(lens)):
        for c in range(got["chain_off"][r], got["chain_off"][r + 1]):
            e = got["edges"][got["edge_off"][c]:got["edge_off"][c + 1]]
            assert (et[e[:-1]] == ef[e[1:]]).all()
            spelled = em[e][em[e] != ord("_")]
            assert np.array_equal(spelled, seqs[r][got["begin"][c]:got["end"][c] + 1])
    O.close()
