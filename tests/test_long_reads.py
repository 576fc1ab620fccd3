"""Long-read mode (HLA-LA.pl --longReads; processBAM::alignOneLongRead, mapper/processBAM.cpp:3618-3838; assignMappingQualities_unpaired :3900-4059;
indel rates 0.075, extensionAligner.cpp:58-64). CPU half: the restatement against the compiled reference and against the golden fixture;
GPU half: hlala_align_long_reads through the C ABI against the strongest oracle, the golden fixture, and per-level coverage."""
import hashlib
import os

import numpy as np
import pytest

import harness as H

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "long_reads_small.npz")
COLS = ("level", "edge", "gchar", "schar", "from_seed", "mapq")


def pack(cols, n):
    return np.concatenate([cols[r, :n[r]] for r in range(len(n))])


def assert_gold(got, ll_key, mq_key, b):
    gold = np.load(GOLD)
    assert np.array_equal(gold["input_sha1"], np.frombuffer(hashlib.sha1(b"".join(b[k].tobytes() for k in H.BATCH_KEYS)).digest(), np.uint8)), "generator output changed: regenerate the fixture"
    assert np.array_equal(got["n_cols"], gold["n_cols"]) and np.array_equal(got["read_reverse"], gold["read_reverse"])
    for k in COLS:
        assert np.array_equal(pack(got[k], got["n_cols"]), gold[k]), k
    assert np.array_equal(got[ll_key], gold["read_ll"]), "log-likelihoods differ"
    assert np.allclose(got[mq_key], gold["read_mapq"], rtol=0, atol=1e-12)


@pytest.mark.skipif(not H.have_ref(), reason="oracle/_ref/libhlala_ref.so not built (reference tree absent)")
def test_restatement_equals_compiled_reference(dataset):
    d, b, mu, sd = dataset("long")
    want = H._ref_subprocess(d, b, 0.0, 1.0, 8192, "long_reads"); got = H.Oracle(d).long_reads(b, 8192)
    for k in ("n_cols", "read_reverse", "read_mapq", "read_ll") + COLS:
        assert np.array_equal(got[k], want[k]), k
    assert (want["read_mapq"] < 1).sum() >= 3 and want["n_cols"].max() > 4000, "dataset must hold ambiguous reads and long alignments"


def test_restatement_golden(dataset):
    d, b, mu, sd = dataset("long_small")
    assert_gold(H.Oracle(d).long_reads(b, 4096), "read_ll", "read_mapq", b)


_products = {}


def product(d):
    if d not in _products:
        P = H.Product(d); P.to_gpu(0); _products[d] = P
    return _products[d]


@pytest.mark.gpu
def test_gpu_golden(dataset):
    d, b, mu, sd = dataset("long_small")
    assert_gold(product(d).long_reads(b, 4096), "pair_ll", "pair_mapq", b)


@pytest.mark.gpu
@pytest.mark.parametrize("name,cap", [("long", 8192), ("long8k", 10240)])
def test_gpu_long_reads_match_oracle(dataset, name, cap):
    d, b, mu, sd = dataset(name)
    want = H.oracle_long_reads(d, b, cap); got = product(d).long_reads(b, cap)
    assert name != "long8k" or want["n_cols"].max() > 8200
    n = want["n_cols"]
    assert np.array_equal(got["n_cols"], n) and np.array_equal(got["read_reverse"], want["read_reverse"])
    for r in range(len(n)):
        for k in COLS:
            assert np.array_equal(got[k][r, :n[r]], want[k][r, :n[r]]), "read %d: %s" % (r, k)
    assert np.array_equal(got["pair_ll"], want["read_ll"]), "log-likelihoods differ (tolerance 1e-6 allowed by north_star; we require 0)"
    assert np.allclose(got["pair_mapq"], want["read_mapq"], rtol=0, atol=1e-12) and np.array_equal(got["pair_mapq"], got["read_mapq"])
    cov = np.zeros_like(got["bases_per_level"])     # processBAM.cpp:2289-2296
    for r in range(len(n)):
        lv = want["level"][r, :n[r]]; g = want["gchar"][r, :n[r]]; m = (lv != -1) & (g != ord("_"))
        np.add.at(cov, lv[m], 1)
    assert np.array_equal(cov, got["bases_per_level"])


@pytest.mark.gpu
def test_gpu_long_reads_edge_cases(dataset):
    """an empty batch; a batch of one read; max_columns too small for the read is a reported capacity error, not a silent truncation"""
    d, b, mu, sd = dataset("long_small")
    P = product(d)
    one = {k: b[k] for k in H.BATCH_KEYS}
    nch = int(b["chain_off"][1]); ncg = int(b["cigar_off"][nch])
    one.update(read_off=b["read_off"][:2].copy(), bases=b["bases"][:b["read_off"][1]].copy(), quals=b["quals"][:b["read_off"][1]].copy(), chain_off=b["chain_off"][:2].copy(),
               chain_contig=b["chain_contig"][:nch].copy(), chain_pos=b["chain_pos"][:nch].copy(), chain_flag=b["chain_flag"][:nch].copy(), chain_as=b["chain_as"][:nch].copy(),
               cigar_off=b["cigar_off"][:nch + 1].copy(), cigar=b["cigar"][:ncg].copy())
    got = P.long_reads(one, 4096); full = P.long_reads(b, 4096)
    m = int(full["n_cols"][0])
    assert got["n_cols"][0] == m and np.array_equal(got["edge"][0, :m], full["edge"][0, :m]) and np.array_equal(got["mapq"][0, :m], full["mapq"][0, :m]) and got["pair_ll"][0] == full["pair_ll"][0]
    empty = {k: b[k][:0].copy() for k in H.BATCH_KEYS}
    empty.update(read_off=np.zeros(1, np.int64), chain_off=np.zeros(1, np.int32), cigar_off=np.zeros(1, np.int32))
    assert len(P.long_reads(empty, 4096)["n_cols"]) == 0
    with pytest.raises(RuntimeError, match="capacity"):
        P.long_reads(one, 512)
