"""The CPU oracle (oracle/hlala_oracle.cpp) against golden vectors generated from the compiled reference (tests/golden/)."""
import hashlib
import os

import numpy as np

import harness as H

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "alignment_small.npz"))


def pack(cols, n):
    return np.concatenate([cols[r, :n[r]] for r in range(len(n))])


def test_generator_is_deterministic(dataset):
    d, b, mu, sd = dataset("small")
    sha = np.frombuffer(hashlib.sha1(b"".join(b[k].tobytes() for k in H.BATCH_KEYS)).digest(), np.uint8)
    assert np.array_equal(sha, GOLD["input_sha1"]), "hlala-synth no longer reproduces the inputs the golden vectors were made from"


def test_oracle_graph_matches_golden_digest(dataset):
    d, b, mu, sd = dataset("small")
    g = H.Oracle(d).graph()
    sha = np.frombuffer(hashlib.sha1(g["edge_from"].tobytes() + g["edge_to"].tobytes() + g["edge_emis"].tobytes() + g["path_edges"].tobytes() + g["gap_stretch"].tobytes()).digest(), np.uint8)
    assert np.array_equal(sha, GOLD["graph_sha1"])


def test_oracle_chains_match_golden(dataset):
    d, b, mu, sd = dataset("small")
    ch = H.Oracle(d).chains(b, 512)
    for k in ("chain_order", "status", "n_cols", "seed_begin", "seed_end"):
        assert np.array_equal(ch[k], GOLD["chain_" + k if k != "chain_order" else k]), k
    assert np.array_equal(ch["ll"], GOLD["chain_ll"]), "per-chain log-likelihoods must be bit-identical (tolerance stated by north_star: 1e-6; achieved: 0)"
    for k in ("level", "edge", "gchar", "schar", "from_seed"):
        assert np.array_equal(pack(ch[k], ch["n_cols"]), GOLD["chain_" + k]), k


def test_oracle_pairs_match_golden(dataset):
    d, b, mu, sd = dataset("small")
    pr = H.Oracle(d).pairs(b, mu, sd, 512)
    assert np.array_equal(pr["n_cols"], GOLD["n_cols"])
    for k in ("level", "edge", "gchar", "schar", "from_seed", "mapq"):
        assert np.array_equal(pack(pr[k], pr["n_cols"]), GOLD[k]), k
    assert np.array_equal(pr["pair_mapq"], GOLD["pair_mapq"]) and np.array_equal(pr["read_mapq"], GOLD["read_mapq"])
    assert np.array_equal(pr["read_reverse"], GOLD["read_reverse"])
    assert (GOLD["pair_mapq"] < 1).sum() > 0, "the fixture must exercise the multi-combination mapQ path"


def test_oracle_matches_cascade_regression_fixture(tmp_path):
    """The restatement on the 19 tie-after-gap-jump pairs of tests/golden/cascade_regress_pairs.npz (expected = compiled reference)."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_cascade_regress import PRG_KW
    gold = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cascade_regress_pairs.npz"))
    d = str(tmp_path / "prg"); H.synth_prg(d, **PRG_KW)
    b = {k[3:]: gold[k] for k in gold.files if k.startswith("in_")}
    pr = H.Oracle(d).pairs(b, 100.0, 10.0, 640)
    assert np.array_equal(pr["n_cols"], gold["n_cols"])
    for k in ("level", "edge", "gchar", "schar", "from_seed", "mapq"):
        assert np.array_equal(pack(pr[k], pr["n_cols"]), gold[k]), k
    assert np.array_equal(pr["pair_mapq"], gold["pair_mapq"]) and np.array_equal(pr["read_mapq"], gold["read_mapq"])
