"""Pins the CPU oracle against the UNMODIFIED reference translation units (oracle/_ref), when that library was built."""
import numpy as np
import pytest

import harness as H

pytestmark = pytest.mark.skipif(not H.have_ref(), reason="oracle/_ref/libhlala_ref.so not built (reference tree absent)")

_ref = {}


def ref_for(d):
    if not _ref:
        _ref[d] = H.quiet(H.Ref, d)
    if d not in _ref:
        pytest.skip("one compiled-reference graph per process")
    return _ref[d]


def test_graph_gap_paths_stretches(dataset):
    d, b, mu, sd = dataset("S")
    rg = ref_for(d).graph(); og = H.Oracle(d).graph()
    for k in og:
        assert np.array_equal(np.asarray(rg[k]), np.asarray(og[k])), k
    assert len(og["path_off"]) > 100 and og["gap_stretch"].sum() > 0


def test_chains_bit_equal(dataset):
    d, b, mu, sd = dataset("S")
    rc = H.quiet(ref_for(d).chains, b); oc = H.Oracle(d).chains(b)
    for k in ("chain_order", "status", "n_cols", "seed_begin", "seed_end", "ll", "level", "edge", "gchar", "schar", "from_seed"):
        assert np.array_equal(rc[k], oc[k]), k
    assert ((rc["status"] == 0) & ((rc["seed_begin"] != 0) | (rc["seed_end"] != 99))).sum() > 100, "dataset must exercise the extension DP"


def test_pairs_bit_equal(dataset):
    d, b, mu, sd = dataset("S")
    rp = H.quiet(ref_for(d).pairs, b, mu, sd); op = H.Oracle(d).pairs(b, mu, sd)
    for k in ("pair_mapq", "read_mapq", "read_reverse", "n_cols", "level", "edge", "gchar", "schar", "from_seed", "mapq"):
        assert np.array_equal(rp[k], op[k]), k


@pytest.mark.parametrize("seed", [2, 11])
def test_fuzz_random_graphs(seed):
    """random levelled graphs with dense gap structure, contigs = random walks, soft-clipped reads (tests/align_fuzz.py, own process per graph):
    oracle restatement and the host build of the GPU's extension DP against the compiled reference, bit for bit"""
    import os
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "align_fuzz.py"), str(seed)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0 and "pairs ok" in r.stdout, r.stdout[-3000:]
