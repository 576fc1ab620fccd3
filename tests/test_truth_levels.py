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
