#!/usr/bin/env python
"""Generates tests/golden/long_reads_small.npz from the COMPILED REFERENCE (oracle/_ref/libhlala_ref.so: the unmodified processBAM::alignOneLongRead,
assignMappingQualities_unpaired and extensionAligner::scoreOneAlignment in long-read mode) for the deterministic 'long_small' dataset of tests/conftest.py.
Run in the build container:  python tests/golden/make_golden_long.py"""
import hashlib
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import harness as H  # noqa: E402
from conftest import DATASETS  # noqa: E402
from make_golden import pack  # noqa: E402


def main():
    prg_kw, rd_kw, mu, sd = DATASETS["long_small"]
    d = tempfile.mkdtemp(prefix="golden_long_")
    H.synth_prg(d, **prg_kw)
    b = H.synth_reads(d, os.path.join(d, "seeds.bin"), **rd_kw)
    R = H.quiet(H.Ref, d)
    r = H.quiet(R.long_reads, b, 4096)
    out = {"input_sha1": np.frombuffer(hashlib.sha1(b"".join(b[k].tobytes() for k in H.BATCH_KEYS)).digest(), np.uint8),
           "read_mapq": r["read_mapq"], "read_ll": r["read_ll"], "read_reverse": r["read_reverse"], "n_cols": r["n_cols"]}
    for k in ("level", "edge", "gchar", "schar", "from_seed", "mapq"):
        out[k] = pack(r[k], r["n_cols"])
    path = os.path.join(HERE, "long_reads_small.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "reads", len(r["n_cols"]), "columns", int(r["n_cols"].sum()), "reads with mapQ < 1:", int((r["read_mapq"] < 1).sum()))


if __name__ == "__main__":
    main()
