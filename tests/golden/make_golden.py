#!/usr/bin/env python
"""Generates tests/golden/alignment_small.npz from the COMPILED REFERENCE (oracle/_ref/libhlala_ref.so, i.e. the unmodified
reference translation units): per-chain and per-pair outputs for the deterministic 'small' dataset of tests/conftest.py.
Run in the build container (needs /root/reference to have been compiled by oracle/Makefile.ref):  python tests/golden/make_golden.py"""
import hashlib
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import harness as H  # noqa: E402
from conftest import DATASETS  # noqa: E402


def pack(cols, n):
    """ragged [rows, cap] arrays -> flat concatenation of the first n[r] entries of each row"""
    return np.concatenate([cols[r, :n[r]] for r in range(len(n))]) if len(n) else cols[:0, 0]


def main():
    prg_kw, rd_kw, mu, sd = DATASETS["small"]
    d = tempfile.mkdtemp(prefix="golden_")
    H.synth_prg(d, **prg_kw)
    b = H.synth_reads(d, os.path.join(d, "seeds.bin"), **rd_kw)
    R = H.quiet(H.Ref, d)
    ch = H.quiet(R.chains, b, 512)
    pr = H.quiet(R.pairs, b, mu, sd, 512)
    g = R.graph()
    out = {"input_sha1": np.frombuffer(hashlib.sha1(b"".join(b[k].tobytes() for k in H.BATCH_KEYS)).digest(), np.uint8),
           "graph_sha1": np.frombuffer(hashlib.sha1(g["edge_from"].tobytes() + g["edge_to"].tobytes() + g["edge_emis"].tobytes() + g["path_edges"].tobytes() + g["gap_stretch"].tobytes()).digest(), np.uint8),
           "chain_order": ch["chain_order"], "chain_status": ch["status"], "chain_n_cols": ch["n_cols"], "chain_seed_begin": ch["seed_begin"], "chain_seed_end": ch["seed_end"], "chain_ll": ch["ll"],
           "pair_mapq": pr["pair_mapq"], "read_mapq": pr["read_mapq"], "read_reverse": pr["read_reverse"], "n_cols": pr["n_cols"]}
    for k in ("level", "edge", "gchar", "schar", "from_seed"):
        out["chain_" + k] = pack(ch[k], ch["n_cols"])
    for k in ("level", "edge", "gchar", "schar", "from_seed", "mapq"):
        out[k] = pack(pr[k], pr["n_cols"])
    path = os.path.join(HERE, "alignment_small.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(ch["status"]), "chains,", len(pr["pair_mapq"]), "pairs")


if __name__ == "__main__":
    main()
