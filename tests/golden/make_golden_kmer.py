#!/usr/bin/env python
"""Generates tests/golden/kmer_small.npz from the COMPILED REFERENCE (oracle/_ref/libhlala_ref.so = the unmodified
Graph/GraphAndEdgeIndex.cpp): digest of the whole k-mer index and the findChains results for seeded walk reads on the deterministic
'small' PRG of tests/conftest.py, for k = 25 (the reference's own choice, HLA-LA.cpp:230) and k = 11.
Run in the build container:  python tests/golden/make_golden_kmer.py"""
import hashlib
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import harness as H  # noqa: E402
from conftest import DATASETS  # noqa: E402

KS = (25, 11)
N_READS, READ_LEN, SEED = 400, 120, 17


def index_digest(ix):
    return np.frombuffer(hashlib.sha1(ix["kmers"].tobytes() + ix["pos_off"].tobytes() + ix["edge_off"].tobytes() + ix["edges"].tobytes()).digest(), np.uint8)


def main():
    prg_kw = DATASETS["small"][0]
    d = tempfile.mkdtemp(prefix="golden_kmer_")
    H.synth_prg(d, **prg_kw)
    R = H.quiet(H.Ref, d)
    G = R.graph()
    off, bases = H.walk_reads(G, N_READS, READ_LEN, SEED)
    out = {"read_off": off, "bases": bases}
    for k in KS:
        H.quiet(R.kmer_index, k)
        ix = R.kmer_dump(); ch = R.find_chains(off, bases)
        out["k%d_index_sha1" % k] = index_digest(ix)
        out["k%d_index_dims" % k] = np.array([len(ix["kmers"]), len(ix["edge_off"]) - 1, len(ix["edges"])], np.int64)
        for key in ("chain_off", "begin", "end", "edge_off", "edges"):
            out["k%d_%s" % (k, key)] = ch[key]
    path = os.path.join(HERE, "kmer_small.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
