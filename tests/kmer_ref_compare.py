"""Run by tests/test_kmer.py in a process of its own: the oracle restatement (oracle/hlala_oracle_kmer.cpp) and the product's host index
builder against the UNMODIFIED Graph/GraphAndEdgeIndex.cpp (oracle/_ref), on two synthetic PRGs and several k."""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import harness as H  # noqa: E402


def same(a, b, keys):
    return all(a[k].shape == b[k].shape and (a[k] == b[k]).all() for k in keys)


def main():
    d = tempfile.mkdtemp(prefix="kmer_ref_")
    H.synth_prg(d, levels=4000, haps=8, genes=2, alleles=64, seed=11)
    R = H.quiet(H.Ref, d)
    G = R.graph()
    P = H.Product(d)
    for k in (25, 12, 6):
        H.quiet(R.kmer_index, k)
        O = H.OracleKmer(G, k); P.kmer_index(k)
        ref = R.kmer_dump()
        assert same(ref, O.kmer_dump(), ("kmers", "pos_off", "edge_off", "edges")), "oracle index != reference (k=%d)" % k
        assert same(ref, P.kmer_dump(), ("kmers", "pos_off", "edge_off", "edges")), "product index != reference (k=%d)" % k
        off, bases = H.walk_reads(G, 250, 120, 40 + k)
        a = R.find_chains(off, bases); b = O.find_chains(off, bases)
        assert H.same_chains(a, b), "oracle findChains != reference (k=%d)" % k
        assert len(a["begin"]) > 200
        O.close()
    print("ok")


if __name__ == "__main__":
    main()
