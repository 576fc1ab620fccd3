"""k-mer index / findChains fuzz on small random graphs (own process per graph: one compiled-reference graph per process).
A synthetic PRG directory is generated and its PRG/graph.txt replaced by a random levelled graph over the same levels: 1-3 nodes per level,
random extra edges, gap edges with probability p_gap (bubbles that spell the same k-mer through different gap placements, gap runs, gaps at
the start), NODES and EDGES lines in shuffled order (so that canonical order != level order). Compared: the product's host index builder and
the oracle restatement against the unmodified GraphAndEdgeIndex (whole index, order included), and the oracle's findChains against the
reference's on walk reads.   usage: kmer_fuzz.py <seed> [k ...]"""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import harness as H  # noqa: E402


def write_random_graph(d, rng, max_w=3, p_gap=0.22, p_extra=0.35):
    path = os.path.join(d, "PRG", "graph.txt")
    names = []
    for ln in open(path, "rb").read().split(b"\n"):
        if ln == b"NODES:":
            break
        if b"|||" in ln:
            nm = ln.split(b"|||")[0].decode()
            if not names or names[-1] != nm:
                names.append(nm)
    N = len(names)                      # edge levels; node levels 0..N
    widths = [1] + [int(rng.randint(1, max_w + 1)) for _ in range(N)]
    node_id = {}; nodes = []
    for lvl, w in enumerate(widths):
        for z in range(w):
            node_id[(lvl, z)] = None; nodes.append((lvl, z))
    order = list(range(len(nodes))); rng.shuffle(order)
    for new_id, i in enumerate(order):
        node_id[nodes[i]] = new_id + 1
    syms = "ACGT_"
    edges = []
    for lvl in range(N):
        wf, wt = widths[lvl], widths[lvl + 1]; have = set()

        def add(f, t):
            for _ in range(20):
                s = "_" if (rng.rand() < p_gap and lvl != N - 1 and lvl != 0) else syms[rng.randint(0, 4)]
                if (f, t, s) not in have:
                    have.add((f, t, s)); edges.append((lvl, f, t, s)); return
        for t in range(wt):
            add(int(rng.randint(0, wf)), t)
        for f in range(wf):
            if not any(e[1] == f for e in edges if e[0] == lvl):
                add(f, int(rng.randint(0, wt)))
        for f in range(wf):
            for t in range(wt):
                if rng.rand() < p_extra:
                    add(f, t)
    eorder = list(range(len(edges))); rng.shuffle(eorder)
    with open(path, "wb") as f:
        f.write(b"CODE:\n")
        for nm in names:
            for i, s in enumerate(syms):
                f.write(("%s|||%s|||%d\n" % (nm, s, 49 + i)).encode())
        f.write(b"NODES:\n")
        for i in order:
            lvl, z = nodes[i]
            f.write(("%d|||%d|||%d\n" % (node_id[(lvl, z)], lvl, 1 if lvl == N else 0)).encode())
        f.write(b"EDGES:\n")
        for k, i in enumerate(eorder):
            lvl, fr, to, s = edges[i]
            f.write(("%d|||%s|||1|||" % (k + 1, names[lvl])).encode() + bytes([49 + syms.index(s)]) + ("|||%d|||%d||||||0\n" % (node_id[(lvl, fr)], node_id[(lvl + 1, to)])).encode())
    for c in ("graph.hlala_b200.cache",):
        p = os.path.join(d, "PRG", c)
        if os.path.exists(p):
            os.remove(p)
    return N, len(nodes), len(edges)


def same(a, b, keys):
    return all(a[k].shape == b[k].shape and (a[k] == b[k]).all() for k in keys)


def main():
    seed = int(sys.argv[1]); ks = [int(x) for x in sys.argv[2:]] or [3, 5]
    os.environ["HLALA_NO_GRAPH_CACHE"] = "1"
    rng = np.random.RandomState(seed)
    d = tempfile.mkdtemp(prefix="kmer_fuzz_")
    H.synth_prg(d, levels=int(rng.randint(40, 120)), haps=3, genes=0, alleles=4, seed=seed)
    dims = write_random_graph(d, rng, max_w=int(rng.randint(2, 4)), p_gap=float(rng.choice([0.0, 0.1, 0.22, 0.35])))
    R = H.quiet(H.Ref, d)
    G = R.graph()
    P = H.Product(d)
    assert P.dims()["n_nodes"] == dims[1] and P.dims()["n_edges"] == dims[2]
    IX = ("kmers", "pos_off", "edge_off", "edges")
    for k in ks:
        H.quiet(R.kmer_index, k)
        ref = R.kmer_dump()
        O = H.OracleKmer(G, k); P.kmer_index(k)
        assert same(ref, O.kmer_dump(), IX), "seed %d k %d: oracle index != reference" % (seed, k)
        assert same(ref, P.kmer_dump(), IX), "seed %d k %d: product index != reference" % (seed, k)
        off, bases = H.walk_reads(G, 120, 40, seed + k, err=0.03)
        a = R.find_chains(off, bases); b = O.find_chains(off, bases)
        assert H.same_chains(a, b), "seed %d k %d: oracle findChains != reference" % (seed, k)
        O.close()
        print("seed %d k %d: %d k-mers, %d positions, %d chains ok" % (seed, k, len(ref["kmers"]), len(ref["edge_off"]) - 1, len(a["begin"])))


if __name__ == "__main__":
    main()
