"""Alignment-path fuzz on small random graphs (own process per graph). The graph of a synthetic PRG directory is replaced by a random levelled
graph (tests/kmer_fuzz.py: gap bubbles, gap runs, shuffled node / edge order) and its contigs by random walks through that graph; reads
with soft clips and indels are drawn from the contigs. Compared, bit for bit: the oracle restatement (oracle/hlala_oracle.cpp) against the
UNMODIFIED reference (per-chain columns + log-likelihoods, per-pair columns + mapping qualities), and the extension DP the GPU executes
(hla-la_b200/csrc/extend_dp.h on the host) against both. With --gpu (a box with a GPU) the product's kernels are compared as well (not part of
the suite yet: first GPU run pending).   usage: align_fuzz.py <seed> [--gpu]"""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import harness as H  # noqa: E402
import kmer_fuzz  # noqa: E402
from test_dp_host import check_dp_host, check_lean_host  # noqa: E402


def rewrite_contigs(d, G, rng):
    """every contig of sequences.txt becomes a random walk from level 0 to the last level"""
    ef, et, em, nl = G["edge_from"], G["edge_to"], G["edge_emis"], G["node_level"]
    order = np.argsort(ef, kind="stable"); starts = np.searchsorted(ef[order], np.arange(len(nl) + 1))
    n_levels = int(nl.max()) + 1
    lines = open(os.path.join(d, "sequences.txt")).read().split("\n")
    fa = {}; out_lines = [lines[0]]
    for ln in lines[1:]:
        if not ln.strip():
            continue
        f = ln.split("\t"); sid = int(f[0])
        node = int(np.nonzero(nl == 0)[0][0]); seq = []; lv = []
        while nl[node] < n_levels - 1:
            a, b = starts[node], starts[node + 1]; e = order[a + rng.randint(0, b - a)]
            if em[e] != ord("_"):
                seq.append(chr(em[e])); lv.append(int(nl[node]))
            node = et[e]
        fa["PRG_%d" % sid] = "".join(seq)
        with open(os.path.join(d, "translation", "%d.txt" % sid), "w") as t:
            t.write("".join("%d\n" % x for x in lv))
        f[5] = str(len(seq)); out_lines.append("\t".join(f))
    open(os.path.join(d, "sequences.txt"), "w").write("\n".join(out_lines) + "\n")
    for p in (os.path.join(d, "mapping_PRGonly", "referenceGenome.fa"), os.path.join(d, "extendedReferenceGenome", "extendedReferenceGenome.fa")):
        with open(p, "w") as o:
            for k, s in fa.items():
                o.write(">%s\n" % k + "\n".join(s[i:i + 80] for i in range(0, len(s), 80)) + "\n")


def check_loader(d, R):
    """the product's flat graph (canonical order, gap-edge paths, jump lists in the reference's iteration order, gap stretches) against the
    reference's pointer graph on this shuffled, gap-dense graph"""
    rg = R.graph(); P = H.Product(d)
    node_ord = P.array("node_ord"); eo = P.array("edge_ord"); ef = P.array("edge_from"); et = P.array("edge_to")
    assert np.array_equal(rg["node_level"][node_ord], P.array("node_level"))
    assert np.array_equal(rg["edge_from"][eo], node_ord[ef]) and np.array_equal(rg["edge_to"][eo], node_ord[et]) and np.array_equal(rg["edge_emis"][eo], P.array("edge_emis"))
    assert np.array_equal(P.array("path_off"), rg["path_off"]) and np.array_equal(eo[P.array("path_edges")], rg["path_edges"]), "gap-edge paths"
    assert np.array_equal(P.array("gap_stretch"), rg["gap_stretch"]), "gap stretches"
    for name, key, src, dst in (("jump_fwd", "jumps_fwd", "path_from", "path_to"), ("jump_bwd", "jumps_bwd", "path_to", "path_from")):
        off = P.array(name + "_off"); lst = P.array(name + "_path"); a, b_, c = rg[key]
        mine = [(int(node_ord[n]), int(node_ord[P.array(dst)[p]]), int(p)) for n in np.argsort(node_ord) for p in lst[off[n]:off[n + 1]]]
        assert mine == list(zip(a.tolist(), b_.tolist(), c.tolist())), name
    P.close()


def add_secondary_chains(b, rng, contig_len):
    """about half of the reads get extra secondary records: the primary's CIGAR at a random place of a random contig (same or other strand,
    lower score), and sometimes an exact duplicate of the primary's coordinates — the pair stage then has several combinations to rank,
    chains on the wrong strand to skip and duplicates to drop"""
    nr = len(b["read_off"]) - 1
    out = {k: [] for k in ("chain_contig", "chain_pos", "chain_flag", "chain_as")}; cig = []; cig_off = [0]; chain_off = [0]
    for r in range(nr):
        c0, c1 = int(b["chain_off"][r]), int(b["chain_off"][r + 1])
        recs = [(int(b["chain_contig"][c]), int(b["chain_pos"][c]), int(b["chain_flag"][c]), int(b["chain_as"][c]), list(b["cigar"][b["cigar_off"][c]:b["cigar_off"][c + 1]])) for c in range(c0, c1)]
        if recs and rng.rand() < 0.5:
            ctg, pos, flag, as_, cg = recs[0]
            reflen = sum(int(x) >> 4 for x in cg if (int(x) & 15) in (0, 2, 3, 7, 8))
            for _ in range(int(rng.randint(1, 4))):
                u = rng.rand()
                if u < 0.2:
                    recs.append((ctg, pos, flag | 0x100, as_ - 1, cg))                       # same coordinates: dropped as a duplicate
                else:
                    c2 = int(rng.randint(0, len(contig_len)))
                    if contig_len[c2] > reflen + 2:
                        p2 = int(rng.randint(0, contig_len[c2] - reflen - 1))
                        f2 = (flag | 0x100) ^ (0x10 if rng.rand() < 0.3 else 0)
                        recs.append((c2, p2, f2, as_ - int(rng.randint(0, 30)), cg))
            order = list(range(len(recs))); rng.shuffle(order); recs = [recs[i] for i in order]
        for ctg, pos, flag, as_, cg in recs:
            out["chain_contig"].append(ctg); out["chain_pos"].append(pos); out["chain_flag"].append(flag); out["chain_as"].append(as_)
            cig += [int(x) for x in cg]; cig_off.append(len(cig))
        chain_off.append(len(out["chain_contig"]))
    nb = dict(b)
    nb["chain_off"] = np.array(chain_off, b["chain_off"].dtype); nb["cigar_off"] = np.array(cig_off, b["cigar_off"].dtype); nb["cigar"] = np.array(cig, b["cigar"].dtype)
    for k in out:
        nb[k] = np.array(out[k], b[k].dtype)
    return nb


def main():
    seed = int(sys.argv[1])
    os.environ["HLALA_NO_GRAPH_CACHE"] = "1"
    rng = np.random.RandomState(seed)
    d = tempfile.mkdtemp(prefix="align_fuzz_")
    H.synth_prg(d, levels=int(rng.randint(500, 900)), haps=3, genes=0, alleles=4, seed=seed)
    kmer_fuzz.write_random_graph(d, rng, max_w=int(rng.randint(2, 4)), p_gap=float(rng.choice([0.02, 0.08, 0.15])), p_extra=0.25)
    P = H.Product(d)
    G = dict(node_level=P.array("node_level"), edge_from=P.array("edge_from"), edge_to=P.array("edge_to"), edge_emis=P.array("edge_emis"))
    rewrite_contigs(d, G, rng); P.close()
    Pn = H.Product(d); P_dims = np.diff(Pn.array("contig_off")); Pn.close()
    b = H.synth_reads(d, os.path.join(d, "seeds.bin"), pairs=120, len=60, seed=seed, clip_frac=0.5, indel_rate=0.01, gap_mean=30, gap_sd=6)
    b = add_secondary_chains(b, rng, P_dims)
    R = H.quiet(H.Ref, d); O = H.Oracle(d)
    check_loader(d, R)
    rc = H.quiet(R.chains, b, 512); oc = O.chains(b, 512)
    for k in ("chain_order", "status", "n_cols", "seed_begin", "seed_end", "ll"):
        assert np.array_equal(rc[k], oc[k]), "seed %d: chains %s: oracle != reference" % (seed, k)
    n = rc["n_cols"]
    for i in range(len(n)):
        for k in ("level", "edge", "gchar", "schar", "from_seed"):
            assert np.array_equal(rc[k][i, :n[i]], oc[k][i, :n[i]]), "seed %d slot %d %s" % (seed, i, k)
    rp = H.quiet(R.pairs, b, 30.0, 6.0, 512); op = O.pairs(b, 30.0, 6.0, 512)
    assert np.array_equal(rp["n_cols"], op["n_cols"]) and np.array_equal(rp["pair_mapq"], op["pair_mapq"]) and np.array_equal(rp["read_mapq"], op["read_mapq"]), "seed %d: pairs" % seed
    m = rp["n_cols"]
    for r in range(len(m)):
        for k in ("level", "edge", "gchar", "schar", "from_seed", "mapq"):
            assert np.array_equal(rp[k][r, :m[r]], op[k][r, :m[r]]), "seed %d read %d %s" % (seed, r, k)
    tested = check_dp_host(d, b, rc)
    lean_ok, lean_deferred = check_lean_host(d, b, rc)      # the first GPU tier (extend_lean.h) against extend_dp.h
    if "--gpu" in sys.argv:
        Pg = H.Product(d); Pg.to_gpu(0)
        gc = Pg.chains(b, 512)
        for k in ("chain_order", "status", "n_cols", "seed_begin", "seed_end", "ll"):
            assert np.array_equal(gc[k], rc[k]), "seed %d: GPU chains %s" % (seed, k)
        for i in range(len(n)):
            for k in ("level", "edge", "gchar", "schar", "from_seed"):
                assert np.array_equal(gc[k][i, :n[i]], rc[k][i, :n[i]]), "seed %d GPU slot %d %s" % (seed, i, k)
        gp = Pg.pairs(b, 30.0, 6.0, 512)
        assert np.array_equal(gp["n_cols"], rp["n_cols"]) and np.allclose(gp["pair_mapq"], rp["pair_mapq"], rtol=0, atol=1e-12)
        for r in range(len(m)):
            for k in ("level", "edge", "gchar", "schar", "from_seed", "mapq"):
                assert np.array_equal(gp[k][r, :m[r]], rp[k][r, :m[r]]), "seed %d GPU read %d %s" % (seed, r, k)
        print("seed %d: GPU kernels ok" % seed)
    print("seed %d: %d chains, %d extended by extend_dp.h, lean tier %d equal / %d deferred, %d pairs ok" % (seed, int((rc["status"] == 0).sum()), tested, lean_ok, lean_deferred, len(m) // 2))


if __name__ == "__main__":
    main()
