#!/usr/bin/env python
"""k-mer seeding probe: index build time, hlala_seed_kmers throughput on the bases of a synthetic seed batch, and (optionally) the
compiled reference's findChains on a bounded sample. usage: seed_probe.py [--levels N] [--pairs N] [--k K] [--alleles A] [--ref-reads N]"""
import argparse, json, os, sys, time
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "tests"))
import harness as H


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--levels", type=int, default=1000000); ap.add_argument("--pairs", type=int, default=250000); ap.add_argument("--k", type=int, default=25)
    ap.add_argument("--alleles", type=int, default=8); ap.add_argument("--genes", type=int, default=17); ap.add_argument("--len", type=int, default=150)
    ap.add_argument("--ref-reads", type=int, default=0); ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--dir", default=""); ap.add_argument("--ref-only", type=int, default=0, help="only time the compiled reference on the sample (no GPU needed)")
    a = ap.parse_args()
    d = a.dir or "/tmp/hlala_seed_probe_l%d_g%d_a%d" % (a.levels, a.genes, a.alleles)
    if not os.path.exists(d + "/.complete"):
        os.makedirs(d, exist_ok=True); H.synth_prg(d, levels=a.levels, haps=8, genes=a.genes, alleles=a.alleles, allele_contigs=4, seed=0xB200); open(d + "/.complete", "w").write("ok")
    b = H.synth_reads(d, os.path.join(d, "seeds_p%d.bin" % a.pairs), pairs=a.pairs, len=a.len, seed=0xB200, clip_frac=0.15)
    off = np.ascontiguousarray(b["read_off"], np.int64); bases = np.ascontiguousarray(b["bases"], np.uint8)
    if a.ref_only:
        R = H.quiet(H.Ref, d); t = time.time(); H.quiet(R.kmer_index, a.k); t_ref_index = time.time() - t
        n = min(a.ref_reads, len(off) - 1)
        ref = R.find_chains(off[:n + 1], bases)
        print(json.dumps(dict(ref_reads=n, ref_seconds=ref["seconds"], ref_reads_per_s=n / ref["seconds"], ref_index_build_s=t_ref_index, ref_chains=len(ref["begin"]))))
        return
    P = H.Product(d); P.to_gpu(0)
    t = time.time(); P.kmer_index(a.k); t_index = time.time() - t
    nk = len(P.kmer_dump()["kmers"]) if a.levels <= 2000000 else -1
    out = dict(levels=a.levels, reads=len(off) - 1, k=a.k, index_build_s=t_index, n_kmers=nk)
    for rep in range(a.reps):
        t = time.time(); got = P.find_chains(off, bases); dt = time.time() - t
        out.update(chains=len(got["begin"]), edges=len(got["edges"]), n_failed=got["n_failed"], kernel_ms=got["ms"][0], order_ms=got["ms"][1], e2e_s=dt,
                   reads_per_s_kernel=(len(off) - 1) / (got["ms"][0] / 1e3), reads_per_s_e2e=(len(off) - 1) / dt)
    if a.ref_reads and H.have_ref():
        R = H.quiet(H.Ref, d); H.quiet(R.kmer_index, a.k)
        n = min(a.ref_reads, len(off) - 1)
        ref = R.find_chains(off[:n + 1], bases)
        out.update(ref_reads=n, ref_seconds=ref["seconds"], ref_reads_per_s=n / ref["seconds"])
        same = all(np.array_equal(ref[k], (got[k][:len(ref[k])] if k != "chain_off" else got[k][:n + 1])) for k in ("chain_off", "begin", "end"))
        ne = ref["edge_off"][-1]
        same = same and np.array_equal(ref["edges"], got["edges"][:ne]) and np.array_equal(ref["edge_off"], got["edge_off"][:len(ref["edge_off"])])
        out["matches_reference_on_sample"] = bool(same)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
