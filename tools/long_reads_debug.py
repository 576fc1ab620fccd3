"""debug: long-read stage sample vs the compiled reference, field by field"""
import os, sys, tempfile, json
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import harness as H
d = tempfile.mkdtemp(prefix="hlala_lrd_")
H.synth_prg(d, levels=294118, haps=8, genes=1, alleles=1000, allele_contigs=4, seed=0xB200)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
b = H.synth_reads(d, os.path.join(d, "lr.bin"), pairs=n, len=8000, single=1, indel_rate=0.03, clip_frac=0.5, clip_max=400, gene_frac=0.2, seed=0xB200)
cap = 10240
P = H.Product(d); P.to_gpu(0)
got = P.long_reads(b, cap)
want = H.oracle_long_reads(d, b, cap)
print("oracle", want["oracle"])
bad = 0
for r in range(n):
    m = int(want["n_cols"][r]); diffs = []
    if got["n_cols"][r] != m: diffs.append("n_cols %d vs %d" % (got["n_cols"][r], m))
    if got["pair_ll"][r] != want["read_ll"][r]: diffs.append("ll %.17g vs %.17g" % (got["pair_ll"][r], want["read_ll"][r]))
    if abs(got["pair_mapq"][r] - want["read_mapq"][r]) > 1e-12: diffs.append("mapq %.17g vs %.17g" % (got["pair_mapq"][r], want["read_mapq"][r]))
    if got["read_reverse"][r] != want["read_reverse"][r]: diffs.append("reverse")
    mm = min(m, int(got["n_cols"][r]))
    for key in ("level", "edge", "gchar", "schar", "from_seed", "mapq"):
        w = np.nonzero(got[key][r, :mm] != want[key][r, :mm])[0]
        if len(w): diffs.append("%s at %d cols, first %d: %s vs %s" % (key, len(w), w[0], got[key][r, w[0]], want[key][r, w[0]]))
    if diffs:
        bad += 1
        nch = b["chain_off"][r + 1] - b["chain_off"][r]
        if bad <= 12: print("read", r, "chains", nch, "len", b["read_off"][r + 1] - b["read_off"][r], "|", "; ".join(diffs))
print("reads", n, "different", bad)
