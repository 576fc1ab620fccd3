"""Long-read mode on the GPU: times hlala_align_long_reads (host buffers) on N synthetic reads of length L and compares a sample with the oracle.
usage: long_reads_probe.py [reads] [len] [sample]"""
import json
import os
import sys
import tempfile
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import harness as H  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 8000
    sample = int(sys.argv[3]) if len(sys.argv) > 3 else 100
    d = tempfile.mkdtemp(prefix="hlala_lr_")
    H.synth_prg(d, levels=1000000, haps=8, genes=17, alleles=200, seed=3)
    b = H.synth_reads(d, os.path.join(d, "lr.bin"), pairs=n, len=L, single=1, indel_rate=0.03, clip_frac=0.5, clip_max=400, gene_frac=0.3, seed=9)
    cap = 16384 if L > 7000 else 8192
    P = H.Product(d); P.to_gpu(0)
    out = {"reads": n, "len": L, "chains": int(len(b["chain_contig"])), "max_columns": cap}
    for it in range(3):
        t = time.time(); got = P.long_reads(b, cap, want_levels=True); out["seconds_%d" % it] = time.time() - t
    out["reads_per_s"] = n / out["seconds_2"]; out["bases_per_s"] = n * L / out["seconds_2"]
    out["mapq_lt_1"] = int((got["pair_mapq"] < 1).sum()); out["max_cols"] = int(got["n_cols"].max())
    # sample against the oracle
    k = min(sample, n)
    nch = int(b["chain_off"][k]); ncg = int(b["cigar_off"][nch]); nb = int(b["read_off"][k])
    sb = dict(read_off=b["read_off"][:k + 1].copy(), bases=b["bases"][:nb].copy(), quals=b["quals"][:nb].copy(), chain_off=b["chain_off"][:k + 1].copy(),
              chain_contig=b["chain_contig"][:nch].copy(), chain_pos=b["chain_pos"][:nch].copy(), chain_flag=b["chain_flag"][:nch].copy(), chain_as=b["chain_as"][:nch].copy(),
              cigar_off=b["cigar_off"][:nch + 1].copy(), cigar=b["cigar"][:ncg].copy())
    t = time.time(); want = H.oracle_long_reads(d, sb, cap); out["oracle_seconds"] = time.time() - t; out["oracle"] = want["oracle"]
    bad = 0
    for r in range(k):
        m = int(want["n_cols"][r]); ok = got["n_cols"][r] == m and got["pair_ll"][r] == want["read_ll"][r] and abs(got["pair_mapq"][r] - want["read_mapq"][r]) <= 1e-12
        for key in ("level", "edge", "gchar", "schar", "from_seed", "mapq"):
            ok = ok and np.array_equal(got[key][r, :m], want[key][r, :m])
        bad += 0 if ok else 1
    out["sample"] = k; out["sample_differences"] = bad
    print(json.dumps(out))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
