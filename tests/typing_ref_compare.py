"""Run by tests/test_typing_oracle.py in a process of its own (the compiled-reference driver holds one graph per process):
unmodified reference HLATypeInference vs the oracle restatement on one dataset; prints a JSON verdict."""
import filecmp
import json
import os
import re
import sys

import numpy as np

import harness as H


def main():
    d, seeds, mu, sd, out = sys.argv[1], sys.argv[2], float(sys.argv[3]), float(sys.argv[4]), sys.argv[5]
    b = H.read_arrayfile(seeds)
    R = H.quiet(H.Ref, d)
    aln = H.quiet(R.pairs, b, mu, sd, 512)
    ref_dir = os.path.join(out, "ref", "hla"); or_dir = os.path.join(out, "oracle", "hla")
    r = H.quiet(R.type, d, b, mu, sd, ref_dir)
    T = H.OracleTyping(d, b, aln, mu, sd, or_dir)
    fr = sorted(os.listdir(ref_dir)); fo = sorted(os.listdir(or_dir))
    bad = [f for f in fr if f not in fo or not filecmp.cmp(os.path.join(ref_dir, f), os.path.join(or_dir, f), shallow=False)]
    # The third number of a pile-up summary ("allele x n [mean length; min strand freq; first-mate freq]") counts observations whose
    # verboseSeedChain::fromFirstRead is set. The reference never initialises that member (verboseSeedChain.cpp:13-17) and overwrites the
    # first mate's chain after setting it (processBAM.cpp:3545-3548), so for first-mate observations it reads indeterminate memory: usually 0
    # (what the restatement defines), occasionally not. Files that differ only there are compared with that number masked.
    undefined_only = []
    for f in list(bad):
        if f.startswith("R1_pileup_") and f in fo:
            mask = lambda t: re.sub(r";[-+0-9.e]+\]", ";*]", t)
            if mask(open(os.path.join(ref_dir, f)).read()) == mask(open(os.path.join(or_dir, f)).read()):
                bad.remove(f); undefined_only.append(f)
    dims = [T.locus(i) for i in range(T.n_loci)]
    print(json.dumps({"n_used": r["n_used"], "files_ref": len(fr), "files_oracle": len(fo), "differing": bad, "differing_only_in_uninitialised_first_mate_count": undefined_only, "n_loci": T.n_loci,
                      "min_C": min(x["C"] for x in dims), "min_R": min(x["R"] for x in dims), "finite": bool(all(np.isfinite(x["pair_ll"]).all() for x in dims))}))


if __name__ == "__main__":
    main()
