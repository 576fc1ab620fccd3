#!/usr/bin/env python
"""GPU-box debug helper: per-chain comparison of the warp extension cascade against the scalar DP on tests/golden/cascade_regress_pairs.npz."""
import os, sys, json
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "tests"))
import numpy as np
import harness as H
d = "/tmp/hlala_scale_parity/prg_l250000"
if not os.path.exists(d + "/.complete"):
    os.makedirs(d, exist_ok=True)
    H.synth_prg(d, levels=250000, haps=8, genes=4, alleles=200, allele_contigs=4, seed=0xB200); open(d + "/.complete", "w").write("ok")
_g = np.load(os.path.join(REPO, "tests/golden/cascade_regress_pairs.npz")); b = {k[3:]: _g[k] for k in _g.files if k.startswith("in_")}
P = H.Product(d); P.to_gpu(0)
fast = P.chains(b, cap=640)
os.environ["HLALA_SCALAR_DP"] = "1"
slow = P.chains(b, cap=640)
del os.environ["HLALA_SCALAR_DP"]
slot_read = np.repeat(np.arange(len(b["chain_off"]) - 1), np.diff(b["chain_off"]))
for s in range(len(fast["n_cols"])):
    same = all(np.array_equal(fast[k][s], slow[k][s]) for k in ("status", "n_cols", "ll", "level", "edge", "schar", "gchar", "from_seed"))
    if same:
        continue
    n1, n2 = int(fast["n_cols"][s]), int(slow["n_cols"][s])
    print("slot", s, "read", slot_read[s], "status", fast["status"][s], slow["status"][s], "n_cols", n1, n2, "seed", slow["seed_begin"][s], slow["seed_end"][s], "ll", fast["ll"][s], slow["ll"][s])
    for name, o, n in (("fast", fast, n1), ("slow", slow, n2)):
        fs = o["from_seed"][s, :n]
        nz = np.nonzero(fs)[0]
        lo, hi = (nz[0], nz[-1]) if len(nz) else (0, -1)
        def seg(a, bb):
            return " ".join("%d:%s%s" % (o["level"][s, i], chr(o["gchar"][s, i]), chr(o["schar"][s, i])) for i in range(a, bb))
        print(" ", name, "left :", seg(0, lo))
        print(" ", name, "right:", seg(hi + 1, n))
