#!/usr/bin/env python
"""Parity at scale (GPU box): the production cascade (k_prepare skip + warp extension kernels) against
  (1) the same path forced through the scalar extension DP and with duplicate chains aligned (independent code paths that
      must give identical per-pair results), over the whole batch, and
  (2) the CPU checker (compiled reference if oracle/_ref is present, else the restatement) on every pair where (1) disagrees
      plus a random sample of --check-pairs pairs, column by column.
Prints one JSON line; exit code 1 if anything differs from the checker."""
import argparse
import ctypes as C
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "tests"))
import numpy as np  # noqa: E402
import harness as H  # noqa: E402


subset = H.subset_pairs


def session_results(P, b, maxcol, mu, sd):
    L = P.lib
    sb = H.make_batch_struct(b); sess = C.c_void_p()
    P._chk(L.hlala_session_create(P.g, C.byref(sb), C.c_int32(maxcol), C.byref(sess)))
    t0 = time.time()
    P._chk(L.hlala_session_run(sess, C.c_double(mu), C.c_double(sd), C.c_uint64(0), None))
    nr = len(b["read_off"]) - 1
    o = dict(pair_mapq=np.zeros(nr // 2, np.float64), read_mapq=np.zeros(nr, np.float64), read_reverse=np.zeros(nr, np.uint8), chosen_slot=np.zeros(nr, np.int32), pair_ll=np.zeros(nr // 2, np.float64))
    po = H.PairOut(); po.max_columns = maxcol
    for k, v in o.items():
        setattr(po, k, v.ctypes.data)
    P._chk(L.hlala_session_fetch(sess, C.byref(po)))
    dig = (C.c_int64 * 4)(); sll = C.c_double(0)
    P._chk(L.hlala_session_digest(sess, dig, C.byref(sll)))
    o["digest"] = [int(x) for x in dig] + [sll.value]; o["seconds"] = time.time() - t0
    L.hlala_session_free(sess)
    return o


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--levels", type=int, default=250000)
    ap.add_argument("--pairs", type=int, default=1000000)
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--haps", type=int, default=8)
    ap.add_argument("--genes", type=int, default=4)
    ap.add_argument("--alleles", type=int, default=200)
    ap.add_argument("--check-pairs", type=int, default=3000)
    ap.add_argument("--max-columns", type=int, default=640)
    ap.add_argument("--seed", type=int, default=0xB200)
    ap.add_argument("--dir", default="/tmp/hlala_scale_parity")
    a = ap.parse_args()
    mu, sd = 100.0, 10.0
    d = os.path.join(a.dir, "prg_l%d" % a.levels)
    if not os.path.exists(os.path.join(d, ".complete")):
        os.makedirs(d, exist_ok=True)
        H.synth_prg(d, levels=a.levels, haps=a.haps, genes=a.genes, alleles=a.alleles, allele_contigs=4, seed=a.seed)
        open(os.path.join(d, ".complete"), "w").write("ok\n")
    b = H.synth_reads(d, os.path.join(d, "seeds_p%d.bin" % a.pairs), pairs=a.pairs, len=a.read_len, seed=a.seed + 1, clip_frac=0.15)
    P = H.Product(d); P.to_gpu(0)
    fast = session_results(P, b, a.max_columns, mu, sd)
    os.environ["HLALA_SCALAR_DP"] = "1"; os.environ["HLALA_ALIGN_DUPLICATES"] = "1"
    slow = session_results(P, b, a.max_columns, mu, sd)
    del os.environ["HLALA_SCALAR_DP"]; del os.environ["HLALA_ALIGN_DUPLICATES"]
    diff = np.zeros(a.pairs, bool)
    for k in ("pair_mapq", "pair_ll"):
        diff |= fast[k] != slow[k]
    for k in ("read_mapq", "chosen_slot", "read_reverse"):
        diff |= (fast[k] != slow[k]).reshape(-1, 2).any(1)
    dpairs = np.nonzero(diff)[0]
    rng = np.random.default_rng(a.seed)
    sample = rng.choice(a.pairs, size=min(a.check_pairs, a.pairs), replace=False)
    chk = np.unique(np.concatenate([dpairs[:2000], sample]))
    sub = subset(b, chk)
    got = P.pairs(sub, mu, sd, cap=a.max_columns, want_levels=False)
    t0 = time.time()
    want = H.oracle_pairs(d, sub, mu, sd, cap=a.max_columns)
    cpu_s = time.time() - t0
    n = want["n_cols"]

    def compare(got):
        bad = []
        for r in range(len(n)):
            ok = got["n_cols"][r] == n[r] and got["read_mapq"][r] == want["read_mapq"][r] and got["read_reverse"][r] == want["read_reverse"][r]
            if ok:
                for k in ("level", "edge", "gchar", "schar", "from_seed", "mapq"):
                    if not (got[k][r, :n[r]] == want[k][r, :n[r]]).all():
                        ok = False; break
            if not ok:
                bad.append(int(chk[r // 2]))
        return bad
    bad = compare(got)
    os.environ["HLALA_SCALAR_DP"] = "1"; os.environ["HLALA_ALIGN_DUPLICATES"] = "1"
    bad_scalar = compare(P.pairs(sub, mu, sd, cap=a.max_columns, want_levels=False))
    del os.environ["HLALA_SCALAR_DP"]; del os.environ["HLALA_ALIGN_DUPLICATES"]
    pm_bad = np.nonzero(got["pair_mapq"] != want["pair_mapq"])[0]
    # the sub-batch results of the checked pairs must also equal what the full-batch fast run produced for them
    full_vs_sub = int((fast["pair_ll"][chk] != got["pair_ll"]).sum() + (fast["pair_mapq"][chk] != got["pair_mapq"]).sum())
    line = {"tool": "scale_parity", "pairs": a.pairs, "levels": a.levels, "read_len": a.read_len, "cascade_digest": fast["digest"], "scalar_nodedup_digest": slow["digest"],
            "pairs_differing_cascade_vs_scalar": int(len(dpairs)), "first_differing": [int(x) for x in dpairs[:20]],
            "checked_against": want["oracle"], "pairs_checked": int(len(chk)), "pairs_differing_from_checker": sorted(set(bad))[:50], "scalar_path_pairs_differing_from_checker": sorted(set(bad_scalar))[:50],
            "pair_mapq_differing_from_checker": [int(chk[i]) for i in pm_bad[:50]], "full_batch_vs_subbatch_mismatches": full_vs_sub,
            "seconds": {"cascade": fast["seconds"], "scalar": slow["seconds"], "checker": cpu_s}}
    print(json.dumps(line))
    return 1 if (bad or bad_scalar or len(pm_bad) or len(dpairs) or full_vs_sub) else 0


if __name__ == "__main__":
    sys.exit(main())
