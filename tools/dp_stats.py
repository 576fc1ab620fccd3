#!/usr/bin/env python
"""Work statistics of the extension DP on a bench-like workload (CPU only): runs hla-la_b200/csrc/extend_dp.h on the host with
-DHLALA_DP_STATS over every extension of a sample and prints what the capacity choices of the GPU tiers are based on."""
import argparse
import ctypes as C
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "tests"))
import numpy as np  # noqa: E402
import harness as H  # noqa: E402
import test_dp_host as T  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--levels", type=int, default=1250000)
    ap.add_argument("--genes", type=int, default=4)
    ap.add_argument("--alleles", type=int, default=1000)
    ap.add_argument("--haps", type=int, default=8)
    ap.add_argument("--pairs", type=int, default=1500)
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--dir", default="/tmp/hlala_dp_stats")
    a = ap.parse_args()
    lib = os.path.join(REPO, "tests", "native", "build", "libdp_host_stats.so")
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-DHLALA_DP_STATS", "-o", lib, "tests/native/dp_host.cpp", "hla-la_b200/host/prg_graph.cpp"], cwd=REPO, check=True)
    d = os.path.join(a.dir, "prg_l%d_g%d_a%d" % (a.levels, a.genes, a.alleles))
    if not os.path.exists(os.path.join(d, ".complete")):
        os.makedirs(d, exist_ok=True)
        H.synth_prg(d, levels=a.levels, haps=a.haps, genes=a.genes, alleles=a.alleles, allele_contigs=4, seed=0xB200)
        open(os.path.join(d, ".complete"), "w").write("ok\n")
    b = H.synth_reads(d, os.path.join(d, "seeds_stats_p%d.bin" % a.pairs), pairs=a.pairs, len=a.read_len, seed=0xB201, clip_frac=1.0)
    oc = H.Oracle(d).chains(b, 1024)
    T.LIB = lib
    n = T.check_dp_host(d, b, oc)
    L = C.CDLL(lib)
    L.dp_host_ext_stats.restype = C.c_longlong
    buf = np.zeros((4000000, 11), np.int32)
    m = L.dp_host_ext_stats(H.p(buf), C.c_longlong(len(buf)))
    s = buf[:m]
    np.save(os.path.join(a.dir, "ext_stats.npy"), s)
    names = ["clip", "diags", "cells", "max_td", "max_m1", "max_w", "max_vspread", "first_long_jump", "max_jump_len", "revisits", "max_deg"]
    print("chains compared", n, "extensions", m)
    for i, nm in enumerate(names):
        c = s[:, i]
        print("%-16s mean %9.2f  p50 %6d  p90 %6d  p99 %6d  max %6d" % (nm, c.mean(), np.percentile(c, 50), np.percentile(c, 90), np.percentile(c, 99), c.max()))
    cells = s[:, 2].astype(np.int64)
    for td in (8, 12, 16, 24, 32):
        for w in (2, 4, 8):
            for lj in (0, 1):
                ok = (s[:, 3] <= td) & (s[:, 5] <= w) & ((s[:, 7] == 0) | (lj == 1))
                print("max_td<=%2d max_w<=%d long-jumps %s: tasks %.3f  cells %.3f" % (td, w, "ok " if lj else "no ", ok.mean(), cells[ok].sum() / cells.sum()))


if __name__ == "__main__":
    main()
