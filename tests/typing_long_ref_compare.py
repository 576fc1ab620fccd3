"""Run by tests/test_long_read_typing.py in a process of its own (the compiled-reference driver holds one graph per process): the UNMODIFIED reference in
long-read mode (alignOneLongRead per read, gene filter, HLATyper::HLATypeInference with unpaired reads and longReadsMode "ont2d") against the product —
mode "host": hla-la_b200/host/hla_typing.cpp with the test-only loop stand-in for the two kernels, fed with the reference's own alignments;
mode "gpu":  hlala_align_long_reads -> hlala_typing_blob_from_long_reads -> hlala_typer_infer on cuda:0.   Prints a JSON verdict."""
import ctypes as C
import filecmp
import json
import os
import sys

import numpy as np

import harness as H


def main():
    d, seeds, out, mode = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4]
    cap = int(sys.argv[5]) if len(sys.argv) > 5 else 4096
    b = dict(np.load(seeds)) if seeds.endswith(".npz") else H.read_arrayfile(seeds)
    R = H.quiet(H.Ref, d)
    ref_dir = os.path.join(out, "ref", "hla"); my_dir = os.path.join(out, mode, "hla"); os.makedirs(my_dir)
    r = H.quiet(R.type_long, d, b, ref_dir)
    if mode == "host":
        aln = H.quiet(R.long_reads, b, cap)
        lib = C.CDLL(os.path.join(H.REPO, "tests", "native", "build", "libtyping_host.so")); lib.typing_host_last_error.restype = C.c_char_p
        n = lib.typing_host_run_long(d.encode(), C.c_longlong(len(b["read_off"]) - 1), H.p(b["read_off"]), H.p(b["bases"]), H.p(b["quals"]), C.c_int(cap), H.p(aln["n_cols"]), H.p(aln["level"]),
                                     H.p(aln["gchar"]), H.p(aln["schar"]), H.p(aln["mapq"]), H.p(aln["read_reverse"]), H.p(np.ascontiguousarray(aln["read_mapq"])), my_dir.encode(), C.c_int(1))
        assert n == 17, lib.typing_host_last_error().decode()
        n_sel = r["n_used"]
    else:
        P = H.Product(d); P.to_gpu(0); T = H.ProductTyping(P, d)
        aln = P.long_reads(b, cap)
        blob, n_sel = T.long_read_blob(b, aln, cap)
        T.infer([blob], 0.0, 1.0, my_dir, keep_read_ll=False)
        T.close(); P.close()
    fr = sorted(os.listdir(ref_dir)); fm = sorted(os.listdir(my_dir))
    bad = [f for f in fr if f not in fm or not filecmp.cmp(os.path.join(ref_dir, f), os.path.join(my_dir, f), shallow=False)]
    near = []
    if mode == "gpu":
        # the allele-pair tables print 6 significant digits of sums that differ in the last bits on the device (max-shifted products instead of one exp + log per term):
        # same rows, values within 1e-5, the order may differ only among rows whose LL agree to that precision (as tests/test_gpu_typing.py does for paired reads)
        for f in list(bad):
            if not (f.startswith("R1_PP_") and f in fm):
                continue
            a = open(os.path.join(ref_dir, f)).read().split("\n"); g_ = open(os.path.join(my_dir, f)).read().split("\n")
            if len(a) != len(g_) or a[0] != g_[0]:
                continue
            ra = {l.split("\t")[0]: [float(x) for x in l.split("\t")[1:]] for l in a[1:] if l}; rg = {l.split("\t")[0]: [float(x) for x in l.split("\t")[1:]] for l in g_[1:] if l}
            if ra.keys() != rg.keys():
                continue
            ka = [l.split("\t")[0] for l in a[1:] if l]; kg = [l.split("\t")[0] for l in g_[1:] if l]
            va = np.array([ra[k] for k in ka]); vg = np.array([rg[k] for k in ka])
            ok = np.allclose(va, vg, rtol=1e-5, atol=0) and all(abs(ra[x][1] - ra[y][1]) <= 1e-5 * abs(ra[x][1]) for x, y in zip(ka, kg) if x != y)
            if ok:
                bad.remove(f); near.append(f)
    # how much of the long-read branch the dataset exercises, read off the reference's own pile-up: deepest column, columns where an allele was seen on one strand only >= 100 times
    deepest = 0; one_strand_100 = 0
    for f in fr:
        if f.startswith("R1_pileup_"):
            for line in open(os.path.join(ref_dir, f)):
                x = line.rstrip("\n").split("\t")
                if len(x) >= 3:
                    deepest = max(deepest, int(x[2]))
    print(json.dumps({"n_used": r["n_used"], "n_selected": int(n_sel), "files_ref": len(fr), "files_mine": len(fm), "differing": bad, "pair_tables_equal_to_1e-5": near, "deepest_pileup_column": deepest}))


if __name__ == "__main__":
    main()
