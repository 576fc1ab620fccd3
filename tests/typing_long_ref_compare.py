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
    b = H.read_arrayfile(seeds)
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
    # how much of the long-read branch the dataset exercises, read off the reference's own pile-up: deepest column, columns where an allele was seen on one strand only >= 100 times
    deepest = 0; one_strand_100 = 0
    for f in fr:
        if f.startswith("R1_pileup_"):
            for line in open(os.path.join(ref_dir, f)):
                x = line.rstrip("\n").split("\t")
                if len(x) >= 3:
                    deepest = max(deepest, int(x[2]))
    print(json.dumps({"n_used": r["n_used"], "n_selected": int(n_sel), "files_ref": len(fr), "files_mine": len(fm), "differing": bad, "deepest_pileup_column": deepest}))


if __name__ == "__main__":
    main()
