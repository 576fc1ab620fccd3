"""Runs the compiled reference (oracle/_ref) on one batch in a process of its own: the reference's graph arena holds one PRG per process, so
harness.oracle_pairs() sends every further dataset of a test session here.   usage: ref_worker.py <prg_dir> <in.npz> <out.npz> <mean> <sd> <cap> [pairs|chains|long_reads]"""
import sys

import numpy as np

import harness as H


def main():
    prg, fin, fout, mu, sd, cap = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4]), float(sys.argv[5]), int(sys.argv[6])
    b = dict(np.load(fin))
    what = sys.argv[7] if len(sys.argv) > 7 else "pairs"
    R = H.quiet(H.Ref, prg)
    r = H.quiet(R.chains, b, cap) if what == "chains" else H.quiet(R.long_reads, b, cap) if what == "long_reads" else H.quiet(R.pairs, b, mu, sd, cap)
    np.savez(fout, **{k: np.asarray(v) for k, v in r.items() if isinstance(v, (np.ndarray, float))})


if __name__ == "__main__":
    main()
