"""Ad-hoc GPU probe (development aid): product chains vs the compiled reference on a fresh synthetic PRG."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np
import harness as H


def quiet(fn, *a, **k):
    devnull = os.open(os.devnull, os.O_WRONLY); so, se = os.dup(1), os.dup(2)
    sys.stdout.flush(); sys.stderr.flush(); os.dup2(devnull, 1); os.dup2(devnull, 2)
    try:
        return fn(*a, **k)
    finally:
        os.dup2(so, 1); os.dup2(se, 2); os.close(devnull)


def main(which="S"):
    d = tempfile.mkdtemp(prefix="prg")
    if which == "S":
        H.synth_prg(d, levels=25000, haps=4, genes=2, alleles=64)
        b = H.synth_reads(d, d + "/seeds.bin", pairs=2000, len=100, clip_frac=0.15)
    else:   # gene blocks: ~14 nodes per level, ~11 chains per read
        H.synth_prg(d, levels=60000, haps=8, genes=17, alleles=1000)
        b = H.synth_reads(d, d + "/seeds.bin", pairs=600, len=150, clip_frac=0.15, gene_frac=1.0)
    R = quiet(H.Ref, d)
    t = time.time(); rc = quiet(R.chains, b); t_ref = time.time() - t
    P = H.Product(d); P.to_gpu(0)
    t = time.time(); pc = P.chains(b); t_gpu = time.time() - t
    print("chains", len(rc["status"]), "ref s", round(t_ref, 2), "gpu s", round(t_gpu, 3))
    print("status ref", np.unique(rc["status"], return_counts=True), "gpu", np.unique(pc["status"], return_counts=True))
    assert (rc["chain_order"] == pc["chain_order"]).all()
    ok = (rc["status"] == 0) & (pc["status"] == 0)
    L = b["read_off"][1] - b["read_off"][0]
    noext = ok
    print("ok", ok.sum(), "with extension", (ok & ~((rc["seed_begin"] == 0) & (rc["seed_end"] == L - 1))).sum())
    sb = (rc["seed_begin"] == pc["seed_begin"]) & (rc["seed_end"] == pc["seed_end"])
    print("seed begin/end mismatch among ok:", (~sb & ok).sum())
    bad = 0; badll = 0
    for i in np.nonzero(noext)[0]:
        n = rc["n_cols"][i]
        same = n == pc["n_cols"][i] and all((rc[k][i, :n] == pc[k][i, :n]).all() for k in ("level", "edge", "gchar", "schar", "from_seed"))
        if not same:
            bad += 1
            if bad <= 3:
                print("MISMATCH slot", i, "n", n, pc["n_cols"][i])
                for k in ("level", "edge", "gchar", "schar"):
                    print(k, rc[k][i, :n][:60], pc[k][i, :pc["n_cols"][i]][:60])
        if rc["ll"][i] != pc["ll"][i]:
            badll += 1
            if badll <= 3:
                print("LL differs", i, repr(rc["ll"][i]), repr(pc["ll"][i]))
    print("column mismatches", bad, "LL not bit-equal", badll, "of", noext.sum())
    # ---- pairs
    t = time.time(); rp = quiet(R.pairs, b, 100.0, 10.0); t_ref = time.time() - t
    t = time.time(); pp = P.pairs(b, 100.0, 10.0); t_gpu = time.time() - t
    print("pairs", len(rp["pair_mapq"]), "ref s", round(t_ref, 2), "gpu s", round(t_gpu, 3))
    badp = 0; badq = 0; badmq = 0
    for r in range(len(rp["n_cols"])):
        n = rp["n_cols"][r]
        same = n == pp["n_cols"][r] and all((rp[k][r, :n] == pp[k][r, :n]).all() for k in ("level", "edge", "gchar", "schar", "from_seed"))
        if not same:
            badp += 1
            if badp <= 3:
                print("PAIR MISMATCH read", r, n, pp["n_cols"][r]); print(rp["level"][r, :n][:40], pp["level"][r, :pp["n_cols"][r]][:40])
        elif not (rp["mapq"][r, :n] == pp["mapq"][r, :n]).all():
            badq += 1
            if badq <= 3:
                print("MAPQ chars differ read", r, rp["mapq"][r, :n][:50], pp["mapq"][r, :n][:50])
    dm = np.abs(rp["pair_mapq"] - pp["pair_mapq"]).max(); dr = np.abs(rp["read_mapq"] - pp["read_mapq"]).max()
    print("pair column mismatches", badp, "mapq-char mismatches", badq, "max |mapQ diff|", dm, dr, "pairs with mapQ<1", (rp["pair_mapq"] < 1).sum(),
          "bit-equal pair mapQ", (rp["pair_mapq"] == pp["pair_mapq"]).mean())
    print("bases_per_level sum", pp["bases_per_level"].sum())


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "S")
