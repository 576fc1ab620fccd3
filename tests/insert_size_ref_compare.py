"""Run by tests/test_bam.py in a process of its own (one compiled-reference graph per process): the insert-size estimate of the product
(sample selection of hlala_bam_read + alignment of the sample's primary records + hlala_insert_size_from_levels; with --gpu the whole
hlala_bam_insert_size on the GPU) against the UNMODIFIED processBAM::estimateInsertSize (oracle/_ref), which reads the same records
through the in-memory stand-in for BamTools::BamReader with its translation tables loaded lazily as in a fresh run.
On the CPU the alignments of the sample come from the oracle restatement (pinned to the reference elsewhere)."""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import harness as H  # noqa: E402


def main():
    gpu = "--gpu" in sys.argv
    d = tempfile.mkdtemp(prefix="is_ref_")
    H.synth_prg(d, levels=40000, haps=4, genes=2, alleles=48, seed=9)
    b = H.synth_reads(d, os.path.join(d, "seeds.bin"), pairs=7000, len=100, clip_frac=0.15, seed=9, gap_mean=180, gap_sd=25)
    R = H.quiet(H.Ref, d)
    P = H.Product(d)
    lens = np.diff(P.array("contig_off")); prg_ids = P.array("contig_prg_id")
    refs = [("PRG_%d" % i, int(lens[j])) for j, i in enumerate(prg_ids)]
    rng = np.random.RandomState(4); recs = []; nr = len(b["read_off"]) - 1
    for r in range(nr):
        pair = r // 2; name = "q%d" % pair
        seq = bytes(b["bases"][b["read_off"][r]:b["read_off"][r + 1]]).decode(); qual = bytes(b["quals"][b["read_off"][r]:b["read_off"][r + 1]] - 33)
        for c in range(b["chain_off"][r], b["chain_off"][r + 1]):
            flag = int(b["chain_flag"][c] & 0x110) | 0x1 | (0x40 if r % 2 == 0 else 0x80)
            if pair % 40 == 7 and r % 2 == 1:
                flag |= 0x100                                    # a mate without a primary record -> incomplete pair
            cig = [("MIDNSHP=X"[int(x) & 15], int(x) >> 4) for x in b["cigar"][b["cigar_off"][c]:b["cigar_off"][c + 1]]]
            sec = bool(flag & 0x100)
            recs.append(dict(name=name, flag=flag, ref=int(b["chain_contig"][c]), pos=int(b["chain_pos"][c]), cigar=cig, seq="" if sec else seq, qual=b"" if sec else qual, tags={b"AS": int(b["chain_as"][c])}))
    order = sorted(range(len(recs)), key=lambda i: (recs[i]["ref"], recs[i]["pos"], rng.rand()))      # coordinate-sorted file
    recs = [recs[i] for i in order]
    bam = os.path.join(d, "t.bam"); H.write_bam(bam, refs, recs)
    if gpu:
        P.to_gpu(0)
    sample, loaded, est = P.bam_insert_size(bam, gpu=gpu)
    # the reference on the same records
    ref_a = np.array([x["ref"] for x in recs], np.int32); pos_a = np.array([x["pos"] for x in recs], np.int32); flag_a = np.array([x["flag"] for x in recs], np.uint16); as_a = np.array([x["tags"][b"AS"] for x in recs], np.int32)
    cig_off = np.zeros(len(recs) + 1, np.int32); cig = []; seq_off = np.zeros(len(recs) + 1, np.int64); seqs = []; quals = []
    for i, x in enumerate(recs):
        cig += [(ln << 4) | "MIDNSHP=X".index(op) for op, ln in x["cigar"]]; cig_off[i + 1] = len(cig)
        seqs.append(x["seq"].encode()); quals.append(bytes(q + 33 for q in x["qual"])); seq_off[i + 1] = seq_off[i] + len(x["seq"])
    seq = np.frombuffer(b"".join(seqs) + b"\0", np.uint8).copy(); qual = np.frombuffer(b"".join(quals) + b"\0", np.uint8).copy()
    want = (R.estimate_insert_size if "--verbose" in sys.argv else lambda *a: H.quiet(R.estimate_insert_size, *a))([x["name"] for x in recs], ref_a, pos_a, flag_a, as_a, cig_off, np.array(cig, np.uint32), seq_off, seq, qual)
    # host arithmetic on the restatement's alignments of the sample
    oc = H.Oracle(d).chains(sample, 640)
    assert (oc["status"] == 0).all() and (oc["chain_order"] == np.arange(len(oc["status"]))).all()
    fl = np.full(len(oc["status"]), -1, np.int32); ll = np.full(len(oc["status"]), -1, np.int32)
    for i, n in enumerate(oc["n_cols"]):
        lv = oc["level"][i, :n]; lv = lv[lv != -1]
        if len(lv):
            fl[i], ll[i] = lv[0], lv[-1]
    rev = ((sample["chain_flag"] & 0x10) != 0).astype(np.uint8)
    got = P.insert_size_from_levels(fl, ll, rev, loaded)
    if "--verbose" in sys.argv:
        print("all contigs loaded:", P.insert_size_from_levels(fl, ll, rev, np.arange(len(refs), dtype=np.int32)), "loaded", list(loaded), "of", len(refs))
    n_pairs = (len(sample["read_off"]) - 1) // 2
    assert 1500 < n_pairs < 4000 and 0 < len(loaded) <= len(refs), (n_pairs, len(loaded), len(refs))      # the scan stopped early: only part of the pairs are in the sample
    assert P.insert_size_from_levels(fl, ll, rev, loaded[:2])[:2] != got[:2]                               # ... and which translations are loaded by then matters
    assert (got[0], got[1]) == want, (got, want)
    if gpu:
        assert (est[0], est[1]) == want and est[2] == got[2] and est[3] == got[3], (est, got, want)
    print("ok: insert size %g +- %g from %d pairs of the sample (%d skipped), %d of %d contigs loaded%s" % (got[0], got[1], got[2], got[3], len(loaded), len(refs), "; GPU estimate identical" if gpu else ""))


if __name__ == "__main__":
    main()
