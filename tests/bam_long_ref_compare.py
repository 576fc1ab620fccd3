"""Run by tests/test_bam.py in a process of its own: hlala_bam_read_long (long-read mode of the BAM ingest: primary records only, unpaired, one read per name) against the
UNMODIFIED processBAM::extractSeeds2(.., longReadMode) + protoSeeds::isComplete_unpaired over the same records (in-memory BamReader stand-in)."""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import harness as H  # noqa: E402


def main():
    d = tempfile.mkdtemp(prefix="bam_long_ref_")
    H.synth_prg(d, levels=25000, haps=4, genes=2, alleles=32)
    b = H.synth_reads(d, os.path.join(d, "seeds.bin"), pairs=300, len=1200, single=1, indel_rate=0.02, clip_frac=0.4, clip_max=150, seed=3)
    R = H.quiet(H.Ref, d)
    P = H.Product(d)
    n_contigs = P.dims()["n_contigs"]; lens = np.diff(P.array("contig_off")); prg_ids = P.array("contig_prg_id")
    refs = [("PRG_%d" % i, int(lens[j])) for j, i in enumerate(prg_ids)] + [("chrUn", 1000000)]
    rng = np.random.RandomState(5)
    recs = []
    nr = len(b["read_off"]) - 1
    for r in range(nr):
        name = "q%d" % r
        seq = bytes(b["bases"][b["read_off"][r]:b["read_off"][r + 1]]).decode(); qual = bytes(b["quals"][b["read_off"][r]:b["read_off"][r + 1]] - 33)
        for c in range(b["chain_off"][r], b["chain_off"][r + 1]):
            flag = int(b["chain_flag"][c] & 0x110)               # secondary records stay in the file: the long-read mode must drop them
            if r % 40 == 7:
                flag |= 0x100                                    # a read without any primary record -> never seen at all
            if r % 40 == 9 and not (flag & 0x100):
                flag |= 0x800                                    # supplementary records count as primary (BamTools::IsPrimaryAlignment looks at 0x100 only)
            if r % 40 == 11 and c == b["chain_off"][r]:
                flag |= 0x4
            ref = int(b["chain_contig"][c])
            if r % 40 == 13 and not (flag & 0x100):
                ref = n_contigs                                  # the primary record sits on a reference that is not a PRG contig
            cig = [("MIDNSHP=X"[int(x) & 15], int(x) >> 4) for x in b["cigar"][b["cigar_off"][c]:b["cigar_off"][c + 1]]]
            sec = bool(flag & 0x100)
            recs.append(dict(name=name, flag=flag, ref=ref, pos=int(b["chain_pos"][c]), cigar=cig, seq="" if sec else seq, qual=b"" if sec else qual, tags={b"AS": int(b["chain_as"][c])}))
        if r % 40 == 15:                                         # a second primary (supplementary-like) record of the same read elsewhere: both are kept, the first after AS sorting stands for the read
            c = b["chain_off"][r]
            cig = [("MIDNSHP=X"[int(x) & 15], int(x) >> 4) for x in b["cigar"][b["cigar_off"][c]:b["cigar_off"][c + 1]]]
            recs.append(dict(name=name, flag=int(b["chain_flag"][c] & 0x10) | 0x800, ref=int(b["chain_contig"][c]), pos=int(b["chain_pos"][c]), cigar=cig, seq=seq, qual=qual, tags={b"AS": int(b["chain_as"][c]) - 3}))
    order = sorted(range(len(recs)), key=lambda i: (recs[i]["ref"], recs[i]["pos"], rng.rand()))
    recs = [recs[i] for i in order]
    bam = os.path.join(d, "t.bam"); H.write_bam(bam, refs, recs)
    got, names, st = P.bam_read(bam, threads=3, long_reads=True)
    ref_a = np.array([x["ref"] for x in recs], np.int32); pos_a = np.array([x["pos"] for x in recs], np.int32); flag_a = np.array([x["flag"] for x in recs], np.uint16); as_a = np.array([x["tags"][b"AS"] for x in recs], np.int32)
    cig_off = np.zeros(len(recs) + 1, np.int32); cig = []
    for i, x in enumerate(recs):
        cig += [(ln << 4) | "MIDNSHP=X".index(op) for op, ln in x["cigar"]]; cig_off[i + 1] = len(cig)
    cig = np.array(cig, np.uint32)
    R.lib.hlala_ref_extract_seeds_long_mode(1)
    rn, comp, n1, n2, rr = H.quiet(R.extract_seeds, [x["name"] for x in recs], ref_a, pos_a, flag_a, as_a, cig_off, cig)
    assert rn == sorted(rn) and len(rn) == st["names"] and (n2 == 0).all(), (len(rn), st)
    want_names = [n for n, c in zip(rn, comp) if c]
    assert names == want_names, "reads / their order differ from the reference's complete seeds"
    assert st["incomplete"] == int((comp == 0).sum()) and st["used"] == len(rr)
    at = 0; k = 0
    for n, c, a1 in zip(rn, comp, n1):
        idx = rr[at:at + a1]; at += a1
        if not c:
            continue
        c0, c1 = got["chain_off"][k], got["chain_off"][k + 1]
        assert c1 - c0 == len(idx), n
        assert np.array_equal(got["chain_pos"][c0:c1], pos_a[idx]) and np.array_equal(got["chain_contig"][c0:c1], ref_a[idx])
        assert np.array_equal(got["chain_flag"][c0:c1], flag_a[idx]) and np.array_equal(got["chain_as"][c0:c1], as_a[idx])
        for j, ri in enumerate(idx):
            assert np.array_equal(got["cigar"][got["cigar_off"][c0 + j]:got["cigar_off"][c0 + j + 1]], cig[cig_off[ri]:cig_off[ri + 1]])
        # the read's bases are those of its first primary record after the AS sort
        r = int(n[1:]); assert got["read_off"][k + 1] - got["read_off"][k] == b["read_off"][r + 1] - b["read_off"][r]
        assert np.array_equal(got["bases"][got["read_off"][k]:got["read_off"][k + 1]], b["bases"][b["read_off"][r]:b["read_off"][r + 1]])
        k += 1
    assert k == len(names) and k > 250 and len(rn) < nr and (n1 > 1).sum() >= 5
    print("ok: %d names, %d complete reads, %d records kept of %d" % (len(rn), k, len(rr), len(recs)))


if __name__ == "__main__":
    main()
