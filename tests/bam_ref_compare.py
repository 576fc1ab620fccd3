"""Run by tests/test_bam.py in a process of its own (one compiled-reference graph per process): the record selection, grouping, name
order and completeness rule of hlala_bam_read against the UNMODIFIED processBAM::extractSeeds2 + protoSeeds (oracle/_ref), which is fed
the same records through the in-memory stand-in for BamTools::BamReader. What stays unpinned is BamTools' own record decoding."""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import harness as H  # noqa: E402


def main():
    d = tempfile.mkdtemp(prefix="bam_ref_")
    H.synth_prg(d, levels=25000, haps=4, genes=2, alleles=64)
    b = H.synth_reads(d, os.path.join(d, "seeds.bin"), pairs=1500, len=100, clip_frac=0.15)
    R = H.quiet(H.Ref, d)
    P = H.Product(d)
    n_contigs = P.dims()["n_contigs"]; lens = np.diff(P.array("contig_off")); prg_ids = P.array("contig_prg_id")
    refs = [("PRG_%d" % i, int(lens[j])) for j, i in enumerate(prg_ids)] + [("chrUn", 1000000)]
    rng = np.random.RandomState(3)
    recs = []
    nr = len(b["read_off"]) - 1
    for r in range(nr):
        pair = r // 2; name = "q%d" % pair                      # unpadded: byte order of the names differs from the numeric order
        seq = bytes(b["bases"][b["read_off"][r]:b["read_off"][r + 1]]).decode(); qual = bytes(b["quals"][b["read_off"][r]:b["read_off"][r + 1]] - 33)
        for c in range(b["chain_off"][r], b["chain_off"][r + 1]):
            flag = int(b["chain_flag"][c] & 0x110) | 0x1 | (0x40 if r % 2 == 0 else 0x80)
            if pair % 50 == 7 and r % 2 == 1:
                flag |= 0x100                                    # second mate without a primary record -> incomplete pair
            if pair % 50 == 11 and c == b["chain_off"][r]:
                flag |= 0x4                                      # an unmapped record
            ref = int(b["chain_contig"][c])
            if pair % 50 == 13 and c == b["chain_off"][r + 1] - 1:
                ref = n_contigs                                  # a record on a reference that is not a PRG contig
            cig = [("MIDNSHP=X"[int(x) & 15], int(x) >> 4) for x in b["cigar"][b["cigar_off"][c]:b["cigar_off"][c + 1]]]
            pos = int(b["chain_pos"][c])
            if pair % 50 == 17 and c == b["chain_off"][r]:
                pos = refs[ref][1] - 10                          # hangs over the end of its contig: not inside the interval
            sec = bool(flag & 0x100)
            recs.append(dict(name=name, flag=flag, ref=ref, pos=pos, cigar=cig, seq="" if sec else seq, qual=b"" if sec else qual, tags={b"AS": int(b["chain_as"][c])}))
    order = sorted(range(len(recs)), key=lambda i: (recs[i]["ref"], recs[i]["pos"], rng.rand()))      # coordinate-sorted file
    recs = [recs[i] for i in order]
    bam = os.path.join(d, "t.bam"); H.write_bam(bam, refs, recs)
    got, names, st = P.bam_read(bam, threads=3)
    # the same records through the reference
    ref_a = np.array([x["ref"] for x in recs], np.int32); pos_a = np.array([x["pos"] for x in recs], np.int32); flag_a = np.array([x["flag"] for x in recs], np.uint16); as_a = np.array([x["tags"][b"AS"] for x in recs], np.int32)
    cig_off = np.zeros(len(recs) + 1, np.int32); cig = []
    for i, x in enumerate(recs):
        cig += [(ln << 4) | "MIDNSHP=X".index(op) for op, ln in x["cigar"]]; cig_off[i + 1] = len(cig)
    cig = np.array(cig, np.uint32)
    rn, comp, n1, n2, rr = H.quiet(R.extract_seeds, [x["name"] for x in recs], ref_a, pos_a, flag_a, as_a, cig_off, cig)
    assert rn == sorted(rn) and len(rn) == st["names"], (len(rn), st)
    want_names = [n for n, c in zip(rn, comp) if c]
    assert names == want_names, "pairs / their order differ from the reference's complete seeds"
    assert st["incomplete"] == int((comp == 0).sum()) and st["used"] == len(rr)
    at = 0; k = 0
    for n, c, a1, a2 in zip(rn, comp, n1, n2):
        idx1 = rr[at:at + a1]; idx2 = rr[at + a1:at + a1 + a2]; at += a1 + a2
        if not c:
            continue
        for m, idx in enumerate((idx1, idx2)):
            r = 2 * k + m; c0, c1 = got["chain_off"][r], got["chain_off"][r + 1]
            assert c1 - c0 == len(idx), (n, m)
            assert np.array_equal(got["chain_pos"][c0:c1], pos_a[idx]) and np.array_equal(got["chain_contig"][c0:c1], ref_a[idx])
            assert np.array_equal(got["chain_flag"][c0:c1], flag_a[idx]) and np.array_equal(got["chain_as"][c0:c1], as_a[idx])
            for j, ri in enumerate(idx):
                assert np.array_equal(got["cigar"][got["cigar_off"][c0 + j]:got["cigar_off"][c0 + j + 1]], cig[cig_off[ri]:cig_off[ri + 1]])
        k += 1
    assert k == len(names) and k > 1300 and (comp == 0).sum() > 20
    print("ok: %d names, %d complete pairs, %d records kept of %d" % (len(rn), k, len(rr), len(recs)))


if __name__ == "__main__":
    main()
