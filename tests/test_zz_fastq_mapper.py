"""FASTQ input without bwa: hlala_fastq_map_pairs places paired reads on the PRG's linear contigs and returns the batch hlala_bam_read would return
for their `bwa mem -a -M` BAM. There is no oracle for the mapper itself (bwa is an external program, absent here); what is measured is what the reference's own
testPRGMapping actions measure (HLA-LA.cpp:1386-1621): the fraction of bases of simulated reads that the alignment path puts on their true level — against the
same figure for the generator's own seeds (the chains a perfect mapper would report). The file sorts last on purpose: the mapper is the newest part."""
import gzip
import os
import subprocess

import numpy as np
import pytest

import harness as H
from conftest import DATASETS


def _mapped(dataset, name, tmp_path):
    d, _b, mu, sd = dataset(name)
    pre = str(tmp_path / "R")
    b = H.synth_reads(d, str(tmp_path / "seeds.bin"), levels_prefix=pre, **DATASETS[name][1])
    P = H.Product(d)
    mb, names, cnt = P.fastq_map(pre + "_1.fq", pre + "_2.fq", threads=4)
    return d, b, mu, sd, pre, P, mb, names, cnt


def _fraction(aln, pre, names):
    _per, tot, _n = H.truth_evaluate(aln, pre + "_1.levels", pre + "_2.levels", names=names)
    return tot[1] / tot[0]


@pytest.mark.parametrize("name,floor", [("S", 0.985), ("genes", 0.99)])
def test_mapped_fastq_reaches_the_accuracy_of_the_generators_seeds(dataset, tmp_path, name, floor):
    d, b, mu, sd, pre, P, mb, names, cnt = _mapped(dataset, name, tmp_path)
    npairs = DATASETS[name][1]["pairs"]
    assert cnt["incomplete"] == 0 and len(names) == npairs and names == sorted(names) and cnt["records"] == cnt["used"] == len(mb["chain_contig"])
    nr = len(mb["read_off"]) - 1
    qlen = np.zeros(len(mb["chain_contig"]), np.int64)
    for c in range(len(qlen)):
        ops = mb["cigar"][mb["cigar_off"][c]:mb["cigar_off"][c + 1]]
        qlen[c] = sum(int(x >> 4) for x in ops if (x & 15) in (0, 1, 4))
        assert (ops[0] & 15) in (0, 4) and (ops[-1] & 15) in (0, 4) and all((x & 15) in (0, 1, 2, 4) for x in ops)
        inner = [x & 15 for x in ops if (x & 15) != 4]
        assert inner[0] == 0 and inner[-1] == 0                                   # an alignment starts and ends with matched bases
    for r in range(nr):
        c0, c1 = mb["chain_off"][r], mb["chain_off"][r + 1]
        fl = mb["chain_flag"][c0:c1]; prim = np.nonzero((fl & 0x100) == 0)[0]
        assert len(prim) == 1                                                     # exactly one primary record per mate (protoSeeds.cpp:252-314)
        assert mb["chain_as"][c0:c1].max() == mb["chain_as"][c0 + prim[0]] >= 30  # ... the best-scoring one
        assert ((fl & 0x40) != 0).all() == (r % 2 == 0) and ((fl & 0x80) != 0).all() == (r % 2 == 1)
        key = list(zip(mb["chain_contig"][c0:c1], mb["chain_pos"][c0:c1])); assert key == sorted(key)   # coordinate order like a sorted BAM
        assert (qlen[c0:c1] == mb["read_off"][r + 1] - mb["read_off"][r]).all()
    # SEQ as in the FASTQ for a forward primary, reverse-complemented otherwise: either way the generator's BAM-orientation bases when the strand is right
    same = sum(bytes(mb["bases"][mb["read_off"][2 * p]:mb["read_off"][2 * p + 1]]) == bytes(b["bases"][b["read_off"][2 * int(n[1:])]:b["read_off"][2 * int(n[1:]) + 1]]) for p, n in enumerate(names))
    assert same >= 0.99 * npairs
    O = H.Oracle(d)
    aln = H.quiet(O.pairs, mb, mu, sd, 1024)
    if H.have_ref():      # the compiled reference (a process of its own: one graph per process): the mapper's batches are new ground for the restatement too
        ref = H._ref_subprocess(d, mb, mu, sd, 1024, "pairs")
        for k in ("n_cols", "level", "edge", "schar", "mapq", "read_reverse"):
            assert np.array_equal(aln[k], ref[k]), "restatement and compiled reference differ in %s" % k
    got = _fraction(aln, pre, names)
    want = _fraction(H.quiet(O.pairs, b, mu, sd, 1024), pre, None)
    assert got >= floor and got >= want - 0.004, "bases on their true level: %.4f with the mapper's seeds, %.4f with the generator's" % (got, want)
    # a gzip-compressed copy gives the same batch
    for m in ("1", "2"):
        with open(pre + "_%s.fq" % m, "rb") as f, gzip.open(pre + "_%s.fq.gz" % m, "wb") as z:
            z.write(f.read())
    mb2, names2, _ = P.fastq_map(pre + "_1.fq.gz", pre + "_2.fq.gz", threads=2)
    assert names2 == names and all((mb2[k] == mb[k]).all() for k in H.BATCH_KEYS)
    P.close()


def test_unplaceable_reads_and_malformed_input(dataset, tmp_path):
    d, _b, _mu, _sd = dataset("small")
    P = H.Product(d)
    rng = np.random.RandomState(1)
    seq = P.array("contig_seq"); off = P.array("contig_off")
    good = bytes(seq[off[0] + 500:off[0] + 600]).decode(); mate = bytes(seq[off[0] + 700:off[0] + 800]).decode()
    comp = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}
    mate_rc = "".join(comp[c] for c in reversed(mate)); junk = "".join("ACGT"[i] for i in rng.randint(0, 4, 100))
    f1, f2 = str(tmp_path / "a_1.fq"), str(tmp_path / "a_2.fq")
    open(f1, "w").write("@p2/1 extra\n%s\n+\n%s\n@p1/1\n%s\n+p1\n%s\n" % (good, "I" * 100, junk, "I" * 100))
    open(f2, "w").write("@p2/2\n%s\n+\n%s\n@p1/2\n%s\n+\n%s\n" % (mate_rc, "5" * 100, mate_rc, "5" * 100))
    mb, names, cnt = P.fastq_map(f1, f2)
    assert names == ["p2"] and cnt["names"] == 2 and cnt["incomplete"] == 1          # the pair with a random mate has no placement for it and is dropped
    c0 = mb["chain_off"][0]; prim = [c for c in range(c0, mb["chain_off"][1]) if not mb["chain_flag"][c] & 0x100][0]
    assert mb["chain_contig"][prim] == 0 and mb["chain_pos"][prim] == 500 and mb["chain_as"][prim] == 100 and list(mb["cigar"][mb["cigar_off"][prim]:mb["cigar_off"][prim + 1]]) == [100 << 4]
    prim2 = [c for c in range(mb["chain_off"][1], mb["chain_off"][2]) if not mb["chain_flag"][c] & 0x100][0]
    assert mb["chain_flag"][prim2] & 0x10 and mb["chain_pos"][prim2] == 700 and bytes(mb["bases"][100:200]).decode() == mate and mb["chain_flag"][prim] & 0x20
    open(f2, "w").write("@p2/2\n%s\n+\n%s\n" % (mate_rc, "5" * 100))
    with pytest.raises(RuntimeError, match="different numbers of reads"):
        P.fastq_map(f1, f2)
    open(f2, "w").write("@p2/2\n%s\n+\n%s\n@px/2\n%s\n+\n%s\n" % (mate_rc, "5" * 100, mate_rc, "5" * 100))
    with pytest.raises(RuntimeError, match="names differ"):
        P.fastq_map(f1, f2)
    open(f2, "w").write("@p2/2\n%s\n+\n%s\n" % (mate_rc, "5" * 99))
    with pytest.raises(RuntimeError, match="malformed FASTQ"):
        P.fastq_map(f1, f2)
    with pytest.raises(RuntimeError, match="cannot open"):
        P.fastq_map(f1, str(tmp_path / "none.fq"))
    P.close()


@pytest.mark.gpu
def test_gpu_alignment_of_mapped_fastq_equals_the_oracle_and_keeps_the_accuracy(dataset, tmp_path):
    """the batch from the mapper through hlala_align_pairs on the GPU: identical to the oracle on the same batch (columns, mapping qualities, coverage),
    and the bases-on-true-level figure with it"""
    d, b, mu, sd, pre, P, mb, names, cnt = _mapped(dataset, "S", tmp_path)
    P.to_gpu(0)
    got = P.pairs(mb, mu, sd, 1024)
    try:
        want = H.oracle_pairs(d, mb, mu, sd, 1024)
    except RuntimeError:      # a compiled-reference graph opened outside the harness cache earlier in this process: use a process of its own
        want = H._ref_subprocess(d, mb, mu, sd, 1024, "pairs")
    from test_gpu_parity import assert_pairs_equal
    assert_pairs_equal(got, want)
    assert _fraction(got, pre, names) >= 0.985
    P.close()


@pytest.mark.gpu
def test_cli_runs_from_fastq(dataset, tmp_path):
    """hlala-b200 --action HLA --FASTQ1 --FASTQ2 (the arguments HLA-LA.pl:563 builds for short reads): reads placed here, then the usual path; the calls equal those
    of the same run from the generator's BAM where the loci are covered"""
    d, _b, _mu, _sd = dataset("typing")
    pre = str(tmp_path / "R")
    H.synth_reads(d, str(tmp_path / "seeds.bin"), levels_prefix=pre, **DATASETS["typing"][1])
    out = str(tmp_path / "out")
    r = subprocess.run([H.CLI, "--action", "HLA", "--sampleID", "S", "--FASTQ1", pre + "_1.fq", "--FASTQ2", pre + "_2.fq", "--outputDirectory", out, "--PRG_graph_dir", d,
                        "--insertSizeMean", "100", "--insertSizeSD", "10"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "reads placed on the PRG contigs" in r.stdout
    assert len(os.listdir(os.path.join(out, "hla"))) == 73 and os.path.getsize(os.path.join(out, "reads_per_level.txt")) > 0
    calls = [l.split("\t") for l in r.stdout.splitlines() if l.count("\t") == 4]
    assert len(calls) == 17 and calls[0][0] == "A"


def test_insert_size_estimate_from_mapped_fastq(tmp_path):
    """the default insert-size model of a FASTQ run: the mapper fills the sample estimateInsertSize works on (primary records of the first 2000 pairs in name order);
    alignments from the oracle restatement here, the arithmetic is hlala_insert_size_from_levels (pinned to the reference in insert_size_ref_compare.py)"""
    d = str(tmp_path / "prg"); H.synth_prg(d, levels=40000, haps=4, genes=2, alleles=48, seed=9)
    pre = str(tmp_path / "R")
    H.synth_reads(d, str(tmp_path / "seeds.bin"), pairs=2500, len=100, clip_frac=0.15, seed=9, gap_mean=180, gap_sd=25, levels_prefix=pre)
    P = H.Product(d)
    sample, loaded, _ = P.bam_insert_size(pre + "_1.fq", fastq2=pre + "_2.fq", threads=4)
    assert len(sample["read_off"]) - 1 == 4000 and (np.diff(sample["chain_off"]) == 1).all() and not (sample["chain_flag"] & 0x100).any()
    oc = H.Oracle(d).chains(sample, 640)
    assert (oc["status"] == 0).all()
    fl = np.full(len(oc["status"]), -1, np.int32); ll = np.full(len(oc["status"]), -1, np.int32)
    for i, n in enumerate(oc["n_cols"]):
        lv = oc["level"][i, :n]; lv = lv[lv != -1]
        if len(lv):
            fl[i], ll[i] = lv[0], lv[-1]
    mean, sd, used, skipped = P.insert_size_from_levels(fl, ll, ((sample["chain_flag"] & 0x10) != 0).astype(np.uint8), loaded)
    assert abs(mean - 180) <= 6 and 15 <= sd <= 35 and used >= 1900, (mean, sd, used, skipped)
    P.close()


def test_test_prg_mapping_action_arguments(dataset, tmp_path):
    """--action testPRGMapping (the reference's synthetic integration test, HLA-LA.cpp:1386-1621): argument check, and a loud failure where there is no GPU"""
    r = subprocess.run([H.CLI, "--action", "testPRGMapping", "--PRG_graph_dir", "x"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 2 and "usage: hlala-b200 --action testPRGMapping" in r.stderr
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    d, _b, _mu, _sd = dataset("small")
    r = subprocess.run([H.CLI, "--action", "testPRGMapping", "--PRG_graph_dir", d, "--outputDirectory", str(tmp_path / "o"), "--qualityMatrixFile", "none.txt"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "uploading the PRG" in r.stderr


@pytest.mark.gpu
def test_test_prg_mapping_action_on_the_gpu(dataset, tmp_path):
    """reads with known levels from a random genome of the graph -> placed on the contigs -> aligned on the GPU -> compared with their true levels: the figure the
    reference prints as "Graph: <bases> <fraction>" (HLA-LA.cpp:1258). The same flow with the oracle's alignments on the CPU gives 0.993 on this PRG."""
    import simulator_ref_compare as S
    d, _b, _mu, _sd = dataset("S")
    mat = str(tmp_path / "m.txt"); S.synthetic_matrix(mat, accurate=True)
    r = subprocess.run([H.CLI, "--action", "testPRGMapping", "--PRG_graph_dir", d, "--outputDirectory", str(tmp_path / "o"), "--qualityMatrixFile", mat, "--seed", "1"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("\tGraph: ")]
    assert len(line) == 1, r.stdout[-2000:]
    total, frac = line[0].split()[1:]
    assert int(total) > 100000 and float(frac) >= 0.985, line


def test_placements_of_constructed_reads(dataset, tmp_path):
    """reads cut from a contig with one known difference each: the placement, CIGAR and score (bwa mem's defaults: 1 / -4 / gap 6 + 1 per base / clipping 5) are the expected
    ones, the oracle and the compiled reference take the batch and agree on it"""
    d, _b, _mu, _sd = dataset("small")
    P = H.Product(d); seq = P.array("contig_seq"); off = P.array("contig_off")
    c0 = bytes(seq[off[0]:off[1]]).decode(); L0 = len(c0)
    comp = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}
    rc = lambda s: "".join(comp[c] for c in reversed(s))
    other = lambda ch: "ACGT"[("ACGT".index(ch) + 1) % 4] if ch in "ACGT" else ch
    cases = [  # name, first read, second read (sequenced from the other strand), expected (pos, score, CIGAR) of the first read's primary, of the second read's
        ("n_in_middle", c0[300:350] + "N" + c0[351:400], rc(c0[600:700]), (300, 95, "100M"), (600, 100, "100M")),
        ("lowercase", c0[1000:1100].lower(), rc(c0[1300:1400]), (1000, 100, "100M"), (1300, 100, "100M")),
        ("short_mate", c0[2000:2100], rc(c0[2300:2315]), None, None),                                       # 15 bases cannot be placed: the pair is dropped
        ("unequal", c0[2500:2600], rc(c0[2800:2860]), (2500, 100, "100M"), (2800, 60, "60M")),
        ("contig_start", "ACGTACGTAC" + c0[0:90], rc(c0[300:400]), (0, 90, "10S90M"), (300, 100, "100M")),
        ("contig_end", c0[L0 - 400:L0 - 300], rc(c0[L0 - 90:L0] + "GATTACAGAT"), (L0 - 400, 100, "100M"), (L0 - 90, 90, "90M10S")),
        ("deletion", c0[3000:3050] + c0[3056:3106], rc(c0[3400:3500]), (3000, 88, "50M6D50M"), (3400, 100, "100M")),
        ("insertion", c0[3600:3650] + "GATTA" + c0[3650:3695], rc(c0[3900:4000]), (3600, 84, "50M5I45M"), (3900, 100, "100M")),
        ("mismatches", "".join(other(ch) if i % 23 == 5 else ch for i, ch in enumerate(c0[4200:4300])), rc(c0[4500:4600]), (4200, 75, "100M"), (4500, 100, "100M")),
    ]
    f1, f2 = str(tmp_path / "a_1.fq"), str(tmp_path / "a_2.fq")
    with open(f1, "w") as a, open(f2, "w") as b:
        for n, x, y, _e1, _e2 in cases:
            a.write("@%s/1\n%s\n+\n%s\n" % (n, x, "I" * len(x))); b.write("@%s/2\n%s\n+\n%s\n" % (n, y, "I" * len(y)))
    mb, names, cnt = P.fastq_map(f1, f2)
    assert names == sorted(n for n, *_ in cases if n != "short_mate") and cnt["incomplete"] == 1
    for n, _x, _y, e1, e2 in cases:
        if e1 is None:
            continue
        p_ = names.index(n)
        for m, e in ((0, e1), (1, e2)):
            r = 2 * p_ + m
            prim = [c for c in range(mb["chain_off"][r], mb["chain_off"][r + 1]) if not mb["chain_flag"][c] & 0x100][0]
            cig = "".join("%d%s" % (x >> 4, "MIDNSHP=X"[x & 15]) for x in mb["cigar"][mb["cigar_off"][prim]:mb["cigar_off"][prim + 1]])
            assert (int(mb["chain_pos"][prim]), int(mb["chain_as"][prim]), cig) == e and mb["chain_contig"][prim] == 0, (n, m, mb["chain_pos"][prim], mb["chain_as"][prim], cig)
            assert bool(mb["chain_flag"][prim] & 0x10) == (m == 1)
    assert not (mb["bases"] >= ord("a")).any()
    rest = H.quiet(H.Oracle(d).pairs, mb, 200.0, 50.0, 640)
    assert (rest["n_cols"] >= np.diff(mb["read_off"])).all()
    if H.have_ref():
        want = H._ref_subprocess(d, mb, 200.0, 50.0, 640, "pairs")
        for k in ("n_cols", "level", "schar", "mapq"):
            assert np.array_equal(want[k], rest[k]), k
    P.close()


@pytest.mark.gpu
def test_last_bit_fixture_on_the_gpu(tmp_path):
    """with CUDA's own exp() the 110 columns came out as phred 190 (profiles/r02_gpu_tests_before_exp_fix.log)"""
    from test_exp_libm import _fixture, _same_as_gold
    d, gold, b = _fixture(tmp_path)
    P = H.Product(d); P.to_gpu(0)
    _same_as_gold(P.pairs(b, 100.0, 10.0, 640, want_levels=False), gold)
    P.close()


def test_mate_rescue(dataset, tmp_path):
    """bwa mem's mate rescue: a mate without a 19-mer of its own (a mismatch every 15 bases) is aligned, on the opposite strand, within 800 bases of its mate's primary
    placement; a mate that matches nothing there stays unplaced and the pair is dropped"""
    d, _b, _mu, _sd = dataset("small")
    P = H.Product(d); seq = P.array("contig_seq"); off = P.array("contig_off")
    c0 = bytes(seq[off[0]:off[1]]).decode()
    comp = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}
    rc = lambda s: "".join(comp[c] for c in reversed(s))
    other = lambda ch: "ACGT"[("ACGT".index(ch) + 2) % 4] if ch in "ACGT" else ch
    noisy = "".join(other(ch) if i % 15 == 7 else ch for i, ch in enumerate(c0[1500:1600]))      # 7 mismatches: score 93 - 28 = 65
    rng = np.random.RandomState(3); junk = "".join("ACGT"[i] for i in rng.randint(0, 4, 100))
    f1, f2 = str(tmp_path / "a_1.fq"), str(tmp_path / "a_2.fq")
    open(f1, "w").write("@down\n%s\n+\n%s\n@up\n%s\n+\n%s\n@none\n%s\n+\n%s\n" % (c0[1200:1300], "I" * 100, rc(noisy), "I" * 100, c0[3000:3100], "I" * 100))
    open(f2, "w").write("@down\n%s\n+\n%s\n@up\n%s\n+\n%s\n@none\n%s\n+\n%s\n" % (rc(noisy), "I" * 100, c0[1200:1300], "I" * 100, junk, "I" * 100))
    mb, names, cnt = P.fastq_map(f1, f2)
    assert names == ["down", "up"] and cnt["incomplete"] == 1 and cnt["is_n"] == 2          # two mates placed by rescue
    for p_, (r_noisy, r_clean) in enumerate(((1, 0), (0, 1))):
        cn, cc = mb["chain_off"][2 * p_ + r_noisy], mb["chain_off"][2 * p_ + r_clean]
        assert mb["chain_off"][2 * p_ + r_noisy + 1] - cn == 1
        assert mb["chain_pos"][cn] == 1500 and mb["chain_as"][cn] == 65 and list(mb["cigar"][mb["cigar_off"][cn]:mb["cigar_off"][cn + 1]]) == [100 << 4]
        assert mb["chain_flag"][cn] & 0x10 and not mb["chain_flag"][cn] & 0x100 and not mb["chain_flag"][cc] & 0x10 and mb["chain_pos"][cc] == 1200
        s = mb["read_off"][2 * p_ + r_noisy]; assert bytes(mb["bases"][s:s + 100]).decode() == noisy       # stored in the orientation of its placement
    aln = H.quiet(H.Oracle(d).pairs, mb, 200.0, 50.0, 640)
    assert (aln["n_cols"] >= 100).all()
    P.close()
