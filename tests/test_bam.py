"""BAM ingest: hlala_bam_read against the rules of processBAM::extractSeeds2 / protoSeeds (mapper/processBAM.cpp:703-862,
mapper/reads/protoSeeds.cpp:24,377) and the command-line front end. BamTools is not part of the reference tree, so there is no reference binary to
pin BamTools' byte-level record decoding against (round-trip tests). What the reference itself decides on decoded records (which are kept, how they
are grouped and ordered, which pairs are complete) is pinned against its unmodified extractSeeds2 / protoSeeds, fed the same records through the
in-memory BamReader stand-in."""
import os
import subprocess

import numpy as np
import pytest

import harness as H


def test_round_trip_of_a_seed_batch(dataset, tmp_path):
    d, b, _mu, _sd = dataset("S")
    bam = str(tmp_path / "s.bam")
    H.synth_bam(d, os.path.join(d, "seeds.bin"), bam)
    P = H.Product(d)
    got, names, st = P.bam_read(bam, threads=4)
    n_pairs = (len(b["read_off"]) - 1) // 2
    assert names == ["r%09d" % p for p in range(n_pairs)] and st["incomplete"] == 0 and st["used"] == len(b["chain_contig"]) == st["records"]
    for k in ("read_off", "bases", "quals", "chain_off", "chain_contig", "chain_pos", "chain_as", "cigar_off", "cigar"):
        assert np.array_equal(got[k], b[k]), k
    assert np.array_equal(got["chain_flag"] & 0x110, b["chain_flag"] & 0x110)
    mate = np.repeat(np.arange(len(b["read_off"]) - 1) & 1, np.diff(b["chain_off"]))
    assert np.array_equal((got["chain_flag"] & 0xC0) == 0x80, mate == 1) and (got["chain_flag"] & 1).all()
    assert st["is_n"] > 100 and 80 < st["is_mean"] < 120     # the generator draws gaps from N(100, 10)
    one, _names1, st1 = P.bam_read(bam, threads=1)
    assert all(np.array_equal(one[k], got[k]) for k in got) and st1 == st      # the thread count changes nothing, insert-size estimate included
    P.close()


def test_record_selection_rules(dataset, tmp_path):
    d, _b, _mu, _sd = dataset("small")
    P = H.Product(d)
    nc = P.dims()["n_contigs"]; off = P.array("contig_off"); lens = np.diff(off)
    refs = [("PRG_%d" % i, int(lens[j])) for j, i in enumerate(P.array("contig_prg_id"))] + [("chrUn", 5000)]
    seq = "ACGTTGCAAC" * 5; q = bytes([30] * 50); M50 = [("M", 50)]
    R = []

    def rec(name, flag, ref, pos, cigar=M50, s=seq, as_=40, **kw):
        R.append(dict(name=name, flag=flag, ref=ref, pos=pos, cigar=cigar, seq=s, qual=q[:len(s)], tags={b"AS": as_} if as_ is not None else {b"NM": 0}, **kw))
    # pair "b": complete, the second mate also has a secondary record with a better score and SEQ '*'
    rec("b", 0x1 | 0x40, 0, 100); rec("b", 0x1 | 0x80 | 0x10, 0, 300, as_=38); rec("b", 0x1 | 0x80 | 0x10 | 0x100, 1, 10, s="", as_=45)
    # pair "a": sorts before "b" although it comes later in the file
    rec("a", 0x1 | 0x80, 1, 50); rec("a", 0x1 | 0x40 | 0x10, 1, 220, cigar=[("S", 5), ("M", 40), ("I", 2), ("M", 3)])
    # pair "c": first mate only has a secondary record -> incomplete
    rec("c", 0x1 | 0x40 | 0x100, 0, 10, s=""); rec("c", 0x1 | 0x80, 0, 200)
    # pair "d": one mate unmapped, the other on a contig that is not part of the PRG -> no record kept
    rec("d", 0x1 | 0x40 | 0x4, 0, 10, cigar=[]); rec("d", 0x1 | 0x80, nc, 100)
    # pair "e": a record hanging over the end of its contig is not inside the interval; the other mate alone is incomplete
    rec("e", 0x1 | 0x40, 0, refs[0][1] - 20); rec("e", 0x1 | 0x80, 0, 500)
    bam = str(tmp_path / "rules.bam"); H.write_bam(bam, refs, R, block=300)
    got, names, st = P.bam_read(bam)
    assert names == ["a", "b"] and st["records"] == len(R) and st["names"] == 4 and st["incomplete"] == 2
    assert list(got["chain_off"]) == [0, 1, 2, 3, 5] and list(got["read_off"]) == [0, 50, 100, 150, 200]
    assert list(got["chain_contig"]) == [1, 1, 0, 0, 1] and list(got["chain_pos"]) == [220, 50, 100, 300, 10] and list(got["chain_as"]) == [40, 40, 40, 38, 45]
    assert list(got["chain_flag"] & 0x110) == [0x10, 0, 0, 0x10, 0x110]
    assert bytes(got["bases"][150:200]).decode() == seq and (got["quals"] == 30 + 33).all()     # the primary's SEQ stands for the read, not the better-scoring secondary
    assert list(got["cigar"][got["cigar_off"][0]:got["cigar_off"][1]]) == [(5 << 4) | 4, (40 << 4) | 0, (2 << 4) | 1, (3 << 4) | 0]
    # errors are loud
    R.append(dict(name="z", flag=0x1 | 0x40, ref=0, pos=5, cigar=M50, seq=seq, qual=q, tags={b"NM": 1}))
    H.write_bam(bam, refs, R)
    with pytest.raises(RuntimeError, match="AS tag"):
        P.bam_read(bam)
    R[-1] = dict(name="z", flag=0x40, ref=0, pos=5, cigar=M50, seq=seq, qual=q, tags={b"AS": 1})
    H.write_bam(bam, refs, R)
    with pytest.raises(RuntimeError, match="unpaired"):
        P.bam_read(bam)
    open(bam, "wb").write(open(bam, "rb").read()[:200])
    with pytest.raises(RuntimeError, match="BGZF|BAM"):
        P.bam_read(bam)
    with pytest.raises(RuntimeError, match="Cannot open"):
        P.bam_read(str(tmp_path / "missing.bam"))
    P.close()


@pytest.mark.skipif(not H.have_ref(), reason="compiled reference (oracle/_ref) not built on this box")
def test_ingest_matches_the_references_extractSeeds2():
    """selection, grouping, name order and completeness against the unmodified processBAM::extractSeeds2 + protoSeeds (own process)"""
    import sys
    r = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "bam_ref_compare.py")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0 and "ok:" in r.stdout, r.stdout[-3000:]


@pytest.mark.skipif(not H.have_ref(), reason="compiled reference (oracle/_ref) not built on this box")
def test_long_read_ingest_matches_the_references_extractSeeds2():
    """long-read mode (primary records only, unpaired, supplementary records count as primary, one read per name) against extractSeeds2(.., longReadMode) + isComplete_unpaired"""
    import sys
    r = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "bam_long_ref_compare.py")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0 and "ok:" in r.stdout, r.stdout[-3000:]


@pytest.mark.skipif(not H.have_ref(), reason="compiled reference (oracle/_ref) not built on this box")
def test_insert_size_estimate_matches_the_references_estimateInsertSize():
    """sample selection (incl. the reference's habit of reading on after its thresholds are met), lazily loaded translations, strand rule, underlying-sequence
    distances and histogram statistics against the unmodified processBAM::estimateInsertSize (own process); alignments from the oracle restatement"""
    import sys
    r = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "insert_size_ref_compare.py")], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0 and "ok:" in r.stdout, r.stdout[-3000:]


@pytest.mark.gpu
@pytest.mark.skipif(not H.have_ref(), reason="compiled reference (oracle/_ref) not built on this box")
def test_insert_size_estimate_on_the_gpu_matches_the_reference():
    """hlala_bam_insert_size (what the command line uses when no --insertSizeMean/--insertSizeSD is given): the sample's primary records through the CUDA chain
    kernels and the extension DP, equal to the unmodified processBAM::estimateInsertSize"""
    import sys
    r = subprocess.run([sys.executable, os.path.join(os.path.dirname(os.path.abspath(__file__)), "insert_size_ref_compare.py"), "--gpu"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0 and "GPU estimate identical" in r.stdout, r.stdout[-3000:]


def test_cli_prepare_graph(dataset):
    """--action prepareGraph (HLA-LA.cpp:1341): needs no GPU; builds the flat-array cache next to graph.txt"""
    d, _b, _mu, _sd = dataset("small")
    cache = os.path.join(d, "PRG", "graph.hlala_b200.cache")
    if os.path.exists(cache):
        os.remove(cache)
    r = subprocess.run([H.CLI, "--action", "prepareGraph", "--PRG_graph_dir", d], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0 and os.path.exists(cache), r.stderr
    P = H.Product(d); dm = P.dims(); P.close()
    assert "%d levels, %d nodes, %d edges, %d gap-edge paths" % (dm["n_levels"], dm["n_nodes"], dm["n_edges"], dm["n_paths"]) in r.stdout


def test_cli_usage_and_loud_failure_without_gpu(dataset, tmp_path):
    r = subprocess.run([H.CLI, "--action", "HLA"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 2 and "usage: hlala-b200 --action HLA" in r.stderr
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    d, _b, _mu, _sd = dataset("small")
    r = subprocess.run([H.CLI, "--action", "HLA", "--sampleID", "x", "--BAM", "none.bam", "--outputDirectory", str(tmp_path / "o"), "--PRG_graph_dir", d], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_cli_writes_what_the_api_path_writes(dataset, tmp_path):
    """remapped BAM -> hlala-b200 --action HLA -> reads_per_level.txt and hla/*: identical to driving the C ABI by hand on the ingested batch"""
    import ctypes as C
    import filecmp
    d, b, mu, sd = dataset("typing")
    bam = str(tmp_path / "t.bam"); H.synth_bam(d, os.path.join(d, "seeds.bin"), bam)
    out = str(tmp_path / "cli")
    r = subprocess.run([H.CLI, "--action", "HLA", "--sampleID", "S1", "--BAM", bam, "--outputDirectory", out, "--PRG_graph_dir", d, "--insertSizeMean", str(mu), "--insertSizeSD", str(sd)],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    P = H.Product(d); P.to_gpu(0)
    ing, names, _st = P.bam_read(bam)
    res = P.pairs(ing, mu, sd, 640)
    lines = open(os.path.join(out, "reads_per_level.txt")).read().splitlines()
    assert len(lines) == P.dims()["n_levels"] - 1
    assert [int(x.split("\t")[2]) for x in lines] == list(res["bases_per_level"]) and lines[5].split("\t")[0] == "5"
    T = H.ProductTyping(P, d)
    sess = H.session_align(P, ing, mu, sd, 640, keep_columns=True)
    blob, n_sel = T.extract(sess, names=names)
    api = str(tmp_path / "api" / "hla"); T.infer([blob], mu, sd, api, keep_read_ll=False)
    P.lib.hlala_session_free(sess)
    files = sorted(os.listdir(api)); assert files == sorted(os.listdir(os.path.join(out, "hla"))) and len(files) == 73
    assert not [f for f in files if not filecmp.cmp(os.path.join(api, f), os.path.join(out, "hla", f), shallow=False)]
    calls = [l.split("\t") for l in r.stdout.splitlines() if l.count("\t") == 4]
    assert len(calls) == 17 and calls[0][0] == "A"
    T.close(); P.close()
