"""HLA typing from long reads (HLA-LA.pl --longReads; hla::HLATyper::HLATypeInference with unpaired reads, hla/HLATyper.cpp:935-946, 1079-1097, 1467-1495, 1797-1880,
1918, 3568-3930): every file the UNMODIFIED reference writes into outDir/hla must be byte-identical — from the product's host logic (CPU, kernels replaced by the
test-only loops) and from the whole GPU path through the C ABI."""
import json
import os
import subprocess
import sys

import pytest

import harness as H

pytestmark = pytest.mark.skipif(not H.have_ref(), reason="oracle/_ref/libhlala_ref.so not built (reference tree absent)")


def compare(dataset, tmp_path, name, mode):
    d, b, mu, sd = dataset(name)
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, os.path.join(here, "typing_long_ref_compare.py"), d, os.path.join(d, "seeds.bin"), str(tmp_path), mode, "4096"],
                       cwd=here, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=1200)
    assert r.returncode == 0, r.stderr[-3000:]
    v = json.loads(r.stdout.strip().split("\n")[-1])
    assert v["files_ref"] == v["files_mine"] == 5 + 4 * 17 and v["n_used"] == v["n_selected"] and v["n_used"] > 300
    assert not v["differing"], "files differ from the reference's: %s" % v["differing"]
    return v


@pytest.mark.parametrize("name", ["long_typing", "long_typing_deep"])
def test_host_logic_writes_the_references_files(dataset, tmp_path, name):
    v = compare(dataset, tmp_path, name, "host")
    if name == "long_typing_deep":
        assert v["deepest_pileup_column"] >= 200, "the deep dataset must reach the strand filter's coverage (100 observations of one allele)"


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["long_typing", "long_typing_deep"])
def test_gpu_path_writes_the_references_files(dataset, tmp_path, name):
    compare(dataset, tmp_path, name, "gpu")


@pytest.mark.gpu
def test_command_line_long_reads(dataset, tmp_path):
    """hlala-b200 --action HLA --longReads ont2d on a BAM of single long reads that also holds their secondary records: the ingest keeps the primary records only
    (extractSeeds2 in long-read mode), so the reference and the library are run on that primary-only batch. Every file without read names is byte-identical to the
    library path's (which is compared with the reference's files), the calls are the reference's, reads_per_level.txt is the per-level count of the alignments."""
    import filecmp
    import numpy as np
    d, b, mu, sd = dataset("long_typing")
    bam = str(tmp_path / "long.bam")
    subprocess.run([H.SYNTH, "bam", "--prg", d, "--seeds", os.path.join(d, "seeds.bin"), "--out", bam, "--single", "1"], check=True, stderr=subprocess.DEVNULL)
    out = str(tmp_path / "out")
    exe = os.path.join(H.PKG, "build", "hlala-b200")
    r = subprocess.run([exe, "--action", "HLA", "--sampleID", "S", "--BAM", bam, "--outputDirectory", out, "--PRG_graph_dir", d, "--longReads", "ont2d", "--maxColumns", "4096"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:] + r.stdout[-2000:]
    assert len(os.listdir(os.path.join(out, "hla"))) == 5 + 4 * 17
    # the batch the ingest hands on: primary records only
    prim = {k: b[k] for k in H.BATCH_KEYS}
    keep = (b["chain_flag"] & 0x100) == 0; idx = np.nonzero(keep)[0]
    prim["chain_off"] = np.concatenate([[0], np.cumsum(np.add.reduceat(keep.astype(np.int64), b["chain_off"][:-1]))]).astype(np.int32)
    for k in ("chain_contig", "chain_pos", "chain_flag", "chain_as"):
        prim[k] = np.ascontiguousarray(b[k][idx])
    co = [0]; cg = []
    for c in idx:
        cg.append(b["cigar"][b["cigar_off"][c]:b["cigar_off"][c + 1]]); co.append(co[-1] + len(cg[-1]))
    prim["cigar_off"] = np.array(co, np.int32); prim["cigar"] = np.ascontiguousarray(np.concatenate(cg))
    assert (np.diff(prim["chain_off"]) == 1).all() and len(idx) < len(keep)
    seeds = str(tmp_path / "prim.npz"); np.savez(seeds, **prim)
    here = os.path.dirname(os.path.abspath(__file__))
    rr = subprocess.run([sys.executable, os.path.join(here, "typing_long_ref_compare.py"), d, seeds, str(tmp_path / "cmp"), "gpu", "4096"],
                        cwd=here, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=1200)
    assert rr.returncode == 0, rr.stderr[-3000:]
    v = json.loads(rr.stdout.strip().split("\n")[-1]); assert not v["differing"], v
    lib_dir = str(tmp_path / "cmp" / "gpu" / "hla"); cli_dir = os.path.join(out, "hla")
    nameless = [f for f in sorted(os.listdir(lib_dir)) if not f.startswith(("R1_pileup_", "R1_readIDs_"))]
    bad = [f for f in nameless if not filecmp.cmp(os.path.join(lib_dir, f), os.path.join(cli_dir, f), shallow=False)]
    def first_diff(f):
        a = open(os.path.join(lib_dir, f)).read().split("\n"); c = open(os.path.join(cli_dir, f)).read().split("\n")
        for i, (x, y) in enumerate(zip(a, c)):
            if x != y:
                return "%s line %d: %r vs %r (%d / %d lines)" % (f, i, x[:120], y[:120], len(a), len(c))
        return "%s: %d vs %d lines" % (f, len(a), len(c))
    assert not bad, [first_diff(f) for f in bad[:3]]
    ref_best = open(str(tmp_path / "cmp" / "ref" / "hla" / "R1_bestguess.txt")).read(); assert open(os.path.join(cli_dir, "R1_bestguess.txt")).read() == ref_best, "calls differ from the reference's"
    P = H.Product(d); P.to_gpu(0)
    got = P.long_reads(prim, 4096)
    cov = [int(x.split("\t")[2]) for x in open(os.path.join(out, "reads_per_level.txt")).read().splitlines()]
    assert np.array_equal(np.array(cov), got["bases_per_level"]), "reads_per_level.txt"
    P.close()
