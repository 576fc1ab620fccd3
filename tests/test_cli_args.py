"""Argument surface of hlala-b200 (no GPU needed): what the reference binary answers for --action testBinary (HLA-LA.cpp:129-132) and a missing --action (:104-108),
and this program's own rules — unknown keys and odd argument counts are errors, paired FASTQ input is mapped by the program itself while unpaired FASTQ input is answered with the mapping step to run first, --longReads takes
0 / ont2d / pacbio (HLA-LA.cpp:759)."""
import os
import subprocess

import harness as H

EXE = os.path.join(H.PKG, "build", "hlala-b200")


def run(*args):
    return subprocess.run([EXE] + list(args), stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=60)


def test_test_binary():
    r = run("--action", "testBinary")
    assert r.returncode == 0 and r.stdout == "\nHLA*LA binary functional!\n\n"


def test_missing_action_and_bad_arguments():
    r = run("--sampleID", "x")
    assert r.returncode != 0 and "Missing --action parameter" in r.stderr
    r = run("--action", "HLA", "--bogus", "1")
    assert r.returncode == 2 and "unknown argument --bogus" in r.stderr
    r = run("--action", "HLA", "--BAM")
    assert r.returncode == 2 and "has no value" in r.stderr
    r = run("--action", "HLA")
    assert r.returncode == 2 and "usage:" in r.stderr


def test_fastq_and_long_read_arguments(tmp_path):
    # paired FASTQ files (what HLA-LA.pl:563 passes) are accepted and mapped by the program itself: the run gets as far as the (missing) PRG
    r = run("--action", "HLA", "--FASTQ1", "a.fq", "--FASTQ2", "b.fq", "--outputDirectory", str(tmp_path), "--PRG_graph_dir", str(tmp_path / "nope"))
    assert r.returncode == 1 and "loading the PRG" in r.stderr
    # unpaired / long-read FASTQ input still needs the external mapping step
    r = run("--action", "HLA", "--FASTQU", "a.fq", "--outputDirectory", str(tmp_path), "--PRG_graph_dir", str(tmp_path))
    assert r.returncode == 2 and "bwa mem -a -M" in r.stderr
    r = run("--action", "HLA", "--FASTQ1", "a.fq", "--outputDirectory", str(tmp_path), "--PRG_graph_dir", str(tmp_path))
    assert r.returncode == 2 and "bwa mem -a -M" in r.stderr
    r = run("--action", "HLA", "--FASTQ1", "a.fq", "--FASTQ2", "b.fq", "--longReads", "ont2d", "--outputDirectory", str(tmp_path), "--PRG_graph_dir", str(tmp_path))
    assert r.returncode == 2 and "--longReads takes a --BAM" in r.stderr
    r = run("--action", "HLA", "--BAM", "x.bam", "--outputDirectory", str(tmp_path), "--PRG_graph_dir", str(tmp_path), "--longReads", "nanopore")
    assert r.returncode == 2 and "ont2d or pacbio" in r.stderr
    # --longReads 0 is the short-read path (HLA-LA.cpp:760-763); --maxThreads is HLA-LA.pl's name for the thread count: both are accepted, the run then fails at the missing PRG
    r = run("--action", "HLA", "--BAM", "x.bam", "--outputDirectory", str(tmp_path), "--PRG_graph_dir", str(tmp_path / "nope"), "--longReads", "0", "--maxThreads", "2")
    assert r.returncode == 1 and "loading the PRG" in r.stderr
