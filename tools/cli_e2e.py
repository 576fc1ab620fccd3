#!/usr/bin/env python
"""End-to-end run of the command line on a synthetic remapped BAM (GPU box): PRG -> seed batch -> BAM (generator), then
hlala-b200 --action HLA, with the wall time of every phase as the binary prints it. One JSON line.
usage: cli_e2e.py [--pairs N] [--levels N] [--alleles A]"""
import argparse, json, os, re, subprocess, sys, tempfile, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "tests"))
import harness as H


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=400000); ap.add_argument("--levels", type=int, default=1000000); ap.add_argument("--alleles", type=int, default=200)
    a = ap.parse_args()
    d = "/tmp/hlala_cli_e2e_l%d_a%d" % (a.levels, a.alleles)
    if not os.path.exists(d + "/.complete"):
        os.makedirs(d, exist_ok=True); H.synth_prg(d, levels=a.levels, haps=8, genes=17, alleles=a.alleles, allele_contigs=4, seed=0xB200); open(d + "/.complete", "w").write("ok")
    seeds = os.path.join(d, "seeds_cli_%d.bin" % a.pairs); bam = os.path.join(d, "remapped_%d.bam" % a.pairs)
    b = H.synth_reads(d, seeds, pairs=a.pairs, len=150, seed=0xB200, clip_frac=0.15, gene_frac=0.05)
    H.synth_bam(d, seeds, bam)
    out = tempfile.mkdtemp(prefix="hlala_cli_out_")
    t = time.time()
    r = subprocess.run([H.CLI, "--action", "HLA", "--sampleID", "S", "--BAM", bam, "--outputDirectory", out, "--PRG_graph_dir", d], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    wall = time.time() - t
    phases = {m.group(2): float(m.group(1)) for m in re.finditer(r"\[\s*([0-9.]+) s\] (.*)", r.stdout)}
    info = [l for l in r.stdout.splitlines() if "records" in l or "overlap" in l]
    print(json.dumps(dict(tool="cli_e2e", rc=r.returncode, pairs=a.pairs, bam_records=int(len(b["chain_contig"])), bam_mb=os.path.getsize(bam) / 1e6, levels=a.levels, wall_s=wall, phases_s=phases,
                          pairs_per_s_whole_run=a.pairs / wall, hla_files=len(os.listdir(os.path.join(out, "hla"))) if os.path.isdir(os.path.join(out, "hla")) else 0, info=info, stderr=r.stderr[-500:])))
    return r.returncode


if __name__ == "__main__":
    sys.exit(main())
