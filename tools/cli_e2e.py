#!/usr/bin/env python
"""End-to-end run of the command line on a synthetic remapped BAM (GPU box): PRG -> seed batch -> BAM (generator), then
hlala-b200 --action HLA, with the wall time of every phase as the binary prints it. One JSON line.
usage: cli_e2e.py [--pairs N] [--levels N] [--alleles A] [--gpus N]
With --gpus N > 1 the run is repeated with `--gpus N` (N host threads, NCCL all-reduces) and its files are compared with the one-GPU run's."""
import argparse, json, os, re, subprocess, sys, tempfile, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "tests"))
import harness as H


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=400000); ap.add_argument("--levels", type=int, default=1000000); ap.add_argument("--alleles", type=int, default=200); ap.add_argument("--gpus", type=int, default=1)
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
    multi = None
    if a.gpus > 1 and r.returncode == 0:
        import filecmp
        out2 = tempfile.mkdtemp(prefix="hlala_cli_out_multi_")
        t = time.time()
        r2 = subprocess.run([H.CLI, "--action", "HLA", "--sampleID", "S", "--BAM", bam, "--outputDirectory", out2, "--PRG_graph_dir", d, "--gpus", str(a.gpus)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        wall2 = time.time() - t
        files = sorted(os.listdir(os.path.join(out, "hla"))) if r2.returncode == 0 else []
        differ = [f for f in files if not (os.path.exists(os.path.join(out2, "hla", f)) and filecmp.cmp(os.path.join(out, "hla", f), os.path.join(out2, "hla", f), shallow=False))]
        cov_same = r2.returncode == 0 and filecmp.cmp(os.path.join(out, "reads_per_level.txt"), os.path.join(out2, "reads_per_level.txt"), shallow=False)
        calls = lambda txt: [l for l in txt.splitlines() if l.count("\t") == 4 and not l.startswith("hlala")]
        multi = dict(gpus=a.gpus, rc=r2.returncode, wall_s=wall2, phases_s={m.group(2): float(m.group(1)) for m in re.finditer(r"\[\s*([0-9.]+) s\] (.*)", r2.stdout)},
                     reads_per_level_identical=bool(cov_same), hla_files=len(files), files_differing=differ, only_pair_tables_differ=all(f.startswith("R1_PP_") for f in differ),
                     calls_identical=calls(r.stdout) == calls(r2.stdout), stderr=r2.stderr[-500:])
    print(json.dumps(dict(tool="cli_e2e", rc=r.returncode, multi_gpu=multi, pairs=a.pairs, bam_records=int(len(b["chain_contig"])), bam_mb=os.path.getsize(bam) / 1e6, levels=a.levels, wall_s=wall, phases_s=phases,
                          pairs_per_s_whole_run=a.pairs / wall, hla_files=len(os.listdir(os.path.join(out, "hla"))) if os.path.isdir(os.path.join(out, "hla")) else 0, info=info, stderr=r.stderr[-500:])))
    return r.returncode if not multi else (r.returncode or multi["rc"] or (0 if multi["reads_per_level_identical"] and multi["calls_identical"] and multi["only_pair_tables_differ"] else 1))


if __name__ == "__main__":
    sys.exit(main())
