"""A/B of the bench's long-read stage over several builds of the library (HLALA_B200_LIB is read by tests/harness.py at import): run once per build in a subprocess."""
import json, os, subprocess, sys, time
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--one":
    sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
    import argparse, bench, harness as H
    args = argparse.Namespace(long_reads=20000, long_ref_reads=0, levels=5000000, genes=17, alleles=1000, haps=8, cpu_levels=294118)
    root = "/tmp/hlala_long_ab"; os.makedirs(root, exist_ok=True)
    t = time.time(); r = bench.stage_long_reads(args, root, H); r["stage_s"] = time.time() - t
    print(json.dumps({k: r[k] for k in ("call_s", "reads_per_s_e2e", "sum_columns", "sum_ll", "stage_s") if k in r}))
else:
    for lib in sys.argv[1:]:
        env = dict(os.environ); env["HLALA_B200_LIB"] = os.path.join(REPO, "ab", lib)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--one"], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        print(lib, r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-500:])
