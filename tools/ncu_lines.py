#!/usr/bin/env python
"""Join an ncu SASS source page with nvdisasm line info: stall samples and executed instructions per CUDA source line.

usage: ncu_lines.py <report.ncu-rep> <library.so> <kernel-substring> [top-N]
The .so must be the build the report was taken from (compile with -lineinfo)."""
import csv, io, os, re, subprocess, sys, tempfile, collections


def main():
    rep, so, kern = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    # the CSV holds one block per kernel: "Kernel Name" row, header row, instruction rows
    blocks = []; i = 0
    while i < len(rows):
        if rows[i] and rows[i][0] == "Kernel Name":
            name = rows[i][1]; hdr = rows[i + 1]; j = i + 2
            while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
                j += 1
            blocks.append((name, hdr, rows[i + 2:j])); i = j
        else:
            i += 1
    blk = [b for b in blocks if kern in b[0]]
    if not blk:
        print("kernels in report:", [b[0][:80] for b in blocks]); return 1
    name, hdr, data = blk[0]
    ci = {h: k for k, h in enumerate(hdr)}
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=d, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    lines = None
    for f in os.listdir(d):
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, f)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
        if "k_" not in txt:
            continue
        # split per function
        parts = re.split(r"\n//-+ \.text\.(\S+) -+\n", txt)
        for k in range(1, len(parts), 2):
            fn, body = parts[k], parts[k + 1]
            dem = subprocess.run(["c++filt", fn], stdout=subprocess.PIPE, text=True).stdout
            if kern in dem and ("<" not in name or dem.split("(")[0].replace(" ", "")[:60] == name.split("(")[0].replace("void ", "").replace(" ", "").replace("(int)", "")[:60] or True):
                cur = None; seq = []
                for ln in body.split("\n"):
                    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
                    if m:
                        cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
                    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
                        seq.append(cur)
                if lines is None or abs(len(seq) - len(data)) < abs(len(lines) - len(data)):
                    lines = seq
    if lines is None:
        print("kernel not found in", so); return 1
    print("kernel:", name[:110]); print("SASS instructions: report %d, disassembly %d" % (len(data), len(lines)))
    n = min(len(data), len(lines))
    agg = collections.defaultdict(lambda: [0.0, 0.0, collections.Counter(), 0.0])
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot_s = tot_i = 0.0
    for k in range(n):
        r = data[k]
        s = float(r[ci["# Samples"]] or 0); ins = float(r[ci["Instructions Executed"]] or 0)
        a = agg[lines[k]]; a[0] += s; a[1] += ins; a[3] += float(r[ci["Thread Instructions Executed"]] or 0) if "Thread Instructions Executed" in ci else 0.0
        for h in stall_cols:
            v = float(r[ci[h]] or 0)
            if v:
                a[2][h] += v
        tot_s += s; tot_i += ins
    src_cache = {}
    def src(loc):
        if not loc:
            return ""
        f, l = loc
        if f not in src_cache:
            for root in ("hla-la_b200/csrc", "hla-la_b200/host"):
                p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), root, f)
                if os.path.exists(p):
                    src_cache[f] = open(p).read().split("\n"); break
            else:
                src_cache[f] = []
        L = src_cache[f]
        return L[l - 1].strip()[:110] if 0 < l <= len(L) else ""
    print("total samples %.0f, total warp instructions %.3g" % (tot_s, tot_i))
    print("%6s %6s %5s  %-26s %-40s %s" % ("smp%", "ins%", "lanes", "location", "top stalls", "source"))
    for loc, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        st = ", ".join("%s %.0f%%" % (h.replace("stall_", ""), 100 * v / max(a[0], 1)) for h, v in a[2].most_common(3))
        print("%6.2f %6.2f %5.1f  %-26s %-40s %s" % (100 * a[0] / max(tot_s, 1), 100 * a[1] / max(tot_i, 1), a[3] / max(a[1], 1), "%s:%d" % loc if loc else "?", st, src(loc)))
    return 0


if __name__ == "__main__":
    sys.exit(main())
