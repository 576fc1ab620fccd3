#!/usr/bin/env python
"""bench.py — paired reads/sec of the read-to-PRG alignment path (BASELINE.json metric) on N B200s of one node.

One "step" = one pass of the whole hot path (chain kernel -> extension DP -> chain finish -> pair kernel) over one batch of
synthetic read pairs with their bwa-style seed chains, against a synthetic PRG in the reference's on-disk format.
  value        pairs/s with the batch already resident in HBM (kernels only, CUDA-event timed, max over ranks)
  e2e          pairs/s through the host-buffer C-ABI call hlala_align_pairs: chain sorting on the host, H2D of the batch from
               pinned memory, all kernels, D2H of the per-pair results and the per-level coverage
  roofline     algorithmic bytes of the dominant kernel / its CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline the unmodified reference C++ (oracle/_ref) timed on this box's host cores on a bounded sample
`--impl reference` times the reference's own CPU implementation of the path on the same workload definition.
Weak scaling: every rank aligns its own batch of --pairs pairs (reads shard by pair); the only exchange is the all-reduce of the
per-level coverage histogram and of the pair counters at the end of a step (NCCL).
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(REPO, "tests"))

import numpy as np  # noqa: E402


def rank_info():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def ensure_built():
    import harness as H
    need = [H.LIB_PRODUCT, H.SYNTH]
    if not all(os.path.exists(p) for p in need):
        sys.path.insert(0, REPO)
        import __graft_entry__
        __graft_entry__.build()


def workload_dirs(args):
    key = hashlib.sha1(("prg-v3|%d|%d|%d|%d" % (args.levels, args.haps, args.genes, args.alleles)).encode()).hexdigest()[:12]
    root = os.environ.get("HLALA_BENCH_CACHE", "/tmp/hlala_bench_cache")
    return os.path.join(root, "prg_" + key), root


def make_prg(args, prg_dir):
    import harness as H
    done = os.path.join(prg_dir, ".complete")
    if not os.path.exists(done):
        os.makedirs(prg_dir, exist_ok=True)
        H.synth_prg(prg_dir, levels=args.levels, haps=args.haps, genes=args.genes, alleles=args.alleles, allele_contigs=4, seed=0xB200)
        open(done, "w").write("ok\n")


def make_reads(args, prg_dir, rank, n_pairs, tag="bench"):
    import harness as H
    path = os.path.join(prg_dir, "seeds_%s_r%d_p%d_l%d.bin" % (tag, rank, n_pairs, args.read_len))
    if os.path.exists(path + ".ok"):
        return H.read_arrayfile(path)
    b = H.synth_reads(prg_dir, path, pairs=n_pairs, len=args.read_len, seed=0xB200 + 7919 * rank, clip_frac=0.15)
    open(path + ".ok", "w").write("ok\n")
    return b


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index; self.rows = []; self.stop_flag = False; self.proc = None

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass
        sm = []; mx = 0; reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for i, n in enumerate(names):
                    if r[2 + i].lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


_REF = {}


def cpu_reference_run(args, prg_dir_small, n_pairs, threads):
    """The unmodified reference C++ (oracle/_ref) over a bounded sample; returns pairs/s."""
    import harness as H
    from gpu_probe import quiet
    b = make_reads(args, prg_dir_small, 0, n_pairs, tag="cpu")
    if prg_dir_small not in _REF:
        _REF[prg_dir_small] = quiet(H.Ref, prg_dir_small)   # one reference graph per process (its arena is not recycled)
    R = _REF[prg_dir_small]
    out = quiet(R.pairs, b, args.is_mean, args.is_sd, 1024, threads, False)
    return n_pairs / out["seconds"], out["seconds"]


def sample_dims(args):
    """The reference arm's PRG: a slice of the bench PRG with the same generator parameters — the same gene-block density (one block per
    levels/genes levels) and the same number of alleles per block — small enough for the reference's pointer graph to load in seconds."""
    genes = max(1, int(round(args.genes * args.cpu_levels / float(args.levels))))
    return args.cpu_levels, genes, args.alleles


def small_prg(args, root):
    import harness as H
    levels, genes, alleles = sample_dims(args)
    d = os.path.join(root, "prg_cpu_sample_l%d_g%d_a%d" % (levels, genes, alleles))
    if not os.path.exists(os.path.join(d, ".complete")):
        os.makedirs(d, exist_ok=True)
        H.synth_prg(d, levels=levels, haps=args.haps, genes=genes, alleles=alleles, allele_contigs=4, seed=0xB200)
        open(os.path.join(d, ".complete"), "w").write("ok\n")
    return d


def sample_text(args):
    levels, genes, alleles = sample_dims(args)
    return ("%d pairs 2x%d per step, same read generator, on a %d-level slice of the PRG generator's output with identical parameters (%d haplotypes, %d gene block(s) x %d alleles: "
            "the bench PRG's block density and allele count)" % (args.cpu_pairs, args.read_len, levels, args.haps, genes, alleles))


def ncu_traffic_bytes(kernel_file):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full summary of the kernel (profiles/), or None"""
    path = os.path.join(REPO, "profiles", kernel_file)
    try:
        tot = 0.0; seen = 0
        for line in open(path):
            f = line.split()
            if len(f) >= 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[f[1]]
                tot += float(f[2]) * mult; seen += 1
        return tot if seen == 2 else None
    except Exception:
        return None


def stage_kmer_seeding(args, root, H):
    """§8 a29 on this box: index build (host) + hlala_seed_kmers (GPU) on the bases of a synthetic batch; the reference's findChains on a sample.
    The index enumerates every k-mer path, which explodes inside 1000-allele gene blocks (the reference's Index() likewise), so this stage
    uses a PRG of the same generator with 8 alleles per gene block."""
    d = os.path.join(root, "prg_seed_stage")
    if not os.path.exists(os.path.join(d, ".complete")):
        os.makedirs(d, exist_ok=True)
        H.synth_prg(d, levels=1000000, haps=args.haps, genes=args.genes, alleles=8, allele_contigs=4, seed=0xB200)
        open(os.path.join(d, ".complete"), "w").write("ok\n")
    b = H.synth_reads(d, os.path.join(d, "seeds_stage.bin"), pairs=250000, len=args.read_len, seed=0xB200, clip_frac=0.15)
    off = np.ascontiguousarray(b["read_off"], np.int64); bases = np.ascontiguousarray(b["bases"], np.uint8)
    P = H.Product(d); P.to_gpu(int(os.environ.get("LOCAL_RANK", "0")))
    t = time.time(); P.kmer_index(25); t_index = time.time() - t

    class RB(C.Structure):
        _fields_ = [("n_reads", C.c_int64), ("read_off", C.c_void_p), ("bases", C.c_void_p)]
    rb = RB(); rb.n_reads = len(off) - 1; rb.read_off = off.ctypes.data; rb.bases = bases.ctypes.data
    best = None
    for _ in range(3):
        res = C.c_void_p(); t = time.time()
        P._chk(P.lib.hlala_seed_kmers(P.g, C.byref(rb), C.byref(res))); dt = time.time() - t
        ms = (C.c_double * 2)(); P._chk(P.lib.hlala_kmer_chains_timing(res, ms))
        nr = C.c_int64(); nc = C.c_int64(); ne = C.c_int64(); nf = C.c_int64()
        P._chk(P.lib.hlala_kmer_chains_dims(res, C.byref(nr), C.byref(nc), C.byref(ne), C.byref(nf)))
        n2 = C.c_int64(); P._chk(P.lib.hlala_kmer_chains_second_tier_reads(res, C.byref(n2)))
        P.lib.hlala_kmer_chains_free(res)
        if best is None or ms[0] < best["kernel_ms"]:
            best = dict(kernel_ms=ms[0], order_gather_ms=ms[1], call_s=dt, chains=nc.value, chain_edges=ne.value, reads_second_tier=n2.value, reads_over_capacity=nf.value)
    n = len(off) - 1
    out = dict(workload="%d reads x %d bp, k=25, PRG of 1000000 levels / %d haplotypes / %d gene blocks x 8 alleles" % (n, args.read_len, args.haps, args.genes),
               index_build_host_s=t_index, reads_per_s_kernel=n / (best["kernel_ms"] / 1e3), reads_per_s_call_device_resident_output=n / best["call_s"], **best)
    if os.path.exists(H.LIB_REF):   # the compiled reference holds one graph per process: its sample runs in a process of its own
        try:
            dref = os.path.join(root, "prg_seed_stage_ref")   # the reference's Index() needs minutes and tens of GB at 1M levels: its sample uses 100000 levels of the same generator
            r = subprocess.run([sys.executable, os.path.join(REPO, "tools", "seed_probe.py"), "--dir", dref, "--levels", "100000", "--alleles", "8", "--genes", str(args.genes), "--pairs", "20000", "--len", str(args.read_len),
                                "--ref-only", "1", "--ref-reads", "20000"],
                               stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=600)
            ref = json.loads(r.stdout.strip().splitlines()[-1])
            out.update(reference_reads_per_s_1_thread=ref["ref_reads_per_s"], reference_index_build_s=ref["ref_index_build_s"],
                       reference_sample="%d reads on a 100000-level PRG of the same generator, unmodified GraphAndEdgeIndex::findChains (single-threaded, as the reference is)" % ref["ref_reads"])
        except Exception as e:
            out["reference_error"] = str(e)[:200]
    P.close()
    return out


def stage_typing(args, root, H):
    """§8 a20-a28 on this box: alignment session with columns kept -> gene filter/extract -> hlala_typer_infer (per-read x cluster kernel,
    allele-pair kernel) on a PRG with 17 gene blocks x 200 alleles and reads concentrated in the gene blocks."""
    import tempfile
    d = os.path.join(root, "prg_typing_stage_a%d" % args.alleles)
    if not os.path.exists(os.path.join(d, ".complete")):
        os.makedirs(d, exist_ok=True)
        H.synth_prg(d, levels=200000, haps=4, genes=17, alleles=args.alleles, seed=11)
        open(os.path.join(d, ".complete"), "w").write("ok\n")
    b = H.synth_reads(d, os.path.join(d, "seeds_stage.bin"), pairs=60000, len=args.read_len, seed=11, gene_frac=0.8)
    P = H.Product(d); P.to_gpu(int(os.environ.get("LOCAL_RANK", "0")))
    T = H.ProductTyping(P, d)
    sess = H.session_align(P, b, args.is_mean, args.is_sd, args.max_columns, keep_columns=True)
    t = time.time(); blob, n_sel = T.extract(sess); t_extract = time.time() - t
    out_dir = tempfile.mkdtemp(prefix="hlala_typing_stage_")
    t = time.time(); T.infer([blob], args.is_mean, args.is_sd, out_dir, device=int(os.environ.get("LOCAL_RANK", "0")), keep_read_ll=False); t_infer = time.time() - t
    tm = T.timing()
    P.lib.hlala_session_free(sess)
    res = dict(workload="60000 pairs 2x%d (80%% inside gene blocks), PRG of 200000 levels / 17 gene blocks x %d alleles (the bench PRG's allele count)" % (args.read_len, args.alleles), pairs_selected=n_sel,
               extract_s=t_extract, infer_call_s=t_infer, read_x_cluster_kernel_ms=tm["ms"][0], allele_pair_kernel_ms=tm["ms"][1], kernel_launches=tm["launches"],
               read_cluster_observation_steps=tm["work"][0], log_avg_evaluations=tm["work"][1],
               log_avg_evaluations_per_s=(tm["work"][1] / (tm["ms"][1] / 1e3)) if tm["ms"][1] > 0 else None, files_written=len(os.listdir(out_dir)), megabytes_written=sum(os.path.getsize(os.path.join(out_dir, f)) for f in os.listdir(out_dir)) / 1e6)
    T.close(); P.close()
    # the allele-pair kernels alone at the cluster count of the real class-I loci (SURVEY config C2: C = 4000), R = 10000 reads
    try:
        rng = np.random.default_rng(1); Cn, R = 4000, 10000
        ll = np.ascontiguousarray(-rng.gamma(2.0, 15.0, size=(Cn, R))); mm = rng.integers(0, 6, size=(Cn, R)).astype(np.int32); npair = Cn * (Cn + 1) // 2
        probe = {}
        lib = C.CDLL(H.LIB_PRODUCT); lib.hlala_typing_pair_probe.argtypes = [C.c_int, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        for termwise in (0, 1):
            pl = np.zeros(npair); ms = C.c_double(0)
            rc = lib.hlala_typing_pair_probe(int(os.environ.get("LOCAL_RANK", "0")), Cn, R, H.p(ll), H.p(mm), termwise, H.p(pl), None, None, C.byref(ms))
            probe["termwise" if termwise else "max_shifted"] = {"kernel_ms": ms.value, "log_avg_terms_per_s": npair * R / (ms.value / 1e3) if rc == 0 and ms.value > 0 else None, "rc": rc, "checksum": float(pl.sum())}
        probe["relative_difference_of_sums"] = abs(probe["max_shifted"]["checksum"] - probe["termwise"]["checksum"]) / abs(probe["termwise"]["checksum"])
        res["allele_pair_kernel_at_C4000_R10000"] = probe
    except Exception as e:
        res["allele_pair_kernel_at_C4000_R10000"] = {"error": str(e)[:200]}
    return res


def stage_long_reads(args, root, H):
    """BASELINE.json configs[3]: long-read mode, 50 k synthetic 8 kb reads on one B200 through hlala_align_long_reads (host buffers in, per-read results and
    per-level coverage out); the unmodified processBAM::alignOneLongRead on the first reads of the same batch, same PRG, one thread (what the reference does).
    The reference has no DP in this mode (the seed is padded to the read, processBAM.cpp:3728-3731), so neither has the GPU path."""
    d = small_prg(args, root)
    n, L, cap = args.long_reads, 8000, 10240
    b = H.synth_reads(d, os.path.join(d, "long_stage.bin"), pairs=n, len=L, single=1, indel_rate=0.03, clip_frac=0.5, clip_max=400, gene_frac=0.2, seed=0xB200)
    P = H.Product(d); P.to_gpu(int(os.environ.get("LOCAL_RANK", "0")))
    nr = len(b["read_off"]) - 1
    o = dict(pair_mapq=np.zeros(nr), read_mapq=np.zeros(nr), read_reverse=np.zeros(nr, np.uint8), chosen_slot=np.zeros(nr, np.int32), pair_ll=np.zeros(nr), n_cols=np.zeros(nr, np.int32))
    po = H.PairOut(); po.max_columns = cap
    for k in o:
        setattr(po, k, o[k].ctypes.data)
    sb = H.make_batch_struct(b); bpl = np.zeros(P.dims()["n_levels"], np.int32); ts = []
    for _ in range(3):
        t = time.perf_counter(); P._chk(P.lib.hlala_align_long_reads(P.g, C.byref(sb), C.byref(po), H.p(bpl))); ts.append(time.perf_counter() - t)
    levels, genes, alleles = sample_dims(args)
    out = dict(workload="%d single reads x %d bases (3 %% indels, half clipped by up to 400 bases, %.1f BAM records per read), PRG of %d levels / %d haplotypes / %d gene block(s) x %d alleles, max_columns %d"
               % (nr, L, len(b["chain_contig"]) / float(nr), levels, args.haps, genes, alleles, cap), call_s=min(ts[1:]), reads_per_s_e2e=nr / min(ts[1:]), bases_per_s_e2e=nr * L / min(ts[1:]),
               reads_with_mapq_lt_1=int((o["pair_mapq"] < 1).sum()), sum_columns=int(o["n_cols"].sum()), sum_ll=float(o["pair_ll"].sum()))
    P.close()
    if os.path.exists(H.LIB_REF):
        try:
            k = min(args.long_ref_reads, nr); nch = int(b["chain_off"][k]); ncg = int(b["cigar_off"][nch]); nb = int(b["read_off"][k])
            sbt = dict(read_off=b["read_off"][:k + 1].copy(), bases=b["bases"][:nb].copy(), quals=b["quals"][:nb].copy(), chain_off=b["chain_off"][:k + 1].copy(),
                       chain_contig=b["chain_contig"][:nch].copy(), chain_pos=b["chain_pos"][:nch].copy(), chain_flag=b["chain_flag"][:nch].copy(), chain_as=b["chain_as"][:nch].copy(),
                       cigar_off=b["cigar_off"][:nch + 1].copy(), cigar=b["cigar"][:ncg].copy())
            if d in _REF:
                want = H.quiet(_REF[d].long_reads, sbt, cap, 1, False)
            else:
                want = H._ref_subprocess(d, sbt, 0.0, 1.0, cap, "long_reads")
            sec = float(want["seconds"]) if "seconds" in want else None
            same = bool(np.array_equal(want["n_cols"], o["n_cols"][:k]) and np.array_equal(want["read_ll"], o["pair_ll"][:k]) and np.allclose(want["read_mapq"], o["pair_mapq"][:k], rtol=0, atol=1e-12))
            out.update(reference_reads=k, reference_seconds_1_thread=sec, reference_reads_per_s_1_thread=(k / sec) if sec else None, sample_identical_to_reference=same)
        except Exception as e:
            out["reference_error"] = str(e)[:200]
    return out


def stage_fastq_mapper(args, root, H):
    """paired FASTQ files -> hlala_fastq_map_pairs (host code: k-mer votes per contig and diagonal, banded alignment with bwa mem's default scores, all
    placements) -> hlala_align_pairs on the GPU -> bases on their true level (hlala_truth_evaluate), next to the same figure for the generator's own seeds. There is no
    reference arm: the reference calls bwa here, which is an external program."""
    d = small_prg(args, root)
    n = args.fastq_pairs
    pre = os.path.join(d, "fq_stage_R")
    b = H.synth_reads(d, os.path.join(d, "fq_stage.bin"), pairs=n, len=args.read_len, clip_frac=0.15, seed=0xB200 + 7, levels_prefix=pre)
    P = H.Product(d); ts = []
    for _ in range(2):
        t = time.perf_counter(); mb, names, cnt = P.fastq_map(pre + "_1.fq", pre + "_2.fq"); ts.append(time.perf_counter() - t)
    levels, genes, alleles = sample_dims(args)
    out = dict(workload="%d pairs of 2 x %d bp as FASTQ, PRG of %d levels / %d haplotypes / %d gene block(s) x %d alleles (%d contig bases)" % (n, args.read_len, levels, args.haps, genes, alleles, int(P.array("contig_off")[-1])),
               host_threads=os.cpu_count(), map_call_s=min(ts), reads_per_s=2 * n / min(ts), pairs_placed=len(names), pairs_with_an_unplaced_mate=int(cnt["incomplete"]),
               placements_per_read=len(mb["chain_contig"]) / float(max(1, len(mb["read_off"]) - 1)), generator_chains_per_read=len(b["chain_contig"]) / float(len(b["read_off"]) - 1))
    try:
        P.to_gpu(int(os.environ.get("LOCAL_RANK", "0")))
        k = min(len(names), 8000)            # the accuracy figure on a slice: the columns of every read come back to the host for the comparison
        sub = H.subset_pairs(mb, np.arange(k)); t = time.perf_counter(); aln = P.pairs(sub, args.is_mean, args.is_sd, cap=args.max_columns, want_levels=False); out["align_slice_s"] = time.perf_counter() - t
        _per, tot, _n = H.truth_evaluate(aln, pre + "_1.levels", pre + "_2.levels", names=names[:k])
        out.update(slice_pairs=k, bases_on_true_level_mapper_seeds=float(tot[1]) / float(tot[0]), reads_below_90_percent=int(tot[2]))
        idx = np.array([int(x[1:]) for x in names[:k]])
        aln = P.pairs(H.subset_pairs(b, idx), args.is_mean, args.is_sd, cap=args.max_columns, want_levels=False)
        _per, tot, _n = H.truth_evaluate(aln, pre + "_1.levels", pre + "_2.levels", names=names[:k])
        out["bases_on_true_level_generator_seeds"] = float(tot[1]) / float(tot[0])
    except Exception as e:
        out["gpu_error"] = str(e)[:300]
    P.close()
    return out


def strong_scaling_leg(args, H, P, prg_dir, rank, local_rank, world, dist, torch, n_levels, batch=None):
    """BASELINE.json configs[2] on N GPUs: --strong-pairs pairs in total, sharded by pair across the ranks (graph replicated). Timed per rank, max over ranks:
    alignment of the shard (columns kept) -> all-reduce of the per-level coverage -> extraction of the gene-overlapping pairs -> all-gather of those (small)
    alignment blobs -> typing with the reads split across ranks and ONE NCCL all-reduce of the allele-pair sums per locus -> rank 0 writes the hla/* files."""
    import tempfile
    n_shard = args.strong_pairs // world if batch is None else (len(batch["read_off"]) - 1) // 2
    b = make_reads(args, prg_dir, 1000 + rank, n_shard, tag="strong%d" % world) if batch is None else batch
    T = H.ProductTyping(P, prg_dir)
    L = P.lib; sb = H.make_batch_struct(b); sess = C.c_void_p()
    P._chk(L.hlala_session_create(P.g, C.byref(sb), C.c_int32(args.max_columns), C.byref(sess))); P._chk(L.hlala_session_set_keep_columns(sess, 1))
    cov = torch.zeros(n_levels - 1, dtype=torch.int32, device="cuda"); stream = torch.cuda.current_stream()
    n_calls = [0]

    def allreduce(ctx, ptr, count, st):     # entered from a worker thread of the library (one locus at a time, in locus order): a new thread starts on device 0
        torch.cuda.set_device(local_rank)
        t = H.dev_f64_tensor(ptr, count, local_rank)
        torch.cuda.synchronize(); dist.all_reduce(t); torch.cuda.synchronize(); n_calls[0] += 1
        return 0

    class _Solo:      # N = 1: the same leg without collectives (the strong-scaling baseline)
        class ReduceOp:
            MAX = None

        @staticmethod
        def barrier():
            pass

        @staticmethod
        def all_reduce(t, op=None):
            return t

        @staticmethod
        def all_gather_object(lst, obj):
            lst[0] = obj
    if dist is None:
        dist = _Solo
    out = tempfile.mkdtemp(prefix="hlala_strong_") if rank == 0 else None
    res = None
    for it in range(2):     # the second pass is the timed one
        cov.zero_(); n_calls[0] = 0
        torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
        P._chk(L.hlala_session_run(sess, C.c_double(args.is_mean), C.c_double(args.is_sd), C.c_uint64(cov.data_ptr()), C.c_void_p(stream.cuda_stream)))
        dist.all_reduce(cov); torch.cuda.synchronize(); t_align = time.perf_counter() - t0
        blob, nsel = T.extract(sess, base=rank * n_shard)
        blobs = [None] * world; dist.all_gather_object(blobs, blob); t_gather = time.perf_counter() - t0
        T.infer(blobs, args.is_mean, args.is_sd, os.path.join(out, "hla%d" % it) if out else None, device=local_rank, rank=rank, world=world, allreduce=allreduce if world > 1 else None, keep_read_ll=False, callback_any_thread=True)
        torch.cuda.synchronize(); t_all = time.perf_counter() - t0
        tt = torch.tensor([t_align, t_gather, t_all, float(nsel)], dtype=torch.float64, device="cuda")
        tmax = tt.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX); tsum = tt.clone(); dist.all_reduce(tsum)
        res = {"total_pairs": n_shard * world, "pairs_per_rank": n_shard, "seconds": float(tmax[2]), "pairs_per_s": n_shard * world / float(tmax[2]),
               "seconds_alignment_plus_coverage_allreduce": float(tmax[0]), "seconds_until_blobs_gathered": float(tmax[1]), "pairs_selected_for_typing": int(tsum[3]),
               "allele_pair_allreduces_per_rank": n_calls[0], "gathered_blob_bytes": int(sum(len(x) for x in blobs)),
               "files_written_by_rank0": len(os.listdir(os.path.join(out, "hla%d" % it))) if out else None,
               "note": "strong scaling: the total is fixed, every rank aligns total/N pairs; wall clock, max over ranks, second of two passes; inputs resident in HBM, typing outputs written by rank 0"}
    L.hlala_session_free(sess); T.close()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--pairs", type=int, default=1000000, help="read pairs per GPU per step (BASELINE.json configs[1]: 1M pairs 2x150)")
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--levels", type=int, default=5000000)
    ap.add_argument("--haps", type=int, default=8)
    ap.add_argument("--genes", type=int, default=17)
    ap.add_argument("--alleles", type=int, default=1000)
    ap.add_argument("--is-mean", type=float, default=100.0)
    ap.add_argument("--is-sd", type=float, default=10.0)
    ap.add_argument("--max-columns", type=int, default=0, help="columns per alignment; 0 = 640 up to 150 bp reads (the headline), 4 x read length + 64 beyond (rare alignments span gap stretches of hundreds of levels)")
    ap.add_argument("--cpu-pairs", type=int, default=6000)
    ap.add_argument("--cpu-levels", type=int, default=294118, help="levels of the reference arm's PRG slice (default: levels / genes = one gene block, the bench PRG's density)")
    ap.add_argument("--fastq-pairs", type=int, default=20000, help="pairs of the FASTQ mapper stage (host code in front of the GPU path)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--stages", type=int, default=1, help="also time the k-mer seeding and typing stages (rank 0, N=1 only)")
    ap.add_argument("--long-reads", type=int, default=50000, help="reads of the long-read stage (BASELINE.json configs[3]: 50 k x 8 kb)")
    ap.add_argument("--long-ref-reads", type=int, default=100, help="reads of that batch the reference's alignOneLongRead is timed on (1 thread)")
    ap.add_argument("--pipeline", type=int, default=1, help="N = 1: also time the headline batch through alignment + typing + files (BASELINE.json configs[1] end to end)")
    ap.add_argument("--strong-single", type=int, default=0, help="run the strong-scaling leg at N = 1 too (its baseline; off by default to keep the default run short)")
    ap.add_argument("--strong-pairs", type=int, default=4000000, help="N > 1: total pairs of the strong-scaling leg (BASELINE.json configs[2]: ~4M pairs sharded over the GPUs, typing with one NCCL all-reduce per locus); 0 = skip")
    args = ap.parse_args()
    if args.max_columns <= 0:
        args.max_columns = 640 if args.read_len <= 150 else min(2040, 4 * args.read_len + 64)
    rank, local_rank, world = rank_info()
    n_gpus = max(world, 1)

    ensure_built()
    import harness as H
    prg_dir, root = workload_dirs(args)
    os.makedirs(root, exist_ok=True)
    config = {"workload": "%d synthetic 2x%dbp pairs per GPU, bwa-style seed chains (15%% soft-clipped), synthetic PRG of %d levels / %d haplotypes / %d gene blocks x %d alleles; "
                          "full seed projection + graph-DP extension + pair likelihood/mapQ + per-level coverage" % (args.pairs, args.read_len, args.levels, args.haps, args.genes, args.alleles),
              "pairs_per_gpu": args.pairs, "read_len": args.read_len, "levels": args.levels, "sharding": "pairs sharded across ranks (weak scaling)",
              "cache": "batch (>1 GB) and per-wave scratch exceed the 126 MB L2, so every step streams its inputs from HBM",
              "reference_arm_sample": "the reference's CPU implementation cannot load or finish the full workload in minutes; `--impl reference` and cpu_baseline run " + sample_text(args)}

    # ------------------------------------------------------------------ reference arm: the reference's own CPU implementation
    if args.impl == "reference":
        if rank != 0:
            return 0
        if not os.path.exists(H.LIB_REF):
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libhlala_ref.so not built on this box"}))
            return 0
        d = small_prg(args, root)
        threads = os.cpu_count() or 1
        vals = []
        n = args.cpu_pairs
        for i in range(args.warmup + args.steps):
            v, sec = cpu_reference_run(args, d, n, threads)
            if i >= args.warmup:
                vals.append((v, sec))
        tot_pairs = n * len(vals); tot_sec = sum(s for _, s in vals)
        value = tot_pairs / tot_sec
        sample = sample_text(args) + "; reference TUs unmodified, outer OpenMP loop over pairs (%d threads)" % threads
        line = {"metric": "paired reads/sec aligned to the PRG (seed projection + extension + pair scoring)", "value": value, "unit": "pairs/s", "n_gpus": 0, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1000.0 * tot_sec / max(len(vals), 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32/f64", "data": "synthetic",
                "config": config, "impl": "reference", "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": "reference", "sample": sample},
                "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm
    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — this path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:   # every rank keeps to its own slice of the host cores (chain sorting and the typing host stage are threaded): the e2e numbers of N ranks then do not fight over the same cores
        try:
            cores = sorted(os.sched_getaffinity(0)); lw = int(os.environ.get("LOCAL_WORLD_SIZE", str(world))); per = max(1, len(cores) // lw)
            os.sched_setaffinity(0, set(cores[local_rank * per:(local_rank + 1) * per]) or set(cores))
            os.environ["HLALA_HOST_THREADS"] = str(per)
        except Exception:
            pass
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    if rank == 0:
        make_prg(args, prg_dir)
    if dist:
        dist.barrier()
    b = make_reads(args, prg_dir, rank, args.pairs)
    P = H.Product(prg_dir)
    P.to_gpu(local_rank)
    L = P.lib
    L.hlala_session_chain_kernel_bytes.restype = C.c_int64
    L.hlala_session_algorithmic_bytes.restype = C.c_int64
    L.hlala_session_dp_kernel_bytes.restype = C.c_int64
    n_levels = P.dims()["n_levels"]
    sb = H.make_batch_struct(b)
    sess = C.c_void_p()
    P._chk(L.hlala_session_create(P.g, C.byref(sb), C.c_int32(args.max_columns), C.byref(sess)))
    cov = torch.zeros(n_levels - 1, dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream()

    def step():
        P._chk(L.hlala_session_run(sess, C.c_double(args.is_mean), C.c_double(args.is_sd), C.c_uint64(cov.data_ptr()), C.c_void_p(stream.cuda_stream)))
        if dist:
            dist.all_reduce(cov)   # the path's one exchange: per-level coverage (reads_per_level.txt is written from the sum)

    for _ in range(args.warmup):
        cov.zero_(); step()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    sampler = ClockSampler(local_rank); sampler.start()
    L.hlala_session_set_timing(sess, 1)
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(args.steps):
        cov.zero_(); step()
    ev1.record()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.finish()
    kms = (C.c_double * 6)(); kl = (C.c_int * 6)()
    P._chk(L.hlala_session_timing(sess, kms, kl))
    dp_bytes_total = int(L.hlala_session_dp_kernel_bytes(sess))
    chain_bytes = int(L.hlala_session_chain_kernel_bytes(sess)); total_bytes = int(L.hlala_session_algorithmic_bytes(sess))
    L.hlala_session_set_timing(sess, 0)
    launches = L.hlala_session_launches(sess)
    dig = (C.c_int64 * 4)(); sll = C.c_double(0)
    P._chk(L.hlala_session_digest(sess, dig, C.byref(sll)))
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = n_gpus * args.pairs * args.steps / (ms_total / 1000.0)
    if rank == 0:
        sys.stderr.write("[bench] resident: %.1f ms/step, %.0f pairs/s; kernel ms/step: seed %.1f extend %.1f finish %.1f pair %.1f; launches %d; errors %d\n" % (
            ms_total / args.steps, value, kms[0] / args.steps, (kms[1] + kms[4] + kms[5]) / args.steps, kms[2] / args.steps, kms[3] / args.steps, launches, dig[3]))
        sys.stderr.write("[bench] extension ms/step: lean %.1f warp tiers %.1f scalar %.1f\n" % (kms[1] / args.steps, kms[4] / args.steps, kms[5] / args.steps))

    # ---- end to end through the host-buffer C-ABI call (pinned host inputs, H2D + kernels + D2H inside the timed region)
    pinned = {k: torch.from_numpy(b[k]).pin_memory() for k in H.BATCH_KEYS}
    sbp = H.SeedBatch(); sbp.n_reads = len(b["read_off"]) - 1
    for k in H.BATCH_KEYS:
        setattr(sbp, k, pinned[k].data_ptr())
    nr = len(b["read_off"]) - 1
    res = {"pair_mapq": torch.empty(nr // 2, dtype=torch.float64).pin_memory(), "read_mapq": torch.empty(nr, dtype=torch.float64).pin_memory(),
           "chosen_slot": torch.empty(nr, dtype=torch.int32).pin_memory(), "pair_ll": torch.empty(nr // 2, dtype=torch.float64).pin_memory()}
    po = H.PairOut(); po.max_columns = args.max_columns
    for k, v in res.items():
        setattr(po, k, v.data_ptr())
    cov_host = np.zeros(n_levels - 1, np.int32)
    h2d = int(sum(b[k].nbytes for k in H.BATCH_KEYS)) + 3 * 4 * len(b["chain_contig"])
    d2h = int(sum(v.numel() * v.element_size() for v in res.values())) + cov_host.nbytes
    L.hlala_session_free(sess)   # the e2e call owns its device buffers (a workspace kept in the graph handle across calls)
    # one extra step on a single lane (no overlap between waves): the per-kernel split that is comparable with a serialised profiler run
    serial = None
    if rank == 0:
        try:
            os.environ["HLALA_LANES"] = "1"
            s1 = C.c_void_p()
            P._chk(L.hlala_session_create(P.g, C.byref(sb), C.c_int32(args.max_columns), C.byref(s1)))
            del os.environ["HLALA_LANES"]
            P._chk(L.hlala_session_run(s1, C.c_double(args.is_mean), C.c_double(args.is_sd), C.c_uint64(0), C.c_void_p(stream.cuda_stream)))
            L.hlala_session_set_timing(s1, 1)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
            P._chk(L.hlala_session_run(s1, C.c_double(args.is_mean), C.c_double(args.is_sd), C.c_uint64(0), C.c_void_p(stream.cuda_stream)))
            e1.record(); torch.cuda.synchronize()
            k1 = (C.c_double * 6)(); l1 = (C.c_int * 6)(); P._chk(L.hlala_session_timing(s1, k1, l1))
            serial = {"ms_per_step": e0.elapsed_time(e1), "kernel_ms": [k1[i] for i in range(6)]}
            L.hlala_session_free(s1)
        except Exception as e:
            os.environ.pop("HLALA_LANES", None)
            serial = {"error": str(e)[:200]}
    torch.cuda.synchronize()
    e2e_times = []
    for i in range(1 + args.e2e_steps):
        if dist:
            dist.barrier()
        t0 = time.perf_counter()
        cov_host[:] = 0
        P._chk(L.hlala_align_pairs(P.g, C.byref(sbp), C.c_double(args.is_mean), C.c_double(args.is_sd), C.byref(po), cov_host.ctypes.data_as(C.c_void_p)))
        dt = time.perf_counter() - t0
        if i > 0:
            e2e_times.append(dt)
    te = torch.tensor([sum(e2e_times)], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = n_gpus * args.pairs * len(e2e_times) / float(te.item()) if e2e_times else None
    if rank == 0:
        sys.stderr.write("[bench] e2e: %s pairs/s (%s s per call)\n" % (e2e_value, [round(x, 3) for x in e2e_times]))

    L.hlala_graph_release_workspace(P.g)     # the device workspace hlala_align_pairs keeps in the graph handle (about the size of a session): the legs below bring their own
    strong = None
    if args.strong_pairs > 0 and (dist or args.strong_single):
        strong = strong_scaling_leg(args, H, P, prg_dir, rank, local_rank, world, dist, torch, n_levels)
    pipeline = None
    if args.pipeline and not dist:
        # BASELINE.json configs[1] end to end on the bench batch itself: alignment (columns kept) -> coverage -> gene filter -> typing kernels -> the 73 hla/* files,
        # batch resident in HBM when the clock starts (the strong-scaling leg at N = 1, on the headline workload)
        try:
            pipeline = strong_scaling_leg(args, H, P, prg_dir, rank, local_rank, world, None, torch, n_levels, batch=b)
            pipeline["note"] = "the headline batch through the whole path of BASELINE config 2: alignment with columns kept, per-level coverage, gene filter + extraction, typing (two kernels per locus) and all hla/* files; wall clock, second of two passes, batch resident in HBM"
            for k in ("allele_pair_allreduces_per_rank", "pairs_per_rank", "gathered_blob_bytes"):
                pipeline.pop(k, None)
        except Exception as e:
            pipeline = {"error": str(e)[:300]}
    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return 0
    peak, peak_src = measured_peak()
    names = ["k_chain_seed", "k_extend_lean", "k_chain_finish", "k_pair", "k_extend_warp(tiny+small+large)", "k_extend(scalar)"]
    ab = (C.c_int64 * 3)(); aligned = None
    try:
        s2 = C.c_void_p(); P._chk(L.hlala_session_create(P.g, C.byref(sb), C.c_int32(args.max_columns), C.byref(s2)))
        P._chk(L.hlala_session_run(s2, C.c_double(args.is_mean), C.c_double(args.is_sd), C.c_uint64(0), C.c_void_p(stream.cuda_stream)))
        torch.cuda.synchronize(); P._chk(L.hlala_session_aligned_bytes(s2, ab)); L.hlala_session_free(s2)
        aligned = {"chains_aligned": int(ab[0]), "chains_in_batch": int(len(b["chain_contig"])), "chain_kernel_bytes_per_step": int(ab[1]), "whole_path_bytes_per_step": int(ab[2])}
    except Exception as e:
        aligned = {"error": str(e)[:200]}
    per_kernel = {names[i]: {"ms_per_step": kms[i] / args.steps, "launches_per_step": kl[i] / args.steps} for i in range(6)}
    dom = max(range(6), key=lambda i: kms[i])
    # dominant kernel: the first tier of the extension DP. Algorithmic bytes of all extension tasks (it attempts every task) over the summed
    # CUDA-event durations of its launches; the launches of different waves overlap on several streams, so this is a per-launch average.
    dp_ach = dp_bytes_total / (kms[1] / 1000.0) / 1e9 if kms[1] > 0 and dp_bytes_total > 0 else None
    chain_ach = chain_bytes * args.steps / (kms[0] / 1000.0) / 1e9 if kms[0] > 0 else None
    traffic = ncu_traffic_bytes("r02_ncu_k_extend_lean.txt")
    roofline = {"bound": "hbm", "kernel": "k_extend_lean", "achieved": dp_ach, "peak": peak, "unit": "GB/s", "frac": (dp_ach / peak) if dp_ach else None, "traffic": traffic,
                "traffic_source": "profiles/r02_ncu_k_extend_lean.txt (ncu --set full of the same launch: one launch over the whole 1 M-pair batch of the default workload)", "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dp_bytes_total / max(kl[1], 1), "launch_ms_avg": kms[1] / max(kl[1], 1),
                "note": "integer DP over a sparse cell set, one thread per extension: bounded by SIMT divergence, dependent shared-memory accesses and issue slots (ncu, profiles/r02_ncu_k_extend_lean.txt: 8.8 of 32 lanes active, issue slots 25 % busy, 8 warps per SM), not by HBM bandwidth; the fraction is reported as measured",
                "chain_kernel": {"kernel": "k_chain_seed", "achieved_counting_all_chains": chain_ach, "frac_counting_all_chains": (chain_ach / peak) if chain_ach else None, "algorithmic_bytes_per_step_all_chains": chain_bytes,
                                 "achieved": (aligned["chain_kernel_bytes_per_step"] * args.steps / (kms[0] / 1000.0) / 1e9) if aligned and "error" not in aligned and kms[0] > 0 else None,
                                 "frac": (aligned["chain_kernel_bytes_per_step"] * args.steps / (kms[0] / 1000.0) / 1e9 / peak) if aligned and "error" not in aligned and kms[0] > 0 else None,
                                 "note": "achieved / frac count only the chains the kernels read (k_prepare skips chains that the pair stage would discard as duplicates: 86 % of this workload's chains)"},
                "aligned_chains": aligned,
                "whole_path_algorithmic_bytes_per_step_all_chains": total_bytes, "whole_path_achieved_counting_all_chains": total_bytes * args.steps / (ms_total / 1000.0) / 1e9,
                "whole_path_algorithmic_bytes_per_step": aligned["whole_path_bytes_per_step"] if aligned and "error" not in aligned else None,
                "whole_path_achieved": (aligned["whole_path_bytes_per_step"] * args.steps / (ms_total / 1000.0) / 1e9) if aligned and "error" not in aligned else None,
                "dominant_kernel_by_time": names[dom], "per_kernel": per_kernel,
                "per_kernel_note": "durations of launches on different streams overlap (and slow each other down); their sum exceeds ms_per_step"}
    if serial and "kernel_ms" in serial:
        tot1 = sum(serial["kernel_ms"]) or 1.0
        roofline["single_lane_step"] = {"ms_per_step": serial["ms_per_step"], "per_kernel_ms": {names[i]: serial["kernel_ms"][i] for i in range(6)},
                                        "per_kernel_share": {names[i]: serial["kernel_ms"][i] / tot1 for i in range(6)},
                                        "note": "one step with HLALA_LANES=1: waves one after the other, no overlap; these shares are the ones to compare with the ncu launch list"}
        roofline["dominant_kernel_by_time"] = names[max(range(6), key=lambda i: serial["kernel_ms"][i])]
    elif serial:
        roofline["single_lane_step"] = serial
    cpu = None
    if os.path.exists(H.LIB_REF) and n_gpus == 1:
        d = small_prg(args, root)
        v1, s1 = cpu_reference_run(args, d, max(500, args.cpu_pairs // 4), 1)
        threads = os.cpu_count() or 1
        vN, sN = cpu_reference_run(args, d, args.cpu_pairs, threads)
        cpu = {"value": vN, "unit": "pairs/s", "cores": threads, "kind": "reference",
               "single_thread_value": v1,
               "sample": sample_text(args) + "; unmodified reference TUs; the reference itself runs this loop on 1 thread (%.0f pairs/s), the value is the courtesy all-cores OpenMP loop over pairs" % v1}
        # the GPU path on EXACTLY the reference arm's sample (same PRG slice, same reads), end to end through the host-buffer call
        try:
            bs = make_reads(args, d, 0, args.cpu_pairs, tag="cpu")
            Ps = H.Product(d); Ps.to_gpu(local_rank)
            ts = []
            for _ in range(3):
                t0 = time.perf_counter(); Ps.pairs(bs, args.is_mean, args.is_sd, cap=args.max_columns); ts.append(time.perf_counter() - t0)
            Ps.close()
            gv = args.cpu_pairs / min(ts[1:])
            cpu["same_workload"] = {"gpu_e2e_pairs_per_s": gv, "reference_pairs_per_s_all_cores": vN, "reference_pairs_per_s_1_thread": v1, "ratio_vs_all_cores": gv / vN, "ratio_vs_1_thread": gv / v1,
                                    "note": "hlala_align_pairs on the reference arm's own sample (host buffers in, per-pair results out); the batch is far too small to fill the GPU, so this is a lower bound of the same-workload ratio"}
        except Exception as e:
            cpu["same_workload"] = {"error": str(e)[:200]}
    stages = None
    if args.stages and n_gpus == 1:
        stages = {}
        for nm, fn in (("kmer_seeding", stage_kmer_seeding), ("typing", stage_typing), ("long_reads", stage_long_reads), ("fastq_mapper", stage_fastq_mapper)):
            try:
                stages[nm] = fn(args, root, H)
            except Exception as e:   # a stage measurement must not cost the headline line
                stages[nm] = {"error": str(e)[:300]}
    line = {"metric": "paired reads/sec aligned to the PRG (seed projection + extension + pair scoring)", "value": value, "unit": "pairs/s", "n_gpus": n_gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32/f64", "data": "synthetic",
            "config": config, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches) * args.steps, "roofline": roofline, "cpu_baseline": cpu,
            "stages": stages, "pipeline_config2": pipeline, "strong_scaling_config3": strong, "check": {"sum_columns": int(dig[0]), "edge_checksum": int(dig[1]), "pairs_mapq_lt_1": int(dig[2]), "errors": int(dig[3]), "sum_pair_ll": sll.value}}
    print(json.dumps(line))
    if dist:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
