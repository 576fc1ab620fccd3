#!/usr/bin/env python
"""Multi-GPU end of the path on real NCCL: run under torchrun with N ranks on one node.
Every rank aligns its contiguous block of pairs, extracts the gene-overlapping pairs, the blobs are all-gathered, every rank runs the
typing kernels on its slice of the reads and the allele-pair sums are combined with ONE NCCL all-reduce per locus through the C-ABI
callback; rank 0 writes hla/*. The per-level coverage is all-reduced too. Rank 0 then repeats everything on one GPU and compares:
coverage identical, calls identical, allele-pair sums within 1e-12 (relative). Prints one JSON line.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/typing_nccl_check.py"""
import ctypes as C
import json
import os
import sys
import tempfile
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "tests")); sys.path.insert(0, os.path.join(REPO, "hla-la_b200"))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import harness as H  # noqa: E402
import hlala_dist  # noqa: E402


def main():
    rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    mu, sd = 100.0, 10.0
    d = "/tmp/hlala_nccl_check_prg"
    if rank == 0 and not os.path.exists(d + "/.complete"):
        os.makedirs(d, exist_ok=True); H.synth_prg(d, levels=120000, haps=4, genes=17, alleles=100, seed=11)
        H.synth_reads(d, d + "/seeds.bin", pairs=40000, len=150, seed=11, gene_frac=0.6); open(d + "/.complete", "w").write("ok")
    dist.barrier()
    b = H.read_arrayfile(d + "/seeds.bin")
    P = H.Product(d); P.to_gpu(local); T = H.ProductTyping(P, d)
    n_pairs = (len(b["read_off"]) - 1) // 2; nl = P.dims()["n_levels"]

    def align_extract(batch, base):
        L = P.lib; sb = H.make_batch_struct(batch); sess = C.c_void_p()
        P._chk(L.hlala_session_create(P.g, C.byref(sb), C.c_int32(640), C.byref(sess))); P._chk(L.hlala_session_set_keep_columns(sess, 1))
        cov = torch.zeros(nl - 1, dtype=torch.int32, device="cuda")
        P._chk(L.hlala_session_run(sess, C.c_double(mu), C.c_double(sd), C.c_uint64(cov.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        torch.cuda.synchronize()
        blob, nsel = T.extract(sess, base=base); L.hlala_session_free(sess)
        return cov, blob, nsel

    t0 = time.time()
    p0, _p1 = hlala_dist.shard_bounds(n_pairs, rank, world)
    cov, blob, nsel = align_extract(hlala_dist.shard_batch(b, rank, world), p0)
    dist.all_reduce(cov)                                   # exchange 1: per-level coverage
    blobs = [None] * world; dist.all_gather_object(blobs, blob)      # exchange 2: the (small) gene-overlapping alignments, rank order == pair order
    n_calls = [0]

    def allreduce(ctx, ptr, count, stream):                # exchange 3: one sum all-reduce of the allele-pair vectors per locus
        t = H.dev_f64_tensor(ptr, count, local)
        torch.cuda.current_stream().synchronize(); dist.all_reduce(t); torch.cuda.synchronize(); n_calls[0] += 1
        return 0
    out = tempfile.mkdtemp(prefix="hlala_nccl_") if rank == 0 else None
    T.infer(blobs, mu, sd, os.path.join(out, "hla") if out else None, device=local, rank=rank, world=world, allreduce=allreduce, keep_read_ll=False)
    multi = [T.locus(i, read_ll=False) for i in range(T.n_loci)]
    t_multi = time.time() - t0
    dist.barrier()
    if rank == 0:
        cov1, blob1, nsel1 = align_extract(b, 0)
        T1 = H.ProductTyping(P, d); out1 = tempfile.mkdtemp(prefix="hlala_single_")
        T1.infer([blob1], mu, sd, os.path.join(out1, "hla"), device=local, keep_read_ll=False)
        single = [T1.locus(i, read_ll=False) for i in range(T1.n_loci)]
        worst = max(float(np.max(np.abs(m["pair_ll"] - s["pair_ll"]) / np.maximum(1.0, np.abs(s["pair_ll"])))) for m, s in zip(multi, single))
        same_calls = all((m["call1"], m["call2"]) == (s["call1"], s["call2"]) for m, s in zip(multi, single))
        bg = open(os.path.join(out, "hla", "R1_bestguess.txt")).read() == open(os.path.join(out1, "hla", "R1_bestguess.txt")).read()
        line = dict(tool="typing_nccl_check", world=world, pairs=n_pairs, pairs_selected_single=nsel1, loci=T.n_loci, allreduce_calls_per_rank=n_calls[0],
                    coverage_identical=bool(torch.equal(cov, cov1)), calls_identical=bool(same_calls), bestguess_file_identical=bool(bg), pair_ll_max_rel_diff=worst,
                    mismatch_sums_identical=bool(all(np.array_equal(m["pair_mavg"], s["pair_mavg"]) and np.array_equal(m["pair_mmin"], s["pair_mmin"]) for m, s in zip(multi, single))),
                    seconds_multi=t_multi)
        print(json.dumps(line))
        ok = line["coverage_identical"] and same_calls and bg and worst < 1e-12 and line["mismatch_sums_identical"]
    dist.barrier(); dist.destroy_process_group()
    return 0 if rank != 0 or ok else 1


if __name__ == "__main__":
    sys.exit(main())
