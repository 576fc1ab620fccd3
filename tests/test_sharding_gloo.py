"""N > 1 host logic on CPU: world_size-2 gloo. Each rank aligns its contiguous block of pairs (with the CPU oracle standing in for
the kernels, which need a GPU) and the all-reduced per-level coverage must equal the single-process result."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import harness as H

sys.path.insert(0, H.PKG)
import hlala_dist  # noqa: E402


def coverage_of(pairs_out, n_levels):
    cov = np.zeros(n_levels - 1, np.int64)
    for r in range(len(pairs_out["n_cols"])):
        n = pairs_out["n_cols"][r]; lv = pairs_out["level"][r, :n]; g = pairs_out["gchar"][r, :n]
        m = (lv != -1) & (g != ord("_"))
        np.add.at(cov, lv[m], 1)
    return cov


def _worker(rank, world, d, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b = H.read_arrayfile(os.path.join(d, "seeds.bin"))
    sb = hlala_dist.shard_batch(b, rank, world)
    o = H.Oracle(d)
    pr = o.pairs(sb, 100.0, 10.0, 512)
    nl = o.graph()["n_levels"]
    cov = torch.from_numpy(coverage_of(pr, nl))
    hlala_dist.allreduce_coverage(cov, dist)
    cnt = torch.tensor([len(pr["pair_mapq"]), int((pr["pair_mapq"] < 1).sum())])
    dist.all_reduce(cnt)
    if rank == 0:
        np.save(ret, np.concatenate([cnt.numpy(), cov.numpy()]))
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process(dataset, tmp_path):
    d, b, mu, sd = dataset("small")
    o = H.Oracle(d); pr = o.pairs(b, mu, sd, 512); nl = o.graph()["n_levels"]
    want = np.concatenate([[len(pr["pair_mapq"]), int((pr["pair_mapq"] < 1).sum())], coverage_of(pr, nl)])
    ret = str(tmp_path / "r0.npy")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, d, port, ret), nprocs=2, join=True)
    got = np.load(ret)
    assert np.array_equal(got, want)


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 100, 101):
        for w in (1, 2, 3, 8):
            blocks = [hlala_dist.shard_bounds(n, r, w) for r in range(w)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1


# ---------------------------------------------------------------------------------------------------------------- typing exchange
def _typing_worker(rank, world, d, port, out_root, fail_rank=-1):
    """each rank runs the product's typing HOST logic (hla_typing.cpp) with the test-only loop stand-in for the kernels on its slice of the reads;
    the allele-pair vectors are combined with ONE gloo all-reduce per locus through the callback, as hlala_typer_infer does with NCCL"""
    import ctypes as C
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b = H.read_arrayfile(os.path.join(d, "seeds.bin"))
    aln = H.Oracle(d).pairs(b, 100.0, 10.0, 512); rmq = np.ascontiguousarray(aln["read_mapq"], np.float64)
    lib = C.CDLL(os.path.join(H.REPO, "tests", "native", "build", "libtyping_host.so")); lib.typing_host_last_error.restype = C.c_char_p
    calls = [0]

    @C.CFUNCTYPE(C.c_int, C.POINTER(C.c_double), C.c_longlong)
    def allreduce(ptr, count):
        t = torch.from_numpy(np.ctypeslib.as_array(ptr, shape=(count,)))
        dist.all_reduce(t); calls[0] += 1
        return 0
    out = os.path.join(out_root, "hla") if rank == 0 else ""
    if out:
        os.makedirs(out, exist_ok=True)
    q = np.zeros(34, np.float64); ps = np.zeros(17, np.float64)
    if rank == fail_rank:
        os.environ["HLALA_TEST_FAIL_LOCUS"] = "3"
    n = lib.typing_host_run_ranked(d.encode(), C.c_longlong(len(b["read_off"]) - 1), H.p(b["read_off"]), H.p(b["bases"]), H.p(b["quals"]), C.c_int(512), H.p(aln["n_cols"]), H.p(aln["level"]),
                                   H.p(aln["gchar"]), H.p(aln["schar"]), H.p(aln["mapq"]), H.p(aln["read_reverse"]), H.p(rmq), C.c_double(100.0), C.c_double(10.0), out.encode(), C.c_int(1),
                                   C.c_int(rank), C.c_int(world), allreduce, H.p(q), H.p(ps))
    if fail_rank >= 0:   # a locus that fails on one rank before its device stage: every rank returns an error, none is left waiting in the collective
        msg = lib.typing_host_last_error().decode()
        assert n != 17 and calls[0] == 17, (n, calls[0], msg)
        assert ("test hook" in msg) if rank == fail_rank else ("another rank" in msg), msg
        open(os.path.join(out_root, "failed_%d" % rank), "w").write(msg)
        dist.barrier(); dist.destroy_process_group()
        return
    assert n == 17, lib.typing_host_last_error().decode()
    assert calls[0] == 17
    if rank == 0:
        np.save(os.path.join(out_root, "q.npy"), np.concatenate([q, ps]))
    dist.barrier(); dist.destroy_process_group()


def test_two_rank_typing_allreduce_matches_single_process(dataset, tmp_path):
    import ctypes as C
    d, b, mu, sd = dataset("typing")
    aln = H.Oracle(d).pairs(b, mu, sd, 512); rmq = np.ascontiguousarray(aln["read_mapq"], np.float64)
    lib = C.CDLL(os.path.join(H.REPO, "tests", "native", "build", "libtyping_host.so")); lib.typing_host_last_error.restype = C.c_char_p
    single = str(tmp_path / "single" / "hla"); os.makedirs(single)
    q1 = np.zeros(34, np.float64); ps1 = np.zeros(17, np.float64)
    n = lib.typing_host_run_ranked(d.encode(), C.c_longlong(len(b["read_off"]) - 1), H.p(b["read_off"]), H.p(b["bases"]), H.p(b["quals"]), C.c_int(512), H.p(aln["n_cols"]), H.p(aln["level"]),
                                   H.p(aln["gchar"]), H.p(aln["schar"]), H.p(aln["mapq"]), H.p(aln["read_reverse"]), H.p(rmq), C.c_double(mu), C.c_double(sd), single.encode(), C.c_int(0),
                                   C.c_int(0), C.c_int(1), None, H.p(q1), H.p(ps1))
    assert n == 17, lib.typing_host_last_error().decode()
    multi = str(tmp_path / "multi"); os.makedirs(multi)
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_typing_worker, args=(2, d, port, multi), nprocs=2, join=True)
    got = np.load(os.path.join(multi, "q.npy"))
    assert np.allclose(got[34:], ps1, rtol=1e-12, atol=1e-9), "allele-pair log-likelihood sums after the all-reduce"
    assert np.allclose(got[:34], q1, rtol=1e-9, atol=1e-12)
    assert open(os.path.join(multi, "hla", "R1_bestguess.txt")).read() == open(os.path.join(single, "R1_bestguess.txt")).read()
    assert sorted(os.listdir(os.path.join(multi, "hla"))) == sorted(os.listdir(single))


def test_two_rank_typing_fails_together(dataset, tmp_path):
    d, b, mu, sd = dataset("typing")
    out = str(tmp_path / "multi"); os.makedirs(out)
    port = 33500 + (os.getpid() % 2000)
    mp.spawn(_typing_worker, args=(2, d, port, out, 1), nprocs=2, join=True)
    assert os.path.exists(os.path.join(out, "failed_0")) and os.path.exists(os.path.join(out, "failed_1"))
