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
