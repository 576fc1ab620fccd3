"""Multi-GPU plumbing of the alignment path (one process per GPU, torch.distributed).

Read pairs are independent through the whole alignment path, so ranks take contiguous blocks of the pair list and
the only exchange is the sum of the per-level coverage histogram (reads_per_level.txt is written from it, processBAM.cpp:1902-1913)
plus a few counters. No data-path collective is invented: the graph is replicated on every GPU."""
import numpy as np

BATCH_KEYS = ["read_off", "bases", "quals", "chain_off", "chain_contig", "chain_pos", "chain_flag", "chain_as", "cigar_off", "cigar"]


def shard_bounds(n_pairs, rank, world):
    """Contiguous block of pairs for `rank`: sizes differ by at most one, order is preserved (rank 0 gets the first block)."""
    base, rem = divmod(n_pairs, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(b, rank, world):
    """Slice a seed batch (dict of the arrays in BATCH_KEYS) to the pairs of `rank`; offsets are rebased to zero."""
    n_pairs = (len(b["read_off"]) - 1) // 2
    p0, p1 = shard_bounds(n_pairs, rank, world)
    r0, r1 = 2 * p0, 2 * p1
    c0, c1 = int(b["chain_off"][r0]), int(b["chain_off"][r1])
    g0, g1 = int(b["cigar_off"][c0]), int(b["cigar_off"][c1])
    q0, q1 = int(b["read_off"][r0]), int(b["read_off"][r1])
    out = {
        "read_off": (b["read_off"][r0:r1 + 1] - q0).copy(), "bases": b["bases"][q0:q1].copy(), "quals": b["quals"][q0:q1].copy(),
        "chain_off": (b["chain_off"][r0:r1 + 1] - c0).copy(),
        "cigar_off": (b["cigar_off"][c0:c1 + 1] - g0).copy(), "cigar": b["cigar"][g0:g1].copy(),
    }
    for k in ("chain_contig", "chain_pos", "chain_flag", "chain_as"):
        out[k] = b[k][c0:c1].copy()
    return out


def allreduce_coverage(cov, dist=None):
    """Sum the per-level coverage over ranks. `cov` is a torch tensor (CUDA with the nccl backend, CPU with gloo)."""
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(cov)
    return cov
