"""Region sharding across the GPUs of one box: every rank owns whole checkpoint blocks, nothing is
exchanged per site, and ONE all-reduce sums the per-shard totals (sum AN, sum AC, sum AC<M>, sites passed, sites).

One process per GPU; torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU tests) is plumbing only.
"""
import numpy as np


def shard_rows(n_rows, shift, rank, world):
    """Rows [beg, end) owned by `rank`: blocks [rank*ceil(B/world), (rank+1)*ceil(B/world)) of B = ceil(n/2^shift)."""
    bs = 1 << shift
    n_blk = (n_rows + bs - 1) // bs
    per = (n_blk + world - 1) // world
    b0 = min(rank * per, n_blk)
    b1 = min(b0 + per, n_blk)
    return min(b0 * bs, n_rows), min(b1 * bs, n_rows)


def allreduce_totals(totals, device=None):
    """Sum a list of int64 totals over all ranks (no-op without an initialised process group)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([int(x) for x in totals], dtype=torch.int64, device=device)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [int(x) for x in t.tolist()]


def sharded_scan(scan_rows, n_rows, shift, rank, world, device=None, gather=True):
    """Run `scan_rows(beg, end) -> dict(counts[n,stride], passed[n], totals[4])` on this rank's shard, all-reduce the
    totals and (optionally) gather the per-site results on rank 0 in row order.

    scan_rows is the only place compute happens: on the GPU box it wraps bgt_b200.scan on a Pbf loaded for the
    shard; the CPU tests inject a stand-in."""
    import torch.distributed as dist
    beg, end = shard_rows(n_rows, shift, rank, world)
    res = scan_rows(beg, end) if end > beg else dict(counts=None, passed=np.zeros(0, np.uint8), totals=[0, 0, 0, 0])
    totals = allreduce_totals(list(res["totals"]) + [end - beg], device=device)
    out = dict(rows=(beg, end), local=res, totals=totals)
    if gather and dist.is_available() and dist.is_initialized() and world > 1:
        parts = [None] * world if rank == 0 else None
        dist.gather_object((beg, end, res["counts"], res["passed"]), parts, dst=0)
        if rank == 0:
            parts = [p for p in parts if p[1] > p[0]]
            parts.sort(key=lambda p: p[0])
            out["counts"] = np.concatenate([p[2] for p in parts]) if parts else None
            out["passed"] = np.concatenate([p[3] for p in parts]) if parts else None
    elif gather:
        out["counts"], out["passed"] = res["counts"], res["passed"]
    return out
