"""Host-side sharding of independent streams over ranks: no data-path collective.

Streams are independent work, so rank r simply owns a contiguous slice of the stream list; the
only cross-rank traffic is the barrier around the timed region and a MAX-reduce of elapsed times /
SUM of processed blocks for reporting (bench.py).  Contiguous slices keep each stream's blocks on
one GPU, which is what the ampmax chain (lib/block.c:626-628) and the decode overlap-add
(lib/block.c:767-823) need.
"""


def stream_slice(nstreams, world, rank):
    """[lo, hi) of the streams owned by `rank`; sizes differ by at most one."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, extra = divmod(nstreams, world)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def reduce_report(dist, local_blocks, local_ms, device=None):
    """(total blocks, max elapsed ms) over all ranks; `dist` is torch.distributed (initialised)."""
    import torch
    t = torch.tensor([float(local_blocks)], dtype=torch.float64, device=device)
    m = torch.tensor([float(local_ms)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    dist.all_reduce(m, op=dist.ReduceOp.MAX)
    return float(t.item()), float(m.item())
