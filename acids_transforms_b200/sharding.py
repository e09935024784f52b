"""Batch sharding over the GPUs of one box: clips are independent (every op on the path is per clip —
utils/misc.py:168-178 flattens batch dims into rows), so rank r takes a contiguous range of clips and the
forward / inverse paths need NO collective.  The only cross-clip quantity is Normalize's global statistics,
fitted once by `scale_data`; `merge_stats` / `all_reduce_stats` combine per-shard statistics exactly
(min, max, mean, unbiased std) when a dataset-level fit is wanted across ranks."""
import math
from typing import List, Sequence, Tuple

import torch


def shard_range(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous split of `total` clips: the first `total % world` ranks get one extra clip."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard(x: torch.Tensor, world: int, rank: int) -> torch.Tensor:
    lo, hi = shard_range(x.shape[0], world, rank)
    return x[lo:hi]


def merge_stats(stats: Sequence[Sequence[float]], counts: Sequence[int]) -> List[float]:
    """Combine per-shard [min, max, mean, unbiased std] into the statistics of the union (Chan et al.)."""
    n_tot, mean, m2 = 0, 0.0, 0.0
    mn, mx = math.inf, -math.inf
    for (a, b, mu, sd), n in zip(stats, counts):
        if n == 0:
            continue
        mn, mx = min(mn, a), max(mx, b)
        m2_i = sd * sd * (n - 1) if n > 1 else 0.0
        delta = mu - mean
        new_n = n_tot + n
        mean += delta * n / new_n
        m2 += m2_i + delta * delta * n_tot * n / new_n
        n_tot = new_n
    return [mn, mx, mean, math.sqrt(m2 / (n_tot - 1)) if n_tot > 1 else float("nan")]


def all_reduce_stats(st: torch.Tensor, count: int, group=None) -> torch.Tensor:
    """All ranks contribute their float64[4] statistics and element count; every rank gets the merged result."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    payload = torch.cat([st.to(torch.float64).reshape(4), torch.tensor([float(count)], dtype=torch.float64, device=st.device)])
    gathered = [torch.empty_like(payload) for _ in range(world)]
    dist.all_gather(gathered, payload, group=group)
    rows = [g.tolist() for g in gathered]
    merged = merge_stats([r[:4] for r in rows], [int(r[4]) for r in rows])
    return torch.tensor(merged, dtype=torch.float64, device=st.device)
