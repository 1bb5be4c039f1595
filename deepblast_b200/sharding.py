"""Batch sharding across GPUs.  The soft-DP path has no cross-pair term
(deepblast/nw.py:110-115, nw_cuda.py:76-79), so ranks own disjoint slices of the
batch and no collective touches the DP; NCCL only gathers the scalar result."""
import numpy as np


def shard_range(B, world, rank):
    """Plain slicing: rank r owns pairs [lo, hi); sizes differ by at most one."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def lpt_assign(cells, world):
    """Longest-processing-time assignment of variable-size pairs to ranks: sort by
    n*m descending, give each pair to the currently lightest rank.  Returns a list of
    index arrays (one per rank), each in descending-work order so a rank's persistent
    CTAs start with its largest lattices."""
    cells = np.asarray(cells, dtype=np.int64)
    order = np.argsort(-cells, kind="stable")
    load = np.zeros(world, dtype=np.int64)
    out = [[] for _ in range(world)]
    for idx in order:
        r = int(np.argmin(load))
        out[r].append(int(idx))
        load[r] += cells[idx]
    return [np.asarray(o, dtype=np.int64) for o in out]


def packing_stats(xlen, ylen, assignment):
    """useful cells, swept cells (strip-granular: rows rounded up to 32), imbalance."""
    xlen = np.asarray(xlen, dtype=np.int64)
    ylen = np.asarray(ylen, dtype=np.int64)
    useful = xlen * ylen
    swept = ((xlen + 31) // 32) * 32 * ylen
    per_rank = np.array([useful[a].sum() for a in assignment], dtype=np.float64)
    return dict(useful=int(useful.sum()), swept=int(swept.sum()),
                packing_efficiency=float(useful.sum() / max(1, swept.sum())),
                imbalance=float(per_rank.max() / max(1.0, per_rank.mean())))
