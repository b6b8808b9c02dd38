"""Dataset-wide prior-weighted k-means over row shards on several GPUs (BASELINE.json
configs[4]).  The reference has no such mode (direct_clustering.py clusters one batch); the
semantics are those of the reference kmeans() (batch_spalign_kmeans.py:136-183) applied to
the concatenation of all ranks' rows, rank r holding a contiguous slice.

Per iteration every rank runs the multi-CTA sweep on its rows, sums its chunk partials in
fixed order, and the ranks exchange ONE buffer of K*(D+2)+1 float64 values (centroid sums,
weight sums, member counts, #changed) with an NCCL all-reduce; the update kernel then runs
redundantly on every rank, so centres and stop flags stay bit-identical across ranks.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def gather_sizes(n_local: int, group=None):
    """Row count of every rank -> (sizes [world], my offset)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [None] * world
    dist.all_gather_object(sizes, int(n_local), group=group)
    return np.asarray(sizes, dtype=np.int64), int(np.sum(sizes[:rank]))


def distributed_init(w_local: np.ndarray, k: int, group=None, seed_rank: int = 0) -> np.ndarray:
    """Seeded init of the GLOBAL problem (batch_spalign_kmeans.py:141-149): the upper median
    of all weights and one shuffle of the low-prior rows are global, so the weights are
    gathered on ``seed_rank``, the init is drawn there from ``np.random`` (seed-compatible
    with the single-process reference) and the slices are scattered back."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    gathered = [None] * world if rank == seed_rank else None
    dist.gather_object(np.asarray(w_local, dtype=np.float64), gathered, dst=seed_rank, group=group)
    parts = None
    if rank == seed_rank:
        sizes = [len(g) for g in gathered]
        w = np.concatenate(gathered)
        n = len(w)
        init = np.zeros(n, dtype=np.int32)
        thr = float(np.sort(w)[n // 2])
        low = w <= thr
        idx = np.arange(int(low.sum())) % (k - 1) + 1
        np.random.shuffle(idx)
        init[low] = idx
        parts = np.split(init, np.cumsum(sizes)[:-1])
    out = [None]
    dist.scatter_object_list(out, parts, src=seed_rank, group=group)
    return out[0]


def allreduce_totals(totals: torch.Tensor, group=None):
    """Sum the per-rank [G, K*(D+2)+1] float64 totals in place (NCCL on CUDA tensors)."""
    dist.all_reduce(totals, op=dist.ReduceOp.SUM, group=group)


def global_kmeans(X_local: torch.Tensor, w_local: torch.Tensor, k: int, n_iter: int = 1000,
                  init_local=None, pos_grid=None, group=None, poll: int = 4):
    """Cluster the union of all ranks' rows.  Returns this rank's KMeansResult (assignments of
    its rows; iters/status/centres identical on every rank)."""
    from . import ops
    sizes, row0 = gather_sizes(X_local.shape[0], group)
    if init_local is None:
        init_local = distributed_init(w_local.detach().cpu().numpy(), k, group)
    init_t = torch.as_tensor(np.asarray(init_local, dtype=np.int32), device=X_local.device)
    km = ops.KMeansLarge(X_local, w_local, init_t, k, [0, X_local.shape[0]], n_iter=n_iter,
                         pos_grid=pos_grid, pos_row0=row0,
                         allreduce=lambda t: allreduce_totals(t, group))
    return km.run(poll=poll)
