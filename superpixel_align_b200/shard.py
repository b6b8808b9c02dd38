"""Image-range sharding, exactly as the reference's shell drivers do it
(utils/create_val_labels.sh:38-52 and siblings): one process per GPU, disjoint index ranges,
no communication on the data path."""
from __future__ import annotations

from typing import Iterator, List, Tuple


def shard_range(n_data: int, n_shards: int, rank: int) -> Tuple[int, int]:
    """``step = n_data / N + 1``; shard r owns ``[r*step, min((r+1)*step, n_data))``."""
    step = n_data // n_shards + 1
    lo = min(rank * step, n_data)
    return lo, min(lo + step, n_data)


def all_ranges(n_data: int, n_shards: int) -> List[Tuple[int, int]]:
    """The ranges the shell loop actually spawns (it stops once i reaches n_data, so fewer
    than N processes may start when N is large)."""
    out, i, step = [], 0, n_data // n_shards + 1
    while i < n_data:
        out.append((i, min(i + step, n_data)))
        i += step
    return out


def batch_ranges(start: int, end: int, batchsize: int) -> Iterator[Tuple[int, int]]:
    """Batches inside one shard, batch_spalign_kmeans.py:538-544: the last batch is re-anchored
    to ``end - batchsize`` to keep the batch size (so it overlaps the previous one; a shard
    shorter than ``batchsize`` yields a start below ``start``, which we clamp at 0 where the
    reference would wrap around with a negative slice)."""
    for i in range(start, end, batchsize):
        if i + batchsize >= end:
            yield max(end - batchsize, 0), end
        else:
            yield i, i + batchsize
