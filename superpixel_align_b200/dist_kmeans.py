"""Dataset-wide prior-weighted k-means over row shards on several GPUs (BASELINE.json
configs[4]).  The reference has no such mode (direct_clustering.py clusters one batch); the
semantics are those of the reference kmeans() (batch_spalign_kmeans.py:136-183) applied to
the concatenation of all ranks' rows, rank r holding a contiguous slice.

Two exchange paths, same results:

* ``exchange='peer'`` (default): the iteration is ONE kernel launch per GPU.  The CTA that
  finishes the local reduction stores this rank's K*(D+2)+1 float64 sums (centroid sums, weight
  sums, member counts, #changed) into every rank's inbox over NVLink (peer-mapped memory, CUDA
  IPC), waits for the peers' flags and adds the vectors in rank order -- no NCCL call and no
  host round trip inside an iteration, Hamerly bounds and running sums stay in use.
* ``exchange='nccl'``: sweep -> reduce -> ``torch.distributed.all_reduce`` -> update, four
  enqueues per iteration, every rank polling the stop flag at fixed iterations.

Either way the update runs redundantly on every rank from bit-identical totals, so centres and
stop flags agree everywhere.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch
import torch.distributed as dist

from . import _lib


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


class PeerComm:
    """Peer-memory communicator (include/spalign.h, spalign_comm_*): one exchange buffer per
    rank, mapped into every other rank's process.  ``pv_cap`` = largest K*(D+2)+1 it will carry.
    All ranks of ``group`` must construct it together (the IPC handles travel through one
    ``all_gather_object``); ranks must live on one node with peer access (NVLink / NVSwitch)."""

    def __init__(self, pv_cap: int, group=None):
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.pv_cap = int(pv_cap)
        lib = _lib.load()
        h = ctypes.c_void_p()
        _lib.check(lib.spalign_comm_create(self.world, self.rank, self.pv_cap, ctypes.byref(h)),
                   'comm_create')
        self.handle = h
        buf = ctypes.create_string_buffer(_lib.COMM_HANDLE_BYTES)
        _lib.check(lib.spalign_comm_handle(self.handle, buf), 'comm_handle')
        handles = [None] * self.world
        dist.all_gather_object(handles, bytes(buf.raw), group=group)
        if self.world > 1:
            _lib.check(lib.spalign_comm_connect(self.handle, b''.join(handles)), 'comm_connect')
        dist.barrier(group=group)   # every inbox is mapped everywhere before anyone writes

    def close(self):
        if self.handle is not None:
            _lib.load().spalign_comm_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _timed_run(km, poll, events):
    if events is None:
        return km.run(poll=poll)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res = km.run(poll=poll)
    e1.record()
    events[:] = [e0, e1]
    return res


def global_kmeans(X_local: torch.Tensor, w_local: torch.Tensor, k: int, n_iter: int = 1000,
                  init_local=None, pos_grid=None, group=None, poll: int = 4,
                  exchange: str = 'peer', comm: PeerComm = None, row0: int = None,
                  events: list = None):
    """Cluster the union of all ranks' rows.  Returns this rank's KMeansResult (assignments of
    its rows; iters/status/centres identical on every rank).  ``comm``: a PeerComm to reuse
    (``exchange='peer'``); one is created and closed around the call otherwise.  ``row0``: global
    index of this rank's first row (gathered when omitted; only the virtual position columns
    need it).  ``events``: receives two CUDA events bracketing the device work (init sums and
    all iterations) for benchmarks."""
    from . import ops
    if row0 is None:
        sizes, row0 = gather_sizes(X_local.shape[0], group)
    if init_local is None:
        init_local = distributed_init(w_local.detach().cpu().numpy(), k, group)
    init_t = torch.as_tensor(np.asarray(init_local, dtype=np.int32), device=X_local.device)
    D = X_local.shape[1] + (2 if pos_grid else 0)
    if exchange == 'nccl':
        km = ops.KMeansLarge(X_local, w_local, init_t, k, [0, X_local.shape[0]], n_iter=n_iter,
                             pos_grid=pos_grid, pos_row0=row0,
                             allreduce=lambda t: allreduce_totals(t, group))
        return _timed_run(km, poll, events)
    if exchange != 'peer':
        raise ValueError("exchange must be 'peer' or 'nccl'")
    own = comm is None
    if own:
        comm = PeerComm(k * (D + 2) + 1, group)
    try:
        km = ops.KMeansLarge(X_local, w_local, init_t, k, [0, X_local.shape[0]], n_iter=n_iter,
                             pos_grid=pos_grid, pos_row0=row0, comm=comm)
        res = _timed_run(km, poll, events)
        if own:
            torch.cuda.synchronize()
        if (res.status == _lib.KM_COMM_TIMEOUT).any().item():
            raise _lib.SpalignError('global_kmeans: a peer never delivered its partial sums')
        return res
    finally:
        if own:
            comm.close()
