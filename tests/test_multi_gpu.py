"""Multi-GPU tests (need >= 2 CUDA devices; skipped otherwise): dataset-wide k-means over row
shards (BASELINE configs[4]) against the single-process oracle, with the exchange inside the
iterate kernel over peer-mapped memory (NVLink) and with the NCCL all-reduce of the totals
buffer between reduce and update."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip('torch')

pytestmark = pytest.mark.gpu

CASES = {
    # name: (N, D, K, centroid columns?)
    'descriptors': (6000, 514, 4, True),       # a handful of chunks per rank: one-level reduction
    'many_chunks': (60000, 18, 4, False),      # > 32 chunks per rank: two-level reduction tree
}


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problem(name):
    N, D, K, pos = CASES[name]
    rs = np.random.RandomState(0)
    cent = rs.standard_normal((K, D)) * (4 if pos else 0.7)
    X = (cent[rs.randint(0, K, N)] + rs.standard_normal((N, D))).astype(np.float32)
    if pos:
        X[:, -2] = rs.uniform(0, 1023, N)
        X[:, -1] = rs.uniform(0, 2047, N)
    w = rs.uniform(0, 1, N)
    return X, w, K


def _worker(rank, world, port, q, exchange, case):
    import time
    import torch.distributed as dist
    from oracle import spalign_oracle as so
    from superpixel_align_b200 import dist_kmeans, shard
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        X, w, K = _problem(case)
        lo, hi = shard.shard_range(len(X), world, rank)
        Xd, wd = torch.from_numpy(X[lo:hi]).to(dev), torch.from_numpy(w[lo:hi]).to(dev)
        msgs = []
        for rep in range(2):           # second run: warm, timed
            np.random.seed(1111)
            torch.cuda.synchronize()
            dist.barrier()
            t0 = time.time()
            res = dist_kmeans.global_kmeans(Xd, wd, K, exchange=exchange)
            torch.cuda.synchronize()
            dt = time.time() - t0
        np.random.seed(1111)
        want, info = so.kmeans(K, X.astype(np.float64), w, return_info=True, verbose=False)
        got = res.assign.cpu().numpy()
        why = []
        if not np.array_equal(got, np.asarray(want)[lo:hi].astype(np.int32)):
            why.append('%d assignments differ' % int((got != np.asarray(want)[lo:hi]).sum()))
        if res.iters[0].item() != info['iters'] or res.status[0].item() != info['status']:
            why.append('iters %d vs %d, status %d vs %d' % (res.iters[0].item(), info['iters'],
                                                          res.status[0].item(), info['status']))
        # centres are replicated: bit-identical on every rank (compare bits: NaN-safe)
        cen = res.centers.clone().view(torch.int64)
        ref = cen.clone()
        dist.broadcast(ref, 0)
        if not torch.equal(cen, ref):
            why.append('centres differ between ranks')
        ok = not why
        msgs.append('%s/%s world %d: %d iterations, %.1f ms wall incl. host init and communicator set-up' %
                    (case, exchange, world, info['iters'], dt * 1e3))
        q.put((rank, 'ok' if ok else '; '.join(why), msgs))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e), []))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('case', sorted(CASES))
@pytest.mark.parametrize('exchange', ['peer', 'nccl'])
@pytest.mark.parametrize('world', [2, 4, 8])
def test_global_kmeans_matches_oracle(world, exchange, case):
    if torch.cuda.device_count() < world:
        pytest.skip('needs %d GPUs' % world)
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, exchange, case))
             for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(60)
    for _, _, msgs in got:
        for m in msgs[:1]:
            print(m)
    assert {r: s for r, s, _ in got} == {r: 'ok' for r in range(world)}, got
