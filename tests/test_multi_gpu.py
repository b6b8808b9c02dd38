"""Multi-GPU tests (need >= 2 CUDA devices; skipped otherwise): dataset-wide k-means over row
shards with the NCCL all-reduce of the totals buffer, against the single-process oracle."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip('torch')

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from oracle import spalign_oracle as so
    from superpixel_align_b200 import dist_kmeans, shard
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device('cuda', rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=dev)
    try:
        rs = np.random.RandomState(0)
        N, D, K = 6000, 514, 4
        cent = rs.standard_normal((K, D)) * 4
        X = (cent[rs.randint(0, K, N)] + rs.standard_normal((N, D))).astype(np.float32)
        X[:, -2] = rs.uniform(0, 1023, N)
        X[:, -1] = rs.uniform(0, 2047, N)
        w = rs.uniform(0, 1, N)
        lo, hi = shard.shard_range(N, world, rank)
        np.random.seed(1111)
        res = dist_kmeans.global_kmeans(torch.from_numpy(X[lo:hi]).to(dev),
                                        torch.from_numpy(w[lo:hi]).to(dev), K)
        np.random.seed(1111)
        want, info = so.kmeans(K, X.astype(np.float64), w, return_info=True, verbose=False)
        got = res.assign.cpu().numpy()
        ok = np.array_equal(got, np.asarray(want)[lo:hi].astype(np.int32)) and \
            res.iters[0].item() == info['iters'] and res.status[0].item() == info['status']
        q.put((rank, 'ok' if ok else 'mismatch iters %d vs %d' % (res.iters[0].item(), info['iters'])))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs 2 GPUs')
def test_global_kmeans_two_gpus_matches_oracle():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(60)
    assert res == {0: 'ok', 1: 'ok'}, res
