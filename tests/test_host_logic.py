"""CPU tests of the host-side logic: reference sharding rule, batch ranges, and the
multi-process protocol of the dataset-wide k-means (gloo, world_size 2)."""
import os
import socket

import numpy as np
import pytest

from superpixel_align_b200 import shard


def test_shard_ranges_follow_reference_shell_rule():
    # utils/create_val_labels.sh: n_data=500, step = 500/N + 1
    assert shard.all_ranges(500, 1) == [(0, 500)]
    assert shard.all_ranges(500, 2) == [(0, 251), (251, 500)]
    assert shard.all_ranges(500, 4) == [(0, 126), (126, 252), (252, 378), (378, 500)]
    r8 = shard.all_ranges(500, 8)
    assert r8[0] == (0, 63) and r8[-1] == (441, 500) and len(r8) == 8
    for n, N in [(500, 8), (300, 7), (19998, 8), (5, 8), (2975, 25)]:
        rs = shard.all_ranges(n, N)
        assert rs[0][0] == 0 and rs[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(rs[:-1], rs[1:]))      # disjoint cover
        for r, rng in enumerate(rs):
            assert shard.shard_range(n, N, r) == rng
        for r in range(len(rs), N):                                     # ranks the shell never starts
            lo, hi = shard.shard_range(n, N, r)
            assert lo == hi


def test_batch_ranges_keep_batchsize_like_reference():
    assert list(shard.batch_ranges(0, 100, 30)) == [(0, 30), (30, 60), (60, 90), (70, 100)]
    assert list(shard.batch_ranges(0, 90, 30)) == [(0, 30), (30, 60), (60, 90)]
    assert list(shard.batch_ranges(10, 25, 30)) == [(0, 25)]
    assert list(shard.batch_ranges(0, 3, 1)) == [(0, 1), (1, 2), (2, 3)]


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    import torch
    from oracle import spalign_oracle as so
    from superpixel_align_b200 import dist_kmeans
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        rs = np.random.RandomState(0)
        N, D, K = 301, 6, 4
        X = (rs.standard_normal((N, D)) + (5 * rs.standard_normal((4, D)))[rs.randint(0, 4, N)]).astype(np.float32)
        w = rs.uniform(0, 1, N)
        lo, hi = (0, 140) if rank == 0 else (140, N)          # uneven contiguous row shards
        sizes, row0 = dist_kmeans.gather_sizes(hi - lo)
        assert list(sizes) == [140, 161] and row0 == lo
        np.random.seed(1111)
        init_local = dist_kmeans.distributed_init(w[lo:hi], K)
        np.random.seed(1111)
        init_ref = so.kmeans_init(K, w).astype(np.int32)
        assert np.array_equal(init_local, init_ref[lo:hi])
        # protocol: local partial sums in the library's totals layout -> all-reduce -> update
        assign = init_ref.copy()
        cen = np.stack([X[assign == k].astype(np.float64).mean(0) for k in range(K)])
        pv = K * (D + 2) + 1
        for it in range(50):
            Xl, wl = X[lo:hi].astype(np.float64), w[lo:hi]
            d = np.linalg.norm(Xl[:, None] - cen[None], axis=2)
            new = d.argmin(1).astype(np.int32)
            tot = np.zeros(pv)
            for k in range(K):
                m = new == k
                om = wl[m] if k == 0 else 1 - wl[m]
                tot[k * (D + 2):k * (D + 2) + D] = (Xl[m] * om[:, None]).sum(0)
                tot[k * (D + 2) + D] = om.sum()
                tot[k * (D + 2) + D + 1] = m.sum()
            tot[-1] = (new != assign[lo:hi]).sum()
            t = torch.from_numpy(tot)
            dist_kmeans.allreduce_totals(t)
            tot = t.numpy()
            assign[lo:hi] = new
            if tot[-1] == 0:
                break
            with np.errstate(all='ignore'):
                cen = np.stack([tot[k * (D + 2):k * (D + 2) + D] / tot[k * (D + 2) + D] for k in range(K)])
            if any(tot[k * (D + 2) + D + 1] == 0 for k in range(K)):
                break                                          # empty cluster stop (:173-181)
        want = so.kmeans(K, X.astype(np.float64), w, init_assign=init_ref.astype(np.float64), verbose=False)
        assert np.array_equal(assign[lo:hi], np.asarray(want)[lo:hi].astype(np.int32))
        q.put((rank, 'ok'))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_global_kmeans_protocol_gloo_world2():
    mp = pytest.importorskip('torch.multiprocessing')
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(30)
    assert res == {0: 'ok', 1: 'ok'}, res


def test_label_archives_read_back_like_the_reference_dataset(tmp_path):
    """zip -0 of *leftImg8bit.npy (README.md:134-136) and np.savez label zips
    (utils/run_train_rounds.py:191-203), read the way
    datasets/zipped_estimated_cityscapes_dataset.py:20-24,65-66 reads them."""
    import zipfile
    from superpixel_align_b200 import results
    rs = np.random.RandomState(0)
    out_dir = tmp_path / 'results' / 'estimated_train_labels'
    out_dir.mkdir(parents=True)
    masks = {}
    for city, seq in (('aachen', '000000_000019'), ('bochum', '000000_000313')):
        stem = '%s_%s_leftImg8bit' % (city, seq)
        m = (rs.rand(16, 32) < 0.4).astype(np.uint8)
        masks[stem] = m
        np.save(out_dir / (stem + '.npy'), m)
        np.save(out_dir / (stem + '_all_cluster.npy'), m * 3)      # must NOT be archived
    zfn = str(tmp_path / 'estimated_train_labels.0.zip')
    names = results.zip_estimated_labels(str(out_dir), zfn, root=str(tmp_path))
    assert sorted(names) == sorted('results/estimated_train_labels/%s.npy' % s for s in masks)
    with zipfile.ZipFile(zfn) as zf:
        assert all(i.compress_type == zipfile.ZIP_STORED for i in zf.infolist())
        label_fns = {'_'.join(os.path.basename(fn).split('_')[:3]): fn
                     for fn in zf.namelist() if fn.endswith('leftImg8bit.npy')}
    loaded = np.load(zfn)
    for stem, m in masks.items():
        key = '_'.join(stem.split('_')[:3])
        assert np.array_equal(loaded[label_fns[key]].astype(np.int32), m)
    # the direct writer produces the same archive content
    zfn2 = str(tmp_path / 'direct.zip')
    with results.LabelZipWriter(zfn2) as w:
        for stem, m in masks.items():
            w.add('/data/leftImg8bit/train/x/%s.png' % stem, m)
    loaded2 = np.load(zfn2)
    for stem, m in masks.items():
        assert np.array_equal(loaded2['results/estimated_train_labels/%s.npy' % stem], m)
    # np.savez label archives
    d = {'a_leftImg8bit': masks[next(iter(masks))], 'a_leftImg8bit_scores': rs.rand(2, 4, 4).astype(np.float32)}
    zfn3 = results.save_label_npz(str(tmp_path / 'iter-10_eval-train.0.zip'), d)
    back = np.load(zfn3)
    assert set(back.files) == set(d) and all(np.array_equal(back[k], v) for k, v in d.items())
