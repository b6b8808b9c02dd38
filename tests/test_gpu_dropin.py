"""GPU tests of the reference-facing drop-in modules: they are called the way the reference's
estimate_road_mask calls its own functions (NumPy in, NumPy out, an argparse-like ``args``)
and compared with the oracle / the golden vectors frozen from the reference."""
import os
import types

import numpy as np
import pytest

torch = pytest.importorskip('torch')

from oracle import spalign_oracle as so  # noqa: E402
from superpixel_align_b200 import synth  # noqa: E402

pytestmark = pytest.mark.gpu


def _args(**kw):
    base = dict(gpu=0, n_clusters=4, n_anchors=10, n_neighbors=4, without_pos=False, y_rel_pos=0.75,
                x_rel_pos=0.5, y_rel_sigma=0.1, x_rel_sigma=0.1, overlap_threshold=0.01)
    base.update(kw)
    return types.SimpleNamespace(**base)


@pytest.fixture()
def batch():
    H, W, fh, fw, C = 128, 256, 16, 32, 32
    labs = np.stack([synth.voronoi_labels(H, W, 6, 10, image_index=i, dtype=np.int64) for i in range(3)])
    feats = np.stack([synth.smooth_features(C, fh, fw, seed=i) for i in range(3)])
    imgs = np.zeros((3, 3, H, W), dtype=np.float32)
    return labs, feats, imgs


def test_spalign_dropin_matches_oracle_end_to_end(batch):
    from superpixel_align_b200 import batch_spalign_kmeans as bsk
    labs, feats, imgs = batch
    args = _args()
    bsk.clear_cache()
    f, n_per = bsk.batch_superpixel_align(args, None, imgs, labs, feats)
    assert n_per == [60, 60, 60] and f.shape == (180, 34) and f.dtype == np.float64
    w = bsk.batch_create_prior(args, labs)
    assert w.shape == (180,) and w.dtype == np.float64
    np.random.seed(1111)
    cres, road = bsk.batch_weighted_kmeans(args, labs, f, w, n_per)
    assert cres.shape == labs.shape and cres.dtype == labs.dtype and road.dtype == bool
    # oracle on the same inputs
    of, ow = [], []
    for i in range(3):
        ip, ix, ct = so.overlap_csr(labs[i], 16, 32, 60)
        area, sy, sx = so.superpixel_stats(labs[i], 60)
        of.append(so.pool_count(ip, ix, ct, feats[i].reshape(32, -1).T, area, sy, sx, True))
        ow.append(so.create_prior(labs[i], 0.75, 0.5, 0.1, 0.1))
    of, ow = np.concatenate(of), np.concatenate(ow)
    np.testing.assert_allclose(f, of, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(w, ow, rtol=1e-12)
    np.random.seed(1111)
    oa = so.kmeans(4, f, w, verbose=False)        # same seeded stream, same X
    ocm, orm = so.weighted_kmeans_paint(labs, oa, n_per)
    assert np.array_equal(cres, ocm) and np.array_equal(road, orm)


def test_kmeans_dropin_reference_golden_and_seed_parity(golden_dir):
    from superpixel_align_b200 import batch_spalign_kmeans as bsk
    g = np.load(os.path.join(golden_dir, 'kmeans_ref.npz'))
    for name in ('k4', 'k2', 'k8', 'k3_pos', 'nan_center'):
        X, w, k = g[name + '__X'].astype(np.float64), g[name + '__w'], int(g[name + '__k'])
        got = bsk.kmeans(k, X, w, init_assign=g[name + '__init'])
        assert np.array_equal(np.asarray(got).astype(np.int32), g[name + '__assign']), name
    # the golden vectors were produced by ONE seeded stream in this order: replaying the
    # stream through the drop-in reproduces the reference without passing the init
    np.random.seed(1111)
    for name in ('k4', 'k2', 'k8', 'k3_pos'):
        X, w, k = g[name + '__X'].astype(np.float64), g[name + '__w'], int(g[name + '__k'])
        got = bsk.kmeans(k, X, w)
        assert np.array_equal(np.asarray(got).astype(np.int32), g[name + '__assign']), name


def test_kmeans_dropin_float64_rows_not_fp32_representable():
    from superpixel_align_b200 import batch_spalign_kmeans as bsk
    rs = np.random.RandomState(4)
    X = np.concatenate([rs.standard_normal((80, 9)) + 4 * rs.standard_normal((1, 9)) for _ in range(4)])
    w = rs.uniform(0, 1, len(X))
    np.random.seed(5)
    want = so.kmeans(4, X, w, verbose=False)
    np.random.seed(5)
    got = bsk.kmeans(4, X, w)
    assert np.array_equal(np.asarray(want).astype(np.int32), np.asarray(got).astype(np.int32))


def test_weighted_kmeans_dropin_reference_golden(golden_dir):
    from superpixel_align_b200 import batch_spalign_kmeans as bsk
    g = np.load(os.path.join(golden_dir, 'weighted_kmeans_ref.npz'))
    cres, road = bsk.weighted_kmeans(g['labs'], g['anchor_features'], g['weights'], int(g['k']),
                                     list(g['n_per']), init_assign=g['init'])
    assert cres.dtype == g['labs'].dtype
    assert np.array_equal(cres, g['cluster_map']) and np.array_equal(road, g['road'])


def test_prior_dropin_reference_golden(golden_dir):
    from superpixel_align_b200 import batch_spalign_kmeans as bsk
    from superpixel_align_b200 import direct_clustering as dc
    g = np.load(os.path.join(golden_dir, 'prior_ref.npz'))
    bsk.clear_cache()
    np.testing.assert_allclose(bsk.create_prior(g['lab_a'], 0.75, 0.5, 0.1, 0.1), g['w_a'], rtol=1e-12)
    np.testing.assert_allclose(bsk.create_prior(g['lab_b'], 0.6, 0.4, 0.2, 0.15), g['w_b'], rtol=1e-12)
    assert np.array_equal(dc.create_prior(28, 28, 0.75, 0.5, 0.1, 0.1), g['cell_28'])
    assert np.array_equal(dc.create_prior(16, 32, 0.75, 0.5, 0.1, 0.1), g['cell_16x32'])


def test_torch_carriers_stay_on_device(batch):
    from superpixel_align_b200 import batch_spalign_kmeans as bsk
    labs, feats, imgs = batch
    args = _args(without_pos=True)
    d = torch.device('cuda', 0)
    bsk.clear_cache()
    lt = torch.from_numpy(labs.astype(np.int32)).to(d)
    ft = torch.from_numpy(feats).to(d).contiguous(memory_format=torch.channels_last)
    f, n_per = bsk.batch_superpixel_align(args, None, None, lt, ft)
    w = bsk.batch_create_prior(args, lt)
    assert f.is_cuda and w.is_cuda and f.shape == (180, 32)
    np.random.seed(3)
    cres, road = bsk.batch_weighted_kmeans(args, lt, f, w, n_per)
    assert cres.is_cuda and cres.dtype == torch.int32 and road.dtype == torch.bool
    np.random.seed(3)
    oa = so.kmeans(4, f.cpu().numpy().astype(np.float64), w.cpu().numpy(), verbose=False)
    ocm, _ = so.weighted_kmeans_paint(labs, oa, n_per)
    assert np.array_equal(cres.cpu().numpy(), ocm.astype(np.int32))


def test_direct_clustering_dropin(batch):
    from superpixel_align_b200 import direct_clustering as dc
    labs, feats, imgs = batch
    args = _args()
    X = so.direct_features(feats)
    prior = so.create_prior_map(16, 32, 0.75, 0.5, 0.1, 0.1).reshape(1, -1).repeat(3, axis=0).reshape(-1)
    np.random.seed(8)
    want = so.kmeans(4, X, prior, verbose=False)
    np.random.seed(8)
    cres, road = dc.estimate_road_mask(feats, args)          # virtual (x, y) columns, no matrix
    assert np.array_equal(cres.reshape(-1), np.asarray(want).astype(np.int32))
    assert np.array_equal(road, cres == 0)
    np.random.seed(8)
    got2 = dc.batch_weighted_kmeans(args, X, prior)           # materialised matrix, as the reference
    assert np.array_equal(np.asarray(got2).astype(np.int32), np.asarray(want).astype(np.int32))


def test_superpixel_overlaps_dropin(batch):
    from superpixel_align_b200 import superpixel_overlaps as spo
    labs, feats, imgs = batch
    args = _args(overlap_threshold=0.02)
    np.random.seed(8)
    refined, cres, road = spo.estimate_road_mask(feats, labs, args)
    assert refined.shape == labs.shape and refined.dtype == np.uint8
    for i in range(3):
        want = so.refine_overlaps_masks(labs[i], so.upsample_nearest(road[i], 128, 256), 0.02)
        assert np.array_equal(refined[i], want)


def test_product_has_no_cpu_fallback():
    from superpixel_align_b200 import _lib, ops
    with pytest.raises(_lib.SpalignError):
        ops.label_max(torch.zeros((1, 8, 8), dtype=torch.int32))   # CPU tensor


def test_host_pipeline_streams_batches_and_matches_device_path():
    from superpixel_align_b200 import ops, pipeline
    d = torch.device('cuda', 0)
    H, W, fh, fw, C = 128, 256, 16, 32, 32
    labs = np.stack([synth.voronoi_labels(H, W, 6, 10, image_index=i) for i in range(5)])
    feats = np.stack([synth.smooth_features(C, fh, fw, seed=i).reshape(C, -1).T for i in range(5)])
    h_lab = torch.from_numpy(labs).pin_memory()
    h_feat = torch.from_numpy(np.ascontiguousarray(feats)).pin_memory()
    batches = [(h_lab[0:2], h_feat[0:2], [60, 60]), (h_lab[2:4], h_feat[2:4], [60, 60]),
               (h_lab[4:5], h_feat[4:5], [60])]
    got = {}
    hp = pipeline.HostPipeline(H, W, fh, fw, C, sub_batch=2, k=4)
    np.random.seed(7)
    hp.process(batches, lambda i, c, m: got.__setitem__(i, (c.clone(), m.clone())))
    assert sorted(got) == [0, 1, 2] and hp.h2d_bytes == labs.nbytes + feats.nbytes
    np.random.seed(7)
    for i, (l, f, n_sp) in enumerate(batches):
        ref = pipeline.run_batch(l.to(d), f.to(d), n_sp, fh, fw, k=4)
        assert torch.equal(ref.cluster_map.cpu(), got[i][0])
        assert torch.equal(ref.road_mask.cpu(), got[i][1])


def test_host_pipeline_reports_label_problems_with_the_result():
    """The K1 condition words travel back with every sub-batch: a label outside [0, n_sp) or ids
    with gaps raise at the point where the result would have been handed over."""
    from superpixel_align_b200 import pipeline
    H, W, fh, fw, C = 128, 256, 16, 32, 32
    labs = np.stack([synth.voronoi_labels(H, W, 6, 10, image_index=i) for i in range(4)])
    feats = np.stack([synth.smooth_features(C, fh, fw, seed=i).reshape(C, -1).T for i in range(4)])
    h_feat = torch.from_numpy(np.ascontiguousarray(feats)).pin_memory()
    bad = labs.copy()
    bad[3, 5, 7] = 1000                       # out of range in the second sub-batch
    h_bad = torch.from_numpy(bad).pin_memory()
    hp = pipeline.HostPipeline(H, W, fh, fw, C, sub_batch=2, k=4)
    seen = []
    np.random.seed(7)
    with pytest.raises(ValueError, match='sub-batch 1'):
        hp.process([(h_bad[0:2], h_feat[0:2], [60, 60]), (h_bad[2:4], h_feat[2:4], [60, 60])],
                   lambda i, c, m: seen.append(i))
    assert seen == [0]
    gap = labs.copy()
    gap[gap == 59] = 60                       # id 59 has no pixels: 0/0 weights downstream
    h_gap = torch.from_numpy(gap).pin_memory()
    np.random.seed(7)
    with pytest.raises(ValueError, match='not contiguous'):
        hp.process([(h_gap[0:2], h_feat[0:2], [61, 61])], None)


def test_staged_transfers_move_every_byte():
    """The pinned-ring transfers of the NumPy drop-in (chunks of 32 MB, ragged tail, ring reuse):
    identical bytes both ways, for the dtypes the drop-in moves."""
    from superpixel_align_b200 import batch_spalign_kmeans as bsk
    d = torch.device('cuda', 0)
    rs = np.random.RandomState(0)
    for dtype, n in ((np.float32, (5 * bsk._STAGE_CHUNK + 12345) // 4), (np.int64, (bsk._STAGE_CHUNK * 6) // 8 + 7),
                     (np.uint8, bsk._STAGE_CHUNK + 1)):
        a = rs.randint(0, 250, size=n).astype(dtype)
        t = bsk._h2d_staged(a, d)
        assert t.shape == a.shape and torch.equal(t.cpu(), torch.from_numpy(a))
        back = bsk._d2h_staged(t)
        assert back.dtype == a.dtype and np.array_equal(back, a)
    big = np.zeros(bsk._STAGE_MIN // 4 + 3, np.float32)
    big[-1] = 7.0
    assert float(bsk._to_dev(big, d)[-1]) == 7.0 and bsk._to_host(bsk._to_dev(big, d))[-1] == 7.0


def test_run_batch_overlapped_on_side_streams_equals_one_batch():
    from superpixel_align_b200 import pipeline
    d = torch.device('cuda', 0)
    H, W, fh, fw, C = 128, 256, 16, 32, 32
    n = 7
    labs = torch.from_numpy(np.stack([synth.voronoi_labels(H, W, 6, 10, image_index=i) for i in range(n)])).to(d)
    feats = torch.from_numpy(np.ascontiguousarray(np.stack(
        [synth.smooth_features(C, fh, fw, seed=i).reshape(C, -1).T for i in range(n)]))).to(d)
    np.random.seed(3)
    ref = pipeline.run_batch(labs, feats, [60] * n, fh, fw, k=4)
    for sb, ns in ((3, 2), (2, 3), (7, 2)):
        np.random.seed(3)
        out = pipeline.run_batch_overlapped(labs, feats, [60] * n, fh, fw, sub_batch=sb, n_streams=ns, k=4)
        torch.cuda.synchronize()
        assert torch.equal(out.cluster_map, ref.cluster_map) and torch.equal(out.road_mask, ref.road_mask)
        assert torch.equal(out.iters, ref.iters) and torch.equal(out.assign, ref.assign)
        assert len(out.parts) == -(-n // sb)


def test_run_batch_paint_overlap_equals_sequential():
    """Ranges of groups finishing on separate streams with their paint-back overlapped: same
    cluster maps, masks, iteration counts and stop reasons as the one-stream form (per-image and
    joint groups; ragged last range)."""
    from superpixel_align_b200 import pipeline
    d = torch.device('cuda', 0)
    H, W, fh, fw, C = 128, 256, 16, 32, 32
    n = 13
    labs = torch.from_numpy(np.stack([synth.voronoi_labels(H, W, 6, 10, image_index=i) for i in range(n)])).to(d)
    feats = torch.from_numpy(np.ascontiguousarray(np.stack(
        [synth.smooth_features(C, fh, fw, seed=i).reshape(C, -1).T for i in range(n)]))).to(d)
    for ipg in (1, 2):
        np.random.seed(5)
        ref = pipeline.run_batch(labs, feats, [60] * n, fh, fw, k=4, images_per_group=ipg)
        for po in (2, 3):
            np.random.seed(5)
            out = pipeline.run_batch(labs, feats, [60] * n, fh, fw, k=4, images_per_group=ipg,
                                     paint_overlap=po)
            torch.cuda.synchronize()
            assert torch.equal(out.cluster_map, ref.cluster_map)
            assert torch.equal(out.road_mask, ref.road_mask)
            assert torch.equal(out.iters, ref.iters) and torch.equal(out.status, ref.status)
            assert torch.equal(out.assign, ref.assign)


def test_results_scores_and_artefacts(tmp_path):
    from superpixel_align_b200 import results
    rs = np.random.RandomState(1)
    cityscapes = rs.randint(0, 12, size=(2, 40, 60)).astype(np.uint8)
    gt = np.stack([results.create_label_mask(c) for c in cityscapes])
    assert set(np.unique(gt)) == {-1, 0, 1}
    road = rs.rand(2, 40, 60) < 0.4
    scores = results.road_scores(road, gt)
    for i in range(2):
        iou, prec, rec, tp, fp, fn = so.road_iou(road[i], gt[i])
        assert (scores[i]['TP'], scores[i]['FP'], scores[i]['FN']) == (tp, fp, fn)
        assert scores[i]['road_iou'] == pytest.approx(iou) and scores[i]['precision'] == pytest.approx(prec)
    args = types.SimpleNamespace(n_clusters=4, batchsize=30)
    info = results.save_info(str(tmp_path), 'a/b/frankfurt_000000_000294_leftImg8bit.png', 'lab.png',
                             road[0], road[0].astype(np.int64) * 2, scores[0], args, {'time_kmeans': 0.1}, None)
    assert np.array_equal(np.load(tmp_path / 'frankfurt_000000_000294_leftImg8bit.npy'), road[0].astype(np.uint8))
    assert np.load(tmp_path / 'frankfurt_000000_000294_leftImg8bit_all_cluster.npy').dtype == np.uint8
    import json
    rec = json.loads(open(tmp_path / 'result.json').read().strip())
    assert rec['road_iou'] == info['road_iou'] and rec['n_clusters'] == 4 and rec['time_kmeans'] == 0.1


def test_superpixel_align_dropin_bilinear_mode(batch):
    from superpixel_align_b200 import batch_spalign_kmeans as bsk
    labs, feats, imgs = batch
    bsk.clear_cache()
    got = bsk.superpixel_align(imgs[0], feats[0], labs[0], 10, 4, False, pooling='bilinear')
    want = so.pool_dense_bilinear(labs[0], feats[0])
    assert got.dtype == np.float32
    np.testing.assert_allclose(got, want.astype(np.float32), rtol=1e-5, atol=1e-6)


# ----------------------------------------------------- inline reference code, frozen (round 2)
def test_refine_dropin_reference_inline_golden(golden_dir):
    """CUDA K1 counts + K5 refine + K4 paint against the reference's own refine loop
    (superpixel_overlaps.py:360-369, run unmodified by oracle/gen_golden.py)."""
    from superpixel_align_b200 import batch_spalign_kmeans as bsk
    from superpixel_align_b200 import superpixel_overlaps as spo
    g = np.load(os.path.join(golden_dir, 'refine_ref.npz'))
    for name in ('s8', 'ratio', 'r224'):
        lab, road = g[name + '__label'], g[name + '__road_cell']
        for thr in (0.01, 0.05, 0.2):
            bsk.clear_cache()
            got = spo.refine_road_masks(lab[None], road[None], thr)
            assert np.array_equal(got[0], g['%s__refined_%g' % (name, thr)]), (name, thr)
    got = spo.refine_road_masks(g['noroad__label'][None], np.zeros((1, 8, 12), bool), 0.01)
    assert not got.any()


def test_direct_clustering_dropin_reference_inline_golden(golden_dir):
    from superpixel_align_b200 import direct_clustering as dc
    g = np.load(os.path.join(golden_dir, 'direct_features_ref.npz'))
    n, h, w = int(g['n']), int(g['h']), int(g['w'])
    # virtual (x, y) columns in the kernel == the matrix the reference materialises (:298-303)
    got = dc.cluster_cells(g['feats'], g['prior'], 4, init_assign=g['init'])
    assert np.array_equal(got.reshape(-1), g['assign'])
    # and the materialised matrix through the reference-named entry point
    got2 = dc.kmeans(4, g['X'], g['prior'], init_assign=g['init'])
    assert np.array_equal(np.asarray(got2).astype(np.int32), g['assign'])
    args = _args()
    res, road = dc.estimate_road_mask(g['feats'], args)      # seeded init drawn inside
    assert res.shape == (n, h, w) and road.dtype == bool


def test_gapped_label_ids_follow_the_reference(golden_dir):
    """1-based ids with a missing id: rows = sorted unique ids, paint-back by enumerate index."""
    from superpixel_align_b200 import batch_spalign_kmeans as bsk
    g = np.load(os.path.join(golden_dir, 'gapped_ref.npz'))
    labs, n_per = g['labs'], [int(v) for v in g['n_per']]
    args = _args(n_clusters=int(g['k']))
    bsk.clear_cache()
    w = bsk.batch_create_prior(args, labs)
    np.testing.assert_allclose(w, g['weights'], rtol=1e-12)
    feats = np.zeros((2, 8, 4, 8), dtype=np.float32)
    f, got_n = bsk.batch_superpixel_align(args, None, None, labs, feats)
    assert got_n == n_per and f.shape == (sum(n_per), 10)
    cres, road = bsk.weighted_kmeans(labs, g['feats'], g['weights'], int(g['k']), n_per,
                                     init_assign=g['init'])
    assert np.array_equal(cres, g['cluster_map']) and np.array_equal(road, g['road'])


def test_state_cache_never_returns_a_stale_batch(batch):
    """A preallocated buffer refilled in place (NumPy ``buf[:] = ...``, torch ``copy_``) must not
    be mistaken for the previous batch; lists are never cached; the explicit handle works."""
    from superpixel_align_b200 import batch_spalign_kmeans as bsk
    labs, feats, imgs = batch
    args = _args()
    other = np.stack([synth.voronoi_labels(128, 256, 6, 10, image_index=50 + i, dtype=np.int64)
                      for i in range(3)])
    want_a = np.concatenate([so.create_prior(l, 0.75, 0.5, 0.1, 0.1) for l in labs])
    want_b = np.concatenate([so.create_prior(l, 0.75, 0.5, 0.1, 0.1) for l in other])
    bsk.clear_cache()
    buf = labs.copy()
    np.testing.assert_allclose(bsk.batch_create_prior(args, buf), want_a, rtol=1e-12)
    buf[:] = other                                   # same address, same shape, new content
    np.testing.assert_allclose(bsk.batch_create_prior(args, buf), want_b, rtol=1e-12)
    d = torch.device('cuda', 0)
    tb = torch.from_numpy(labs).to(d)
    np.testing.assert_allclose(bsk.batch_create_prior(args, tb).cpu().numpy(), want_a, rtol=1e-12)
    tb.copy_(torch.from_numpy(other))
    np.testing.assert_allclose(bsk.batch_create_prior(args, tb).cpu().numpy(), want_b, rtol=1e-12)
    np.testing.assert_allclose(bsk.batch_create_prior(args, [l for l in labs]), want_a, rtol=1e-12)
    np.testing.assert_allclose(bsk.batch_create_prior(args, [l for l in other]), want_b, rtol=1e-12)
    # explicit handle: same results as the array route, labels uploaded once
    st = bsk.prepare_batch(args, labs, feature_shape=feats.shape[-2:])
    f1, n1 = bsk.batch_superpixel_align(args, None, imgs, st, feats)
    w1 = bsk.batch_create_prior(args, st)
    np.random.seed(1111)
    c1, r1 = bsk.batch_weighted_kmeans(args, st, f1, w1, n1)
    bsk.clear_cache()
    f2, n2 = bsk.batch_superpixel_align(args, None, imgs, labs, feats)
    w2 = bsk.batch_create_prior(args, labs)
    np.random.seed(1111)
    c2, r2 = bsk.batch_weighted_kmeans(args, labs, f2, w2, n2)
    assert n1 == n2 and np.array_equal(f1, f2) and np.array_equal(w1.cpu().numpy(), w2)
    assert np.array_equal(c1, c2) and np.array_equal(r1, r2)
    # a caller that edits the descriptors in place gets them clustered, not the cached ones
    f3 = f2.copy()
    f3[:, :-2] *= -1.0
    np.random.seed(1111)
    c3, _ = bsk.batch_weighted_kmeans(args, labs, f3, w2, n2)
    np.random.seed(1111)
    oa = so.kmeans(4, f3, w2, verbose=False)
    ocm, _ = so.weighted_kmeans_paint(labs, oa, n2)
    assert np.array_equal(c3, ocm)


def test_dropin_centroid_columns_are_exact_float64(batch):
    from superpixel_align_b200 import batch_spalign_kmeans as bsk
    labs, feats, imgs = batch
    bsk.clear_cache()
    f, n_per = bsk.batch_superpixel_align(_args(), None, imgs, labs, feats)
    off = 0
    for i, n in enumerate(n_per):
        area, sy, sx = so.superpixel_stats(labs[i], n)
        assert np.array_equal(f[off:off + n, -2], sy / area) and np.array_equal(f[off:off + n, -1], sx / area)
        off += n
