"""CPU tests: pin the oracle (oracle/spalign_oracle.py) against golden vectors frozen from
the unmodified reference functions (oracle/gen_golden.py), and, where /root/reference is
present, against the reference functions executed live."""
import os

import numpy as np
import pytest

from oracle import ref_extract
from oracle import spalign_oracle as so
from superpixel_align_b200 import synth


def _cases(npz):
    names = sorted({k.split('__')[0] for k in npz.files})
    return {n: {k.split('__')[1]: npz[k] for k in npz.files if k.startswith(n + '__')}
            for n in names}


def test_kmeans_matches_reference_golden(golden_dir):
    cases = _cases(np.load(os.path.join(golden_dir, 'kmeans_ref.npz')))
    assert set(cases) == {'k4', 'k2', 'k8', 'k3_pos', 'nan_center'}
    for name, c in cases.items():
        got = so.kmeans(int(c['k']), c['X'].astype(np.float64), c['w'],
                        init_assign=c['init'].astype(np.float64), verbose=False)
        assert np.array_equal(np.asarray(got).astype(np.int32), c['assign']), name


def test_kmeans_nan_center_known_answer(golden_dir):
    c = _cases(np.load(os.path.join(golden_dir, 'kmeans_ref.npz')))['nan_center']
    # k=8 with 6 rows: init clusters 5..7 are empty -> NaN centres -> numpy argmin returns
    # the first NaN column for every row
    assert np.all(c['assign'] == 5)


def test_prior_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, 'prior_ref.npz'))
    np.testing.assert_allclose(so.create_prior(g['lab_a'], 0.75, 0.5, 0.1, 0.1), g['w_a'],
                               rtol=1e-13, atol=0)
    np.testing.assert_allclose(so.create_prior(g['lab_b'], 0.6, 0.4, 0.2, 0.15), g['w_b'],
                               rtol=1e-13, atol=0)
    assert np.array_equal(so.create_prior_map(28, 28, 0.75, 0.5, 0.1, 0.1), g['cell_28'])
    assert np.array_equal(so.create_prior_map(16, 32, 0.75, 0.5, 0.1, 0.1), g['cell_16x32'])
    gy, gx = so.prior_axes(16, 32, 0.75, 0.5, 0.1, 0.1)
    np.testing.assert_allclose(np.outer(gy, gx), g['cell_16x32'], rtol=4e-15, atol=1e-300)  # separable form: few-ulp


def test_paint_and_centroids_match_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, 'weighted_kmeans_ref.npz'))
    labs, n_per = g['labs'], list(g['n_per'])
    assign = so.kmeans(int(g['k']), g['anchor_features'], g['weights'],
                       init_assign=g['init'].astype(np.float64), verbose=False)
    cmap, road = so.weighted_kmeans_paint(labs, assign, n_per)
    assert cmap.dtype == labs.dtype
    assert np.array_equal(cmap, g['cluster_map'])
    assert np.array_equal(road, g['road'])
    # centroid columns of the reference's anchor path == exact integer sums / area
    off = 0
    for i, n in enumerate(n_per):
        area, sy, sx = so.superpixel_stats(labs[i], n)
        # (1 ulp slack: the reference averages 10 identical centroid copies, :270-274)
        np.testing.assert_allclose(g['anchor_features'][off:off + n, -2], sy / area, rtol=1e-15)
        np.testing.assert_allclose(g['anchor_features'][off:off + n, -1], sx / area, rtol=1e-15)
        off += n


def test_overlap_regression_golden(golden_dir):
    cases = _cases(np.load(os.path.join(golden_dir, 'overlap_oracle.npz')))
    for name, c in cases.items():
        ip, ix, ct = so.overlap_csr(c['label'], int(c['fh']), int(c['fw']))
        assert np.array_equal(ip, c['indptr']) and np.array_equal(ix, c['indices']) \
            and np.array_equal(ct, c['counts']), name
        assert ct.sum() == c['label'].size
        for s in range(len(ip) - 1):
            assert np.all(np.diff(ix[ip[s]:ip[s + 1]]) > 0)


@pytest.mark.parametrize('H,W,fh,fw,gy,gx', [(64, 128, 8, 16, 4, 8), (50, 70, 7, 9, 3, 4),
                                             (224, 224, 28, 28, 7, 7)])
def test_pool_count_equals_dense_nearest(H, W, fh, fw, gy, gx):
    lab = synth.voronoi_labels(H, W, gy, gx, image_index=4)
    F = synth.smooth_features(5, fh, fw, seed=3, radius=1)
    ip, ix, ct = so.overlap_csr(lab, fh, fw)
    area, sy, sx = so.superpixel_stats(lab)
    assert np.array_equal(np.diff(np.concatenate([[0], np.cumsum(ct)])[ip]), area)
    got = so.pool_count(ip, ix, ct, F.reshape(5, -1).T, area, sy, sx, True)
    want = so.pool_dense_nearest(lab, F, True)
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12)


def test_cell_map_matches_cv2_nearest():
    cv = pytest.importorskip('cv2')
    for (fh, fw, H, W) in [(128, 256, 1024, 2048), (28, 28, 224, 224), (30, 60, 1024, 2048),
                           (7, 9, 50, 70)]:
        ids = np.arange(fh * fw, dtype=np.float32).reshape(fh, fw)
        up = cv.resize(ids, (W, H), interpolation=cv.INTER_NEAREST).astype(np.int64)
        mine = so.cell_of_pixel(H, fh)[:, None] * fw + so.cell_of_pixel(W, fw)[None, :]
        assert np.array_equal(up, mine)


def test_refine_csr_equals_mask_loop():
    lab = synth.voronoi_labels(64, 96, 4, 6, image_index=9)
    fh, fw = 8, 12
    rs = np.random.RandomState(0)
    road_cell = rs.rand(fh, fw) < 0.3
    ip, ix, ct = so.overlap_csr(lab, fh, fw)
    for thr in (0.01, 0.05, 0.2):
        ov, road_px, keep = so.refine_overlaps_csr(ip, ix, ct, road_cell, thr)
        want = so.refine_overlaps_masks(lab, so.upsample_nearest(road_cell, 64, 96), thr)
        assert road_px == so.upsample_nearest(road_cell, 64, 96).sum()
        assert np.array_equal(keep[lab].astype(np.uint8), want)


def test_direct_features_layout():
    F = np.arange(2 * 3 * 4 * 5, dtype=np.float32).reshape(2, 3, 4, 5)
    X = so.direct_features(F)
    assert X.shape == (40, 5) and X.dtype == np.float64
    assert np.array_equal(X[7, :3], F[0, :, 1, 2]) and tuple(X[7, 3:]) == (2.0, 1.0)  # (x, y)
    assert np.array_equal(X[20 + 19, :3], F[1, :, 3, 4]) and tuple(X[39, 3:]) == (4.0, 3.0)


def test_road_iou():
    gt = np.array([[1, 1, 0, -1], [0, 1, 0, 0]])
    pred = np.array([[1, 0, 1, 1], [0, 1, 0, 0]])
    iou, prec, rec, tp, fp, fn = so.road_iou(pred, gt)
    assert (tp, fp, fn) == (2, 1, 1) and iou == 0.5


# ---- live checks against the reference source (authoring container only) -------------


@pytest.mark.needs_reference
@pytest.mark.parametrize('script', ['batch_spalign_kmeans.py', 'direct_clustering.py',
                                    'superpixel_overlaps.py'])
def test_kmeans_live_reference_same_stream(script):
    ref = ref_extract.load(script, seed=1111)
    rs = np.random.RandomState(3)
    X = np.concatenate([rs.standard_normal((90, 7)) + 5 * rs.standard_normal((1, 7))
                        for _ in range(4)])
    w = rs.uniform(0, 1, len(X))
    for k in (2, 4, 5):
        np.random.seed(1111)
        want = ref.kmeans(k, X, w)
        np.random.seed(1111)
        got = so.kmeans(k, X, w, verbose=False)
        assert np.array_equal(np.asarray(want), np.asarray(got))


@pytest.mark.needs_reference
def test_create_prior_live_reference():
    ref = ref_extract.load('batch_spalign_kmeans.py')
    lab = synth.blob_labels(40, 60, 12, seed=1)
    np.testing.assert_allclose(so.create_prior(lab, 0.75, 0.5, 0.1, 0.1),
                               ref.create_prior(lab, 0.75, 0.5, 0.1, 0.1), rtol=1e-13)


@pytest.mark.parametrize('H,W,fh,fw,gy,gx', [(64, 128, 8, 16, 4, 8), (50, 70, 7, 9, 3, 4), (33, 47, 5, 6, 3, 3)])
def test_bilinear_overlap_equals_dense_resize_mean(H, W, fh, fw, gy, gx):
    lab = synth.voronoi_labels(H, W, gy, gx, image_index=2)
    F = synth.smooth_features(6, fh, fw, seed=5, radius=1)
    ip, ix, wv = so.overlap_bilinear_csr(lab, fh, fw)
    area, _, _ = so.superpixel_stats(lab)
    rows = np.repeat(np.arange(len(area)), np.diff(ip))
    np.testing.assert_allclose(np.bincount(rows, weights=wv), area, rtol=1e-12)   # weights sum to 1 per pixel
    for s in range(len(area)):
        assert np.all(np.diff(ix[ip[s]:ip[s + 1]]) > 0)
    got = so.pool_weighted(ip, ix, wv, F.reshape(6, -1).T, area)
    np.testing.assert_allclose(got, so.pool_dense_bilinear(lab, F), rtol=1e-11, atol=1e-12)


def test_resize_bilinear_matches_torch_align_corners():
    torch = pytest.importorskip('torch')
    F = synth.smooth_features(3, 7, 9, seed=1, radius=1).astype(np.float64)
    want = torch.nn.functional.interpolate(torch.from_numpy(F)[None], size=(50, 70), mode='bilinear',
                                           align_corners=True)[0].numpy()
    np.testing.assert_allclose(so.resize_bilinear(F, 50, 70), want, rtol=1e-12, atol=1e-12)


# ------------------------------------------------ inline reference code (refine, direct build)
def test_refine_matches_reference_inline_loop_golden(golden_dir):
    """refine_ref.npz = the reference's own loop (superpixel_overlaps.py:360-369) run unmodified:
    both oracle formulations (CSR row-dot and mask loop) must reproduce it bit for bit."""
    g = np.load(os.path.join(golden_dir, 'refine_ref.npz'))
    for name in ('s8', 'ratio', 'r224'):
        lab, road = g[name + '__label'], g[name + '__road_cell']
        fh, fw = int(g[name + '__fh']), int(g[name + '__fw'])
        H, W = lab.shape
        ip, ix, ct = so.overlap_csr(lab, fh, fw)
        for thr in (0.01, 0.05, 0.2):
            want = g['%s__refined_%g' % (name, thr)]
            _, _, keep = so.refine_overlaps_csr(ip, ix, ct, road, thr)
            assert np.array_equal(keep[lab].astype(np.uint8), want), (name, thr)
            assert np.array_equal(so.refine_overlaps_masks(lab, so.upsample_nearest(road, H, W), thr),
                                  want), (name, thr)
    lab = g['noroad__label']
    ip, ix, ct = so.overlap_csr(lab, 8, 12)
    _, px, keep = so.refine_overlaps_csr(ip, ix, ct, np.zeros((8, 12), bool), 0.01)
    assert px == 0 and not keep.any() and not g['noroad__refined'].any()


def test_direct_feature_build_matches_reference_inline_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, 'direct_features_ref.npz'))
    X = so.direct_features(g['feats'])
    assert X.dtype == g['X'].dtype and np.array_equal(X, g['X'])
    h, w, n = int(g['h']), int(g['w']), int(g['n'])
    prior = so.create_prior_map(h, w, 0.75, 0.5, 0.1, 0.1).reshape(1, -1).repeat(n, axis=0).reshape(-1)
    assert np.array_equal(prior, g['prior'])
    got = so.kmeans(4, X, prior, init_assign=g['init'].astype(np.float64), verbose=False)
    assert np.array_equal(np.asarray(got).astype(np.int32), g['assign'])


def test_gapped_labels_match_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, 'gapped_ref.npz'))
    labs, n_per = g['labs'], list(g['n_per'])
    w = np.concatenate([so.create_prior(l, 0.75, 0.5, 0.1, 0.1) for l in labs])
    np.testing.assert_allclose(w, g['weights'], rtol=1e-13)
    assign = so.kmeans(int(g['k']), g['feats'], g['weights'], init_assign=g['init'].astype(np.float64),
                       verbose=False)
    cmap, road = so.weighted_kmeans_paint(labs, assign, n_per)
    assert np.array_equal(cmap, g['cluster_map']) and np.array_equal(road, g['road'])


@pytest.mark.needs_reference
def test_inline_reference_code_live():
    """The golden files above regenerate from the live reference (statement nodes taken from
    the parsed source by line range; a drifted range raises)."""
    import types
    f = ref_extract.refine_loop()
    lab = synth.voronoi_labels(40, 56, 3, 4, image_index=5)
    road = np.random.RandomState(1).rand(5, 7) < 0.4
    (ref,) = f(road.copy(), lab, types.SimpleNamespace(overlap_threshold=0.03))
    assert np.array_equal(np.asarray(ref, np.uint8),
                          so.refine_overlaps_masks(lab, so.upsample_nearest(road, 40, 56), 0.03))
    feats = np.stack([synth.smooth_features(5, 4, 6, seed=i, radius=1) for i in range(3)])
    X, n, h, w = ref_extract.direct_feature_build()([feats], np)
    assert (n, h, w) == (3, 4, 6) and np.array_equal(X, so.direct_features(feats))
    with pytest.raises(ValueError):
        ref_extract.load_inline('superpixel_overlaps.py', 'estimate_road_mask', 361, 369, 'f',
                                ['road_mask', 'superpixel', 'args'], ['refined_roadmap'])


def test_slic_oracle_contract():
    """The SLIC restatement (parity unpinned: scikit-image absent): contiguous ids numbered in
    raster order, 4-connected segments, seed grid of skimage.util.regular_grid."""
    from scipy import ndimage
    rs = np.random.RandomState(0)
    img = ndimage.uniform_filter(rs.rand(3, 64, 96), (0, 9, 9)).astype(np.float32)
    img = (img - img.min()) / (img.max() - img.min())
    assert so.slic_grid(1024, 2048, 1000) == (22, 46, 22, 46)
    assert so.slic_grid(224, 224, 100) == (11, 22, 11, 22)
    raw = so.slic(img, 24, return_raw=True)
    assert len(np.unique(raw)) == 24
    lab = so.slic(img, 24)
    S = lab.max() + 1
    assert len(np.unique(lab)) == S and lab[0, 0] == 0
    first = np.full(S, lab.size)
    np.minimum.at(first, lab.ravel(), np.arange(lab.size))
    assert np.all(np.diff(first) > 0)
    assert all(ndimage.label(lab == v)[1] == 1 for v in range(S))
    # known answer of the Lab conversion (skimage.color.rgb2lab of pure white / mid grey)
    np.testing.assert_allclose(so.rgb2lab(np.array([1.0, 1.0, 1.0])), [100.0, -0.00245, 0.00465], atol=2e-5)
    np.testing.assert_allclose(so.rgb2lab(np.array([0.5, 0.5, 0.5]))[0], 53.389, atol=2e-3)


def test_anchor_pooling_matches_reference_golden(golden_dir):
    """anchors_ref.npz = the reference's superpixel_align (anchor path) with the anchors it drew:
    wherever no tie between equidistant cell centres is involved the restatement is exact; the
    centroid columns always are."""
    g = np.load(os.path.join(golden_dir, 'anchors_ref.npz'))
    lab, fm = g['label'], g['feature_map']
    C = fm.shape[0]
    H = lab.shape[0]
    area, sy, sx = so.superpixel_stats(lab)
    for n_select in (1, 10):
        anchors, n_valid = g['anchors_%d' % n_select], g['n_valid_%d' % n_select]
        want = g['features_%d' % n_select]
        got = so.pool_anchors(fm, H, anchors, n_valid)
        rows_ok = g['tie_free_%d' % n_select].all(axis=1) | (n_valid == 0)
        assert rows_ok.sum() >= (20 if n_select == 1 else 1)
        np.testing.assert_allclose(got[rows_ok], want[rows_ok, :C], rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(want[:, C], sy / area, rtol=1e-14)      # center_of_mass: 1 ulp
        np.testing.assert_allclose(want[:, C + 1], sx / area, rtol=1e-14)
        # with ties the restatement may pick other (equidistant) corner cells: bounded effect
        assert np.abs(got - want[:, :C]).max() < 0.5 * np.abs(want[:, :C]).max()


def test_felzenszwalb_oracle_contract():
    """The restatement of skimage 0.13's felzenszwalb: smoothing equals scipy's gaussian_filter
    to rounding, labels are contiguous, every segment reaches min_size, a constant image is one
    segment, and a larger scale merges more."""
    from scipy import ndimage
    rs = np.random.RandomState(3)
    img = ndimage.uniform_filter(rs.rand(3, 36, 52), size=(1, 5, 5)).astype(np.float32)
    hwc = img.transpose(1, 2, 0).astype(np.float64)
    assert np.abs(so.felz_blur(hwc, 0.8) - ndimage.gaussian_filter(hwc, sigma=[0.8, 0.8, 0])).max() < 1e-15
    lab = so.felzenszwalb(img, scale=4.0, sigma=0.8, min_size=9)
    n = int(lab.max()) + 1
    assert n > 3 and np.array_equal(np.unique(lab), np.arange(n))
    assert np.bincount(lab.ravel()).min() >= 9
    assert so.felzenszwalb(img, scale=40.0, sigma=0.8, min_size=9).max() + 1 < n
    assert so.felzenszwalb(np.full((3, 12, 12), 0.5, np.float32), 1.0, 0.8, 4).max() == 0
    # edges: skimage's order and endpoints (right, down, down-right, up-right)
    e = so.felz_edges(3, 4)
    assert len(e) == 3 * 3 + 2 * 4 + 2 * 2 * 3
    assert e[0].tolist() == [1, 0] and e[9].tolist() == [4, 0] and e[17].tolist() == [5, 0] \
        and e[23].tolist() == [1, 4]

