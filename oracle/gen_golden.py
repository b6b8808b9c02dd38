"""Freeze golden vectors from the UNMODIFIED reference functions.  TEST INFRASTRUCTURE.

Run in the authoring container (needs /root/reference):

    python oracle/gen_golden.py

Writes tests/golden/*.npz.  The reference functions are executed through
``oracle.ref_extract`` (AST extraction, NumPy shim); inputs are small seeded synthetic
arrays so the fixtures stay a few hundred KB.  The GPU box has no /root/reference: tests
there only read these files.
"""
from __future__ import annotations

import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_extract  # noqa: E402
from superpixel_align_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')


def _blobs(rs, n, d, k, spread=6.0):
    cent = rs.standard_normal((k, d)) * spread
    lab = rs.randint(0, k, n)
    return cent[lab] + rs.standard_normal((n, d))


def gen_kmeans():
    """Reference kmeans() (batch_spalign_kmeans.py:136-183) on seeded inputs."""
    ref = ref_extract.load('batch_spalign_kmeans.py', seed=1111)
    cases = {}
    rs = np.random.RandomState(7)
    specs = [  # name, N, D, K
        ('k4', 240, 10, 4), ('k2', 150, 6, 2), ('k8', 400, 12, 8), ('k3_pos', 300, 18, 3),
    ]
    for name, n, d, k in specs:
        X = _blobs(rs, n, d, k)
        if name.endswith('pos'):  # last two columns in pixel units like the spalign path
            X[:, -2] = rs.uniform(0, 1023, n)
            X[:, -1] = rs.uniform(0, 2047, n)
        X = X.astype(np.float32).astype(np.float64)  # fp32-representable, as the CUDA path stores it
        w = rs.uniform(0.0, 1.0, n)
        state = np.random.get_state()
        assign = ref.kmeans(k, X, w)
        after = np.random.get_state()
        # recover the init the reference drew, by replaying the stream
        np.random.set_state(state)
        from oracle import spalign_oracle as so
        init = so.kmeans_init(k, w)
        np.random.set_state(after)
        cases[name] = dict(X=X.astype(np.float32), w=w, k=k, init=init.astype(np.int32),
                           assign=np.asarray(assign).astype(np.int32))
    # NaN centre: more clusters than low-prior rows -> empty init clusters
    X = _blobs(rs, 6, 4, 2).astype(np.float32).astype(np.float64)
    w = np.array([0.9, 0.1, 0.8, 0.2, 0.7, 0.3])
    state = np.random.get_state()
    with np.errstate(all='ignore'):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            assign = ref.kmeans(8, X, w)
    np.random.set_state(state)
    from oracle import spalign_oracle as so
    init = so.kmeans_init(8, w)
    cases['nan_center'] = dict(X=X.astype(np.float32), w=w, k=8, init=init.astype(np.int32),
                               assign=np.asarray(assign).astype(np.int32))
    flat = {}
    for name, c in cases.items():
        for key, v in c.items():
            flat['%s__%s' % (name, key)] = v
    np.savez_compressed(os.path.join(OUT, 'kmeans_ref.npz'), **flat)
    print('kmeans_ref.npz:', {n: int(c['assign'].max()) for n, c in cases.items()})


def gen_prior():
    """Reference create_prior() superpixel form (batch_spalign_kmeans.py:111-129) and cell
    form (direct_clustering.py:188-201)."""
    ref = ref_extract.load('batch_spalign_kmeans.py')
    refd = ref_extract.load('direct_clustering.py')
    lab = synth.voronoi_labels(64, 128, 4, 8, image_index=0)
    lab2 = synth.voronoi_labels(56, 56, 5, 5, image_index=3)
    out = dict(
        lab_a=lab, w_a=ref.create_prior(lab, 0.75, 0.5, 0.1, 0.1),
        lab_b=lab2, w_b=ref.create_prior(lab2, 0.6, 0.4, 0.2, 0.15),
        cell_28=refd.create_prior(28, 28, 0.75, 0.5, 0.1, 0.1),
        cell_16x32=refd.create_prior(16, 32, 0.75, 0.5, 0.1, 0.1),
    )
    np.savez_compressed(os.path.join(OUT, 'prior_ref.npz'), **out)
    print('prior_ref.npz:', out['w_a'].shape, out['w_b'].shape)


def gen_weighted_kmeans():
    """Reference weighted_kmeans() (k-means + paint-back, :186-207) on a 2-image batch, and
    the reference anchor-sampled superpixel_align (:210-276) for the centroid columns."""
    ref = ref_extract.load('batch_spalign_kmeans.py', seed=1111)
    H, W, fh, fw, C = 32, 64, 4, 8, 6
    labs = np.stack([synth.voronoi_labels(H, W, 3, 5, image_index=i, dtype=np.int64)
                     for i in range(2)])
    feats = [synth.smooth_features(C, fh, fw, seed=10 + i, radius=1) for i in range(2)]
    imgs = np.zeros((2, 3, H, W), dtype=np.float32)
    sp_feats, n_per = [], []
    for i in range(2):
        f = ref.superpixel_align(imgs[i], feats[i], labs[i], 10, 4, True)
        sp_feats.append(f)
        n_per.append(len(np.unique(labs[i])))
    sp_feats = np.concatenate(sp_feats)
    w = np.concatenate([ref.create_prior(l, 0.75, 0.5, 0.1, 0.1) for l in labs])
    state = np.random.get_state()
    cres, road = ref.weighted_kmeans(labs, sp_feats, w, 3, n_per)
    np.random.set_state(state)
    from oracle import spalign_oracle as so
    init = so.kmeans_init(3, w)
    np.savez_compressed(
        os.path.join(OUT, 'weighted_kmeans_ref.npz'), labs=labs, feats=np.stack(feats),
        anchor_features=sp_feats, weights=w, n_per=np.array(n_per), k=3,
        init=init.astype(np.int32), cluster_map=cres, road=road)
    print('weighted_kmeans_ref.npz:', sp_feats.shape, cres.shape, int(road.sum()))


def gen_overlap():
    """No reference source exists for the count matrix (parity unpinned).  Freeze the
    oracle's CSR for three small maps as a regression fixture, after cross-checking it
    against the reference-style mask loop."""
    from oracle import spalign_oracle as so
    out = {}
    for name, lab, fh, fw in [
        ('vor', synth.voronoi_labels(64, 128, 4, 8, image_index=1), 8, 16),
        ('ragged', synth.voronoi_labels(50, 70, 3, 4, image_index=2), 7, 9),
        ('noise', synth.noise_labels(24, 40, 17, seed=5), 3, 5),
    ]:
        ip, ix, ct = so.overlap_csr(lab, fh, fw)
        H, W = lab.shape
        cell = so.cell_of_pixel(H, fh)[:, None] * fw + so.cell_of_pixel(W, fw)[None, :]
        for s in range(len(ip) - 1):  # mask-loop cross-check
            m = lab == s
            cols, cnt = np.unique(cell[m], return_counts=True)
            assert np.array_equal(cols, ix[ip[s]:ip[s + 1]])
            assert np.array_equal(cnt, ct[ip[s]:ip[s + 1]])
        out.update({name + '__label': lab, name + '__fh': fh, name + '__fw': fw,
                    name + '__indptr': ip, name + '__indices': ix, name + '__counts': ct})
    np.savez_compressed(os.path.join(OUT, 'overlap_oracle.npz'), **out)
    print('overlap_oracle.npz ok')


def gen_refine():
    """The reference's INLINE overlap-refine loop (superpixel_overlaps.py:360-369), run
    unmodified through ref_extract.load_inline: cell-level road masks (nearest-upsampled by the
    reference's own cv.resize call, :362), stride-8 and non-integer-ratio geometries, an image
    without predicted road, thresholds around the decision.  Anchors the overlap counts
    (M @ road_mask) to reference code."""
    import types
    ref_refine = ref_extract.refine_loop()
    out = {}
    rs = np.random.RandomState(11)
    cases = [('s8', 64, 96, 8, 12, 4, 6), ('ratio', 50, 70, 7, 9, 3, 4), ('r224', 224, 224, 28, 28, 7, 7)]
    for name, H, W, fh, fw, gy, gx in cases:
        lab = synth.voronoi_labels(H, W, gy, gx, image_index=21)
        road = rs.rand(fh, fw) < 0.35
        out[name + '__label'] = lab
        out[name + '__road_cell'] = road
        out[name + '__fh'], out[name + '__fw'] = fh, fw
        for thr in (0.01, 0.05, 0.2):
            (refined,) = ref_refine(road.copy(), lab, types.SimpleNamespace(overlap_threshold=thr))
            out['%s__refined_%g' % (name, thr)] = np.asarray(refined, dtype=np.uint8)
    lab = synth.voronoi_labels(64, 96, 4, 6, image_index=22)
    (refined,) = ref_refine(np.zeros((8, 12), dtype=bool), lab, types.SimpleNamespace(overlap_threshold=0.01))
    out['noroad__label'], out['noroad__refined'] = lab, np.asarray(refined, dtype=np.uint8)
    np.savez_compressed(os.path.join(OUT, 'refine_ref.npz'), **out)
    print('refine_ref.npz:', sorted(k for k in out if 'refined' in k))


def gen_direct():
    """The reference's INLINE direct feature build (direct_clustering.py:298-303: concat,
    (x, y) cell-index columns, transpose to rows) and prior tiling (:307-308), then its
    kmeans() on the result -- the whole direct_clustering path on a 2-image batch."""
    import types
    build = ref_extract.direct_feature_build()
    tile = ref_extract.direct_prior_tiling()
    refd = ref_extract.load('direct_clustering.py', seed=1111)
    feats = np.stack([synth.smooth_features(12, 9, 14, seed=40 + i, radius=1) for i in range(2)])
    X, n, h, w = build([feats], np)
    args = types.SimpleNamespace(y_rel_pos=0.75, x_rel_pos=0.5, y_rel_sigma=0.1, x_rel_sigma=0.1)
    (prior,) = tile(h, w, n, args)
    state = np.random.get_state()
    assign = refd.kmeans(4, X, prior)
    np.random.set_state(state)
    from oracle import spalign_oracle as so
    init = so.kmeans_init(4, prior)
    np.savez_compressed(os.path.join(OUT, 'direct_features_ref.npz'), feats=feats, X=X, prior=prior,
                        n=n, h=h, w=w, init=init.astype(np.int32),
                        assign=np.asarray(assign).astype(np.int32))
    print('direct_features_ref.npz:', X.shape, X.dtype, int(np.asarray(assign).max()))


def gen_gapped():
    """Label ids with gaps (1-based ids as skimage >= 0.19 slic yields, plus a missing id):
    the reference's create_prior / weighted_kmeans on them.  Rows follow np.unique order
    (:124, :226); the paint-back addresses pixels by the enumerate index (:195-198), so pixels
    whose label value is >= n_i keep 0."""
    ref = ref_extract.load('batch_spalign_kmeans.py', seed=1111)
    H, W = 32, 64
    labs = np.stack([synth.voronoi_labels(H, W, 3, 5, image_index=30 + i, dtype=np.int64)
                     for i in range(2)]) + 1
    labs[1][labs[1] >= 7] += 2          # ids 7, 8 missing in image 1
    n_per = [len(np.unique(l)) for l in labs]
    w = np.concatenate([ref.create_prior(l, 0.75, 0.5, 0.1, 0.1) for l in labs])
    rs = np.random.RandomState(3)
    feats = (rs.standard_normal((3, 6)) * 3)[rs.randint(0, 3, sum(n_per))] + \
        rs.standard_normal((sum(n_per), 6))
    feats = feats.astype(np.float32).astype(np.float64)
    state = np.random.get_state()
    cres, road = ref.weighted_kmeans(labs, feats, w, 3, n_per)
    np.random.set_state(state)
    from oracle import spalign_oracle as so
    init = so.kmeans_init(3, w)
    np.savez_compressed(os.path.join(OUT, 'gapped_ref.npz'), labs=labs, n_per=np.array(n_per),
                        weights=w, feats=feats, k=3, init=init.astype(np.int32), cluster_map=cres,
                        road=road)
    print('gapped_ref.npz:', n_per, cres.shape)


def gen_anchors():
    """The reference's anchor-sampled superpixel_align (:210-276), run unmodified with
    n_select = 1 and 10; the anchors it drew are recovered by replaying Python's ``random`` from
    the same state.  ``tie_free`` marks anchors whose 4th and 5th nearest cell centres are not
    equidistant: there the reference's unstable argsort cannot matter and the result is pinned."""
    ref = ref_extract.load('batch_spalign_kmeans.py', seed=1111)
    from oracle import spalign_oracle as so
    H, W, fh, fw, C = 64, 128, 8, 16, 8
    lab = synth.voronoi_labels(H, W, 5, 8, image_index=12, dtype=np.int64)
    fm = synth.smooth_features(C, fh, fw, seed=9, radius=1)
    img = np.zeros((3, H, W), dtype=np.float32)
    yy, xx = np.meshgrid(np.arange(fh), np.arange(fw))
    flat = (np.stack([yy, xx]).transpose(1, 2, 0) + 0.5).reshape(-1, 2)
    out = dict(label=lab, feature_map=fm)
    for n_select in (1, 10):
        state = random.getstate()
        f = ref.superpixel_align(img, fm, lab, n_select, 4, True)
        random.setstate(state)
        anchors, n_valid = so.replay_reference_anchors(lab, n_select)
        tie_free = np.zeros(anchors.shape[:2], dtype=bool)
        for s_ in range(anchors.shape[0]):
            for a in range(int(n_valid[s_])):
                py, px = so.anchor_to_feature_coords(int(anchors[s_, a, 0]), int(anchors[s_, a, 1]), H, fh, fw)
                d2 = np.sort(((flat - np.array([py, px])) ** 2).sum(axis=1))
                tie_free[s_, a] = d2[3] < d2[4]
        out.update({'anchors_%d' % n_select: anchors.astype(np.int32),
                    'n_valid_%d' % n_select: n_valid.astype(np.int32),
                    'features_%d' % n_select: np.asarray(f, dtype=np.float64),
                    'tie_free_%d' % n_select: tie_free})
    np.savez_compressed(os.path.join(OUT, 'anchors_ref.npz'), **out)
    print('anchors_ref.npz:', out['features_10'].shape, 'tie-free anchors',
          float(out['tie_free_10'].mean()))


if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    random.seed(1111)
    gen_kmeans()
    gen_prior()
    gen_weighted_kmeans()
    gen_overlap()
    gen_refine()
    gen_direct()
    gen_gapped()
    gen_anchors()
