"""GPU parity at the BASELINE.json configurations: full-size images with 500/1000/2000/4000
superpixels (configs[3]), joint clustering of a batch (the reference's --batchsize, configs[1]
"joint mode"), the reference's default 224x224 / 28x28 operating point with felzenszwalb-shaped
maps, and direct cell clustering at full feature-map size (direct_clustering.py)."""
import types

import numpy as np
import pytest

torch = pytest.importorskip('torch')

from oracle import spalign_oracle as so  # noqa: E402
from superpixel_align_b200 import synth  # noqa: E402

pytestmark = pytest.mark.gpu
PRIOR = (0.75, 0.5, 0.1, 0.1)


def dev():
    return torch.device('cuda', 0)


def _oracle_image(lab, feat_cell, fh, fw, k, init):
    return so.spalign_image_cpu(lab, feat_cell, fh, fw, k=k, prior=PRIOR, append_pos=True,
                                init_assign=init)


@pytest.mark.parametrize('S,grid', [(500, (20, 25)), (2000, (40, 50)), (4000, (50, 80))])
def test_full_size_superpixel_count_sweep(S, grid):
    from superpixel_align_b200 import ops, pipeline
    H, W, fh, fw, C = 1024, 2048, 128, 256, 64
    labs = synth.voronoi_labels_torch(2, H, W, grid[0], grid[1], first_index=3, device=dev())
    g = torch.Generator(device=dev()).manual_seed(S)
    feats = torch.randn((2, fh * fw, C), generator=g, device=dev())
    feats = feats + 3 * torch.sin(torch.arange(fh * fw, device=dev())[None, :, None] / 3000.0)
    np.random.seed(S)
    out = pipeline.run_batch(labs, feats, [S, S], fh, fw, k=4, prior=PRIOR)
    nnz = out.overlap.validate()
    ip = out.overlap.indptr.cpu().numpy()
    lab0 = labs[0].cpu().numpy()
    # replay the seeded stream for image 0: first shuffle drawn is image 0's
    np.random.seed(S)
    w0 = out.weights[:S].cpu().numpy()
    init0 = so.kmeans_init(4, w0)
    ref = _oracle_image(lab0, feats[0].cpu().numpy(), fh, fw, 4, init0)
    assert np.array_equal(ip[:S + 1], ref['indptr'])
    assert np.array_equal(out.overlap.indices[:ip[S]].cpu().numpy(), ref['indices'])
    assert np.array_equal(out.overlap.counts[:ip[S]].cpu().numpy(), ref['counts'])
    np.testing.assert_allclose(w0, ref['weights'], rtol=1e-12)
    gf = out.features[:S].cpu().numpy()
    np.testing.assert_allclose(gf, ref['features'].astype(np.float32), rtol=1e-5, atol=2e-6)
    # k-means on the GPU's own descriptors (what the oracle would get handed)
    oa = so.kmeans(4, gf.astype(np.float64), w0, init_assign=init0, verbose=False)
    assert np.array_equal(out.assign[:S].cpu().numpy(), np.asarray(oa).astype(np.int32))
    cm, rm = so.weighted_kmeans_paint(lab0[None], oa, [S])
    assert np.array_equal(out.cluster_map[0].cpu().numpy(), cm[0].astype(np.uint8))
    assert np.array_equal(out.road_mask[0].cpu().numpy().astype(bool), rm[0])
    assert nnz == ip[-1]


def test_joint_clustering_of_a_batch_matches_oracle():
    # --batchsize 6: the superpixels of 6 images form ONE k-means problem (6000 rows > 4096:
    # host-side seeded init, chunked multi-CTA k-means)
    from superpixel_align_b200 import pipeline
    H, W, fh, fw, C = 256, 512, 32, 64, 48
    n, gy, gx = 6, 25, 40
    S = gy * gx
    labs = np.stack([synth.voronoi_labels(H, W, gy, gx, image_index=i) for i in range(n)])
    feats = np.stack([synth.smooth_features(C, fh, fw, seed=50 + i).reshape(C, -1).T for i in range(n)])
    np.random.seed(99)
    out = pipeline.run_batch(torch.from_numpy(labs).to(dev()), torch.from_numpy(np.ascontiguousarray(feats)).to(dev()),
                             [S] * n, fh, fw, k=4, prior=PRIOR, images_per_group=n)
    gf = out.features.cpu().numpy().astype(np.float64)
    gw = out.weights.cpu().numpy()
    np.random.seed(99)
    want, info = so.kmeans(4, gf, gw, return_info=True, verbose=False)
    assert np.array_equal(out.assign.cpu().numpy(), np.asarray(want).astype(np.int32))
    assert out.iters[0].item() == info['iters'] and out.status[0].item() == info['status']
    cm, rm = so.weighted_kmeans_paint(labs, want, [S] * n)
    assert np.array_equal(out.cluster_map.cpu().numpy(), cm.astype(np.uint8))


def test_reference_default_operating_point_224_felzenszwalb_like_batch30():
    # the shipped drivers run 224x224 inputs, 28x28 features, irregular superpixels, batchsize 30
    from superpixel_align_b200 import batch_spalign_kmeans as bsk
    rs = np.random.RandomState(0)
    n = 30
    labs = np.stack([synth.blob_labels(224, 224, int(rs.randint(25, 70)), seed=i, dtype=np.int64) for i in range(n)])
    feats = np.stack([synth.smooth_features(64, 28, 28, seed=100 + i) for i in range(n)])
    args = types.SimpleNamespace(gpu=0, n_clusters=4, without_pos=False, y_rel_pos=0.75, x_rel_pos=0.5,
                                 y_rel_sigma=0.1, x_rel_sigma=0.1)
    bsk.clear_cache()
    f, n_per = bsk.batch_superpixel_align(args, None, np.zeros((n, 3, 224, 224), np.float32), labs, feats)
    w = bsk.batch_create_prior(args, labs)
    assert n_per == [int(l.max()) + 1 for l in labs]
    np.random.seed(1111)
    cres, road = bsk.batch_weighted_kmeans(args, labs, f, w, n_per)
    of, ow = [], []
    for i in range(n):
        r = so.spalign_image_cpu(labs[i], feats[i].reshape(64, -1).T, 28, 28, init_assign=np.zeros(n_per[i]))
        of.append(r['features']); ow.append(r['weights'])
    np.testing.assert_allclose(f, np.concatenate(of), rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(w, np.concatenate(ow), rtol=1e-12)
    np.random.seed(1111)
    oa = so.kmeans(4, f, w, verbose=False)
    ocm, orm = so.weighted_kmeans_paint(labs, oa, n_per)
    assert np.array_equal(cres, ocm) and np.array_equal(road, orm)


def test_direct_clustering_full_feature_map():
    # direct_clustering.py on one full-size stride-8 map: 32768 rows x (C + 2 virtual columns)
    from superpixel_align_b200 import direct_clustering as dc
    C, h, w = 32, 128, 256
    feats = synth.smooth_features(C, h, w, seed=7)[None]
    args = types.SimpleNamespace(gpu=0, n_clusters=4, y_rel_pos=0.75, x_rel_pos=0.5, y_rel_sigma=0.1, x_rel_sigma=0.1)
    np.random.seed(5)
    cres, road = dc.estimate_road_mask(feats, args)
    X = so.direct_features(feats)
    prior = so.create_prior_map(h, w, *PRIOR).reshape(-1)
    np.random.seed(5)
    want = so.kmeans(4, X, prior, verbose=False)
    assert np.array_equal(cres.reshape(-1), np.asarray(want).astype(np.int32))
    assert road.sum() > 0
