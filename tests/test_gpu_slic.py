"""f3: SLIC superpixels on the device (csrc/slic.cu) against the NumPy restatement of the same
contract (oracle/spalign_oracle.py:slic -- scikit-image is not in the reference tree, so parity
with the reference's skimage call is unpinned), plus the invariants K1 relies on."""
import types

import numpy as np
import pytest

torch = pytest.importorskip('torch')

from oracle import spalign_oracle as so  # noqa: E402

pytestmark = pytest.mark.gpu


def dev():
    return torch.device('cuda', 0)


def _image(H, W, seed, blobs=12):
    """Piecewise-smooth RGB image in 0..1: a few colour regions + low-pass noise."""
    from scipy import ndimage
    rs = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    idx = np.zeros((H, W), dtype=np.int64)
    best = np.full((H, W), np.inf)
    for b in range(blobs):
        cy, cx = rs.uniform(0, H), rs.uniform(0, W)
        d = (yy - cy) ** 2 + (xx - cx) ** 2
        idx = np.where(d < best, b, idx)
        best = np.minimum(best, d)
    cols = rs.uniform(0.1, 0.9, (blobs, 3))
    img = cols[idx].transpose(2, 0, 1) + 0.08 * ndimage.uniform_filter(rs.standard_normal((3, H, W)), (0, 5, 5))
    return np.clip(img, 0, 1).astype(np.float32)


@pytest.mark.parametrize('H,W,n_seg', [(96, 160, 60), (64, 64, 16), (120, 90, 40), (50, 200, 30)])
def test_slic_matches_oracle_bit_exact(H, W, n_seg):
    from superpixel_align_b200 import ops
    imgs = np.stack([_image(H, W, s) for s in (1, 2)])
    labels, n_labels = ops.slic(torch.from_numpy(imgs).to(dev()), n_seg)
    raw, _ = ops.slic(torch.from_numpy(imgs).to(dev()), n_seg, enforce_connectivity=False)
    for i in range(2):
        want_raw = so.slic(imgs[i], n_seg, enforce_connectivity=False)
        assert np.array_equal(raw[i].cpu().numpy(), want_raw), 'k-means phase differs'
        want = so.slic(imgs[i], n_seg)
        got = labels[i].cpu().numpy()
        assert n_labels[i].item() == want.max() + 1
        assert np.array_equal(got, want)


def test_slic_invariants_full_size_and_k1_compatibility():
    from scipy import ndimage
    from superpixel_align_b200 import ops
    H, W = 1024, 2048
    imgs = torch.from_numpy(np.stack([_image(H, W, 7, blobs=40)])).to(dev())
    labels, n_labels = ops.slic(imgs, 1000)
    lab = labels[0].cpu().numpy()
    S = int(n_labels[0].item())
    assert ops._lib.load().spalign_slic_segments(H, W, 1000) == 22 * 45
    assert 600 <= S <= 1100
    assert lab.min() == 0 and lab.max() == S - 1 and len(np.unique(lab)) == S      # contiguous ids
    first = np.full(S, H * W, dtype=np.int64)
    np.minimum.at(first, lab.ravel(), np.arange(H * W))
    assert np.all(np.diff(first) > 0)                                             # raster numbering
    sizes = np.bincount(lab.ravel())
    min_size = int(0.5 * H * W / 1000)
    assert (sizes >= min_size).sum() >= S - 1              # only an unmergeable corner piece may be small
    # 4-connected segments
    for v in np.random.RandomState(0).choice(S, 60, replace=False):
        assert ndimage.label(lab == v)[1] == 1
    # determinism and K1 on the result
    labels2, _ = ops.slic(imgs, 1000)
    assert torch.equal(labels, labels2)
    ov = ops.overlap_csr(labels, 128, 256, [S], prior=(0.75, 0.5, 0.1, 0.1))
    nnz = ov.validate()
    assert not ov.has_empty_rows and ov.area.sum().item() == H * W and nnz > S


def test_batch_superpixel_dropin_feeds_the_pipeline():
    from superpixel_align_b200 import batch_spalign_kmeans as bsk
    H, W = 128, 256
    imgs = np.stack([_image(H, W, s) for s in (3, 4)]) * 255.0       # 0..255 as the reference holds them
    args = types.SimpleNamespace(gpu=0, superpixel_method='slic', n_slic_segments=50, n_clusters=3,
                                 without_pos=False, y_rel_pos=0.75, x_rel_pos=0.5, y_rel_sigma=0.1,
                                 x_rel_sigma=0.1)
    sp = bsk.batch_superpixel(args, imgs)
    assert sp.shape == (2, H, W) and sp.dtype == np.int64
    for i in range(2):
        assert np.array_equal(sp[i], so.slic(imgs[i] / np.float32(255.0), 50))
    feats = np.random.RandomState(0).standard_normal((2, 8, H // 8, W // 8)).astype(np.float32)
    f, n_per = bsk.batch_superpixel_align(args, None, imgs, sp, feats)
    w = bsk.batch_create_prior(args, sp)
    np.random.seed(1)
    cres, road = bsk.batch_weighted_kmeans(args, sp, f, w, n_per)
    assert cres.shape == sp.shape and road.dtype == bool
    with pytest.raises(ValueError):
        bsk.batch_superpixel(types.SimpleNamespace(superpixel_method='watershed'), imgs)


# ------------------------------------------------------------------ felzenszwalb (f3, default)
@pytest.mark.parametrize('H,W,scale,sigma,min_size', [(40, 56, 8.0, 0.8, 20), (33, 47, 3.0, 0.8, 5),
                                                      (24, 64, 20.0, 1.5, 10), (32, 32, 5.0, 0.0, 4),
                                                      (16, 16, 300.0, 0.8, 20)])
def test_felzenszwalb_labels_equal_the_oracle(H, W, scale, sigma, min_size):
    """Bit-identical label maps against the NumPy restatement of skimage 0.13's algorithm
    (float64 in the same operation order; equal costs in edge order), several images side by
    side; ids are contiguous and every id is present."""
    from superpixel_align_b200 import ops
    imgs = np.stack([_image(H, W, s) for s in (1, 2, 3)])
    imgs[2, :, : H // 2] = imgs[2, :, :1, :1]            # a flat region: many zero-cost ties
    lab, n_lab = ops.felzenszwalb(torch.from_numpy(imgs).to(dev()), scale, sigma, min_size)
    lab, n_lab = lab.cpu().numpy(), n_lab.cpu().numpy()
    for i in range(3):
        ref = so.felzenszwalb(imgs[i], scale, sigma, min_size)
        assert np.array_equal(lab[i], ref), (i, int(ref.max()) + 1, int(n_lab[i]))
        assert n_lab[i] == ref.max() + 1 and np.array_equal(np.unique(lab[i]), np.arange(n_lab[i]))


def test_felzenszwalb_larger_than_shared_memory_and_dropin_default():
    """An image whose union-find does not fit shared memory takes the global-memory path (same
    labels as the in-shared-memory path on the part both can run); the drop-in's default method
    is the reference's (felzenszwalb) and its label maps feed K1."""
    from superpixel_align_b200 import batch_spalign_kmeans as bsk, ops
    H, W = 64, 96
    imgs = np.stack([_image(H, W, s) for s in (5, 6)]) * 255.0
    args = types.SimpleNamespace(gpu=0, felzenszwalb_scale=6.0, felzenszwalb_sigma=0.8,
                                 felzenszwalb_min_size=12)
    sp = bsk.batch_superpixel(args, imgs)                 # no superpixel_method: the default
    assert sp.shape == (2, H, W) and sp.dtype == np.int64
    for i in range(2):
        assert np.array_equal(sp[i], so.felzenszwalb(imgs[i] / np.float32(255.0), 6.0, 0.8, 12))
    n_sp = [int(s.max()) + 1 for s in sp]
    ov = ops.overlap_csr(torch.from_numpy(sp).to(dev()), H // 8, W // 8, n_sp)
    assert ov.validate() > 0 and not ov.has_empty_rows
    big = torch.from_numpy(np.stack([_image(240, 320, 9)])).to(dev())   # 76 800 px > 51 200
    lab, n_lab = ops.felzenszwalb(big, 40.0, 0.8, 20)
    lab = lab.cpu().numpy()[0]
    assert int(n_lab[0]) == lab.max() + 1 and np.array_equal(np.unique(lab), np.arange(lab.max() + 1))
    assert np.bincount(lab.ravel()).min() >= 20           # the min_size pass
