"""GPU parity tests: the CUDA path (through the C ABI, via superpixel_align_b200.ops) against
the CPU oracle on the same seeded inputs, the golden fixtures frozen from the reference, and
size-independent properties at the BASELINE.json sizes (1024x2048).

Bars (BASELINE.json north_star): overlap CSR / statistics / paint / refine bit-exact;
pooled features rtol 1e-5 (fp32 vs float64 oracle); prior weights rtol 1e-12; k-means
assignments identical under the same init.
"""
import os

import numpy as np
import pytest

torch = pytest.importorskip('torch')

from oracle import spalign_oracle as so  # noqa: E402
from superpixel_align_b200 import synth  # noqa: E402

pytestmark = pytest.mark.gpu

PRIOR = (0.75, 0.5, 0.1, 0.1)


@pytest.fixture(scope='module')
def ops():
    from superpixel_align_b200 import _lib, ops as _ops
    _lib.load()
    return _ops


def dev():
    return torch.device('cuda', 0)


def _check_overlap(ops, labs, fh, fw, n_sp=None, prior=PRIOR, cap=None):
    """labs: list of [H,W] arrays (same shape/dtype).  Returns the Overlap."""
    labs = [np.asarray(l) for l in labs]
    if n_sp is None:
        n_sp = [int(l.max()) + 1 for l in labs]
    t = torch.from_numpy(np.stack(labs)).to(dev())
    ov = ops.overlap_csr(t, fh, fw, n_sp, prior=prior, nnz_cap_per_image=cap, retry=True)
    nnz = ov.validate()
    ip = ov.indptr.cpu().numpy()
    ix = ov.indices.cpu().numpy()
    ct = ov.counts.cpu().numpy()
    assert ip[0] == 0 and ip[-1] == nnz
    off = 0
    for i, lab in enumerate(labs):
        oip, oix, oct_ = so.overlap_csr(lab, fh, fw, n_sp[i])
        a, b = ip[off], ip[off + n_sp[i]]
        assert np.array_equal(ip[off:off + n_sp[i] + 1] - a, oip)
        assert np.array_equal(ix[a:b], oix)
        assert np.array_equal(ct[a:b], oct_)
        area, sy, sx = so.superpixel_stats(lab, n_sp[i])
        assert np.array_equal(ov.area[off:off + n_sp[i]].cpu().numpy(), area)
        assert np.array_equal(ov.sum_y[off:off + n_sp[i]].cpu().numpy(), sy)
        assert np.array_equal(ov.sum_x[off:off + n_sp[i]].cpu().numpy(), sx)
        if prior is not None and area.min() > 0:
            w = so.create_prior(lab, *prior)
            np.testing.assert_allclose(ov.weights()[off:off + n_sp[i]].cpu().numpy(), w,
                                       rtol=1e-12, atol=1e-300)
        off += n_sp[i]
    return ov


# ---------------------------------------------------------------------------- K1 overlap
@pytest.mark.parametrize('dtype', [np.int32, np.int64])
def test_overlap_stride8_voronoi(ops, dtype):
    labs = [synth.voronoi_labels(64, 128, 4, 8, image_index=i, dtype=dtype) for i in range(3)]
    _check_overlap(ops, labs, 8, 16)


@pytest.mark.parametrize('dtype', [torch.int32, torch.int64])
def test_overlap_stride8_label_pointer_alignment(ops, dtype):
    """emit v3 loads 32 bytes of labels at a time where the map is 32-byte aligned and 2 x 16
    bytes otherwise: a label map that starts 16 bytes into an allocation gives the same matrix."""
    labs = np.stack([synth.voronoi_labels(64, 128, 4, 8, image_index=20 + i) for i in range(3)])
    n_sp = [int(l.max()) + 1 for l in labs]
    aligned = torch.from_numpy(labs).to(dev()).to(dtype)
    shift = 16 // aligned.element_size()
    buf = torch.zeros(aligned.numel() + shift, dtype=dtype, device=dev())
    shifted = buf[shift:].view(aligned.shape)
    shifted.copy_(aligned)
    assert aligned.data_ptr() % 32 == 0 and shifted.data_ptr() % 32 == 16
    a = ops.overlap_csr(aligned, 8, 16, n_sp, prior=PRIOR)
    b = ops.overlap_csr(shifted, 8, 16, n_sp, prior=PRIOR)
    nnz = a.validate()
    assert nnz == b.validate()
    for name in ('indptr', 'area', 'sum_y', 'sum_x', 'sum_prior'):
        assert torch.equal(getattr(a, name), getattr(b, name)), name
    for name in ('indices', 'counts'):
        assert torch.equal(getattr(a, name)[:nnz], getattr(b, name)[:nnz]), name
    _check_overlap(ops, list(labs), 8, 16)


@pytest.mark.parametrize('seed', range(12))
def test_overlap_stride8_randomised_maps(ops, seed):
    """emit v3 on random stride-8 geometries and label patterns (feature widths that are not a
    multiple of the 32-cell tile, one-cell-high maps, noise with up to 64 labels per cell, stripes
    one pixel wide, blocks aligned and misaligned with the cells, both label dtypes): bit-exact
    CSR, areas, centroid sums; prior to 1e-12."""
    rs = np.random.RandomState(1000 + seed)
    fh, fw = int(rs.randint(1, 20)), int(rs.randint(1, 70))
    H, W = 8 * fh, 8 * fw
    kind = seed % 4
    labs = []
    for i in range(int(rs.randint(1, 4))):
        if kind == 0:                                   # noise: many labels per cell
            S = int(rs.randint(2, 200))
            lab = rs.randint(0, S, size=(H, W))
        elif kind == 1:                                 # stripes of random width
            wd = int(rs.randint(1, 12))
            lab = (np.arange(W)[None, :] // wd + np.arange(H)[:, None] // int(rs.randint(1, 12))
                   * ((W + wd - 1) // wd))
        elif kind == 2:                                 # blocks shifted against the cell grid
            b, sy, sx = int(rs.randint(3, 30)), int(rs.randint(0, 8)), int(rs.randint(0, 8))
            lab = ((np.arange(H)[:, None] + sy) // b) * ((W + 8 + b - 1) // b) + \
                (np.arange(W)[None, :] + sx) // b
        else:                                           # a few big blobs + sprinkled pixels
            lab = (np.arange(H)[:, None] * 3 // max(H, 1)) * 3 + np.arange(W)[None, :] * 3 // max(W, 1)
            m = rs.rand(H, W) < 0.02
            lab = np.where(m, rs.randint(9, 40, size=(H, W)), lab)
        u, inv = np.unique(lab, return_inverse=True)
        labs.append(inv.reshape(H, W).astype(np.int64 if seed % 2 else np.int32))
    _check_overlap(ops, labs, fh, fw)


@pytest.mark.parametrize('H,W,fh,fw,gy,gx', [(50, 70, 7, 9, 3, 4), (224, 224, 28, 28, 7, 7),
                                             (33, 47, 33, 47, 3, 3), (40, 40, 1, 1, 2, 2),
                                             (64, 128, 16, 32, 4, 8), (30, 50, 40, 60, 2, 3)])
def test_overlap_generic_geometries(ops, H, W, fh, fw, gy, gx):
    labs = [synth.voronoi_labels(H, W, gy, gx, image_index=7 + i) for i in range(2)]
    _check_overlap(ops, labs, fh, fw)


def test_overlap_golden_fixture(ops, golden_dir):
    g = np.load(os.path.join(golden_dir, 'overlap_oracle.npz'))
    for name in ('vor', 'ragged', 'noise'):
        lab = g[name + '__label']
        fh, fw = int(g[name + '__fh']), int(g[name + '__fw'])
        ov = _check_overlap(ops, [lab], fh, fw, prior=None)
        nnz = ov.validate()
        assert np.array_equal(ov.indptr.cpu().numpy(), g[name + '__indptr'])
        assert np.array_equal(ov.indices.cpu().numpy()[:nnz], g[name + '__indices'])
        assert np.array_equal(ov.counts.cpu().numpy()[:nnz], g[name + '__counts'])


def test_overlap_single_superpixel_heavy_row(ops):
    # S = 1: one row holding every cell (> 256 cells -> dense heavy-row path)
    lab = np.zeros((128, 256), dtype=np.int32)
    _check_overlap(ops, [lab, lab], 16, 32)
    _check_overlap(ops, [np.zeros((90, 100), dtype=np.int64)], 30, 25)  # generic + heavy


def test_overlap_blob_labels_mixed_row_lengths(ops):
    labs = [synth.blob_labels(256, 512, 60, seed=s) for s in (1, 2)]
    n_sp = [int(l.max()) + 1 for l in labs]
    ov = _check_overlap(ops, labs, 32, 64, n_sp=n_sp)
    lens = np.diff(ov.indptr.cpu().numpy())
    assert lens.max() > 256 and lens.min() < 64  # both sort tiers exercised


def test_overlap_noise_labels_stage_overflow_and_retry(ops):
    # every 8x8 cell holds ~40 distinct labels: > 1024 pairs per block (staging overflow) and
    # more pairs than the default capacity (retry path)
    lab = synth.noise_labels(64, 128, 50, seed=3)
    _check_overlap(ops, [lab], 8, 16)
    lab2 = synth.noise_labels(48, 56, 23, seed=4)
    _check_overlap(ops, [lab2, lab2[::-1].copy()], 5, 9)  # generic kernel, many labels per cell


def test_overlap_unequal_superpixel_counts_and_gaps(ops):
    a = synth.voronoi_labels(64, 64, 4, 4, image_index=0)
    b = synth.voronoi_labels(64, 64, 2, 3, image_index=1)
    ov = _check_overlap(ops, [a, b], 8, 8)
    assert not ov.has_empty_rows
    # declare more superpixels than present: rows exist but are empty, flagged
    ov = _check_overlap(ops, [a, b], 8, 8, n_sp=[20, 9], prior=None)
    assert ov.has_empty_rows
    assert ov.area[16:20].sum().item() == 0


def test_overlap_label_out_of_range_is_flagged(ops):
    lab = synth.voronoi_labels(64, 64, 4, 4, image_index=0).copy()
    t = torch.from_numpy(lab[None]).to(dev())
    ov = ops.overlap_csr(t, 8, 8, [10])  # real maximum is 15
    with pytest.raises(ValueError):
        ov.validate()
    lab[3, 3] = -1
    ov = ops.overlap_csr(torch.from_numpy(lab[None]).to(dev()), 8, 8, [16])
    with pytest.raises(ValueError):
        ov.validate()


def test_overlap_capacity_overflow_is_flagged(ops):
    lab = synth.noise_labels(64, 128, 50, seed=3)
    ov = ops.overlap_csr(torch.from_numpy(lab[None]).to(dev()), 8, 16, [50], nnz_cap_per_image=500)
    with pytest.raises(OverflowError):
        ov.validate()


def test_pipeline_overflow_leaves_empty_matrix_and_is_reported(ops):
    # the sync-free pipeline never reads the flags: on overflow K1 must leave an EMPTY matrix so
    # that pooling / k-means / paint stay inside their buffers, and check() reports it
    from superpixel_align_b200 import pipeline
    for H, W, fh, fw in ((64, 128, 8, 16), (50, 70, 7, 9)):          # stride-8 and generic K1
        lab = np.stack([synth.noise_labels(H, W, 50, seed=s) for s in (3, 4)])
        feats = torch.randn((2, fh * fw, 16), device=dev())
        out = pipeline.run_batch(torch.from_numpy(lab).to(dev()), feats, [50, 50], fh, fw,
                                 nnz_cap_per_image=100)
        torch.cuda.synchronize()
        assert int(out.overlap.indptr.abs().sum().item()) == 0
        assert int(out.overlap.area.abs().sum().item()) == 0
        with pytest.raises(OverflowError):
            out.check()
    lab = np.stack([synth.voronoi_labels(64, 128, 4, 8, image_index=i) for i in range(2)])
    feats = torch.randn((2, 8 * 16, 16), device=dev())
    np.random.seed(3)
    out = pipeline.run_batch(torch.from_numpy(lab).to(dev()), feats, [32, 32], 8, 16)
    assert out.check() == 0


def test_label_max(ops):
    labs = np.stack([synth.voronoi_labels(40, 50, 3, 4, image_index=i) for i in range(3)])
    labs[1] = labs[1] % 7
    for dt in (np.int32, np.int64):
        got = ops.label_max(torch.from_numpy(labs.astype(dt)).to(dev())).cpu().numpy()
        assert np.array_equal(got, labs.reshape(3, -1).max(axis=1))


def test_overlap_full_size_properties_and_determinism(ops):
    H, W, fh, fw = 1024, 2048, 128, 256
    labs = synth.voronoi_labels_torch(3, H, W, 25, 40, first_index=0, device=dev())
    n_sp = [1000] * 3
    ov = ops.overlap_csr(labs, fh, fw, n_sp, prior=PRIOR)
    nnz = ov.validate()
    ip = ov.indptr.cpu().numpy().astype(np.int64)
    ix = ov.indices.cpu().numpy()[:nnz]
    ct = ov.counts.cpu().numpy()[:nnz].astype(np.int64)
    assert ct.sum() == 3 * H * W and ct.min() >= 1 and ct.max() <= 64
    rows = np.repeat(np.arange(3000), np.diff(ip))
    assert np.array_equal(np.bincount(rows, weights=ct, minlength=3000).astype(np.int64),
                          ov.area.cpu().numpy())
    srt = np.diff(ix) > 0
    srt[ip[1:-1] - 1] = True  # row boundaries
    assert srt.all() and ix.min() >= 0 and ix.max() < fh * fw
    # each image's cells are covered: column sums = 64 pixels per cell
    for i in range(3):
        a, b = ip[1000 * i], ip[1000 * (i + 1)]
        assert np.array_equal(np.bincount(ix[a:b], weights=ct[a:b], minlength=fh * fw),
                              np.full(fh * fw, 64.0))
    # exact oracle comparison on the first image
    lab0 = labs[0].cpu().numpy()
    oip, oix, oct_ = so.overlap_csr(lab0, fh, fw, 1000)
    assert np.array_equal(ip[:1001], oip) and np.array_equal(ix[:oip[-1]], oix)
    assert np.array_equal(ct[:oip[-1]], oct_)
    np.testing.assert_allclose(ov.weights()[:1000].cpu().numpy(), so.create_prior(lab0, *PRIOR),
                               rtol=1e-12, atol=1e-300)
    # bit-reproducible run to run (fp64 prior included)
    ov2 = ops.overlap_csr(labs, fh, fw, n_sp, prior=PRIOR)
    assert torch.equal(ov.indices[:nnz], ov2.indices[:nnz]) and torch.equal(ov.counts[:nnz], ov2.counts[:nnz])
    assert torch.equal(ov.sum_prior, ov2.sum_prior) and torch.equal(ov.sum_y, ov2.sum_y)


# ---------------------------------------------------------------------------- K2 pooling
@pytest.mark.parametrize('C,layout', [(512, 'nchw'), (512, 'cl'), (24, 'nchw'), (260, 'cl'),
                                      (1024, 'cl')])
def test_pool_matches_oracle(ops, C, layout):
    H, W, fh, fw = 64, 128, 8, 16
    labs = [synth.voronoi_labels(H, W, 4, 8, image_index=i) for i in range(2)]
    feats = np.stack([synth.smooth_features(C, fh, fw, seed=20 + i, radius=1) for i in range(2)])
    ov = _check_overlap(ops, labs, fh, fw)
    f = torch.from_numpy(feats).to(dev())
    if layout == 'cl':
        f = f.contiguous(memory_format=torch.channels_last)
    cell = ops.as_cellmajor(f)
    assert np.array_equal(cell.cpu().numpy(), feats.reshape(2, C, -1).transpose(0, 2, 1))
    for append_pos in (True, False):
        got = ops.pool(cell, ov, append_pos=append_pos).cpu().numpy()
        off = 0
        for i, lab in enumerate(labs):
            oip, oix, oct_ = so.overlap_csr(lab, fh, fw, 32)
            area, sy, sx = so.superpixel_stats(lab, 32)
            want = so.pool_count(oip, oix, oct_, feats[i].reshape(C, -1).T, area, sy, sx,
                                 append_pos)
            np.testing.assert_allclose(got[off:off + 32], want.astype(np.float32), rtol=1e-5,
                                       atol=1e-6)
            if append_pos:  # centroid columns: exact integer sums / area, rounded once
                assert np.array_equal(got[off:off + 32, C], (sy / area).astype(np.float32))
                assert np.array_equal(got[off:off + 32, C + 1], (sx / area).astype(np.float32))
            off += 32


def test_pool_equals_dense_nearest_mean(ops):
    # independent formulation: mean over member pixels of the nearest-upsampled map
    lab = synth.voronoi_labels(50, 70, 3, 4, image_index=2)
    F = synth.smooth_features(8, 7, 9, seed=1, radius=1)
    ov = _check_overlap(ops, [lab], 7, 9)
    got = ops.pool(ops.as_cellmajor(torch.from_numpy(F[None]).to(dev())), ov).cpu().numpy()
    np.testing.assert_allclose(got, so.pool_dense_nearest(lab, F, True), rtol=1e-5, atol=1e-6)


def test_pool_heavy_rows_and_full_size_linearity(ops):
    H, W, fh, fw, C = 1024, 2048, 128, 256, 512
    labs = synth.voronoi_labels_torch(2, H, W, 25, 40, first_index=5, device=dev())
    ov = ops.overlap_csr(labs, fh, fw, [1000, 1000])
    ov.validate()
    g = torch.Generator(device=dev()).manual_seed(1111)
    f1 = torch.randn((2, fh * fw, C), generator=g, device=dev())
    f2 = torch.randn((2, fh * fw, C), generator=g, device=dev())
    p1 = ops.pool(f1, ov, append_pos=False)
    p2 = ops.pool(f2, ov, append_pos=False)
    p12 = ops.pool(f1 + 2 * f2, ov, append_pos=False)
    torch.testing.assert_close(p12, p1 + 2 * p2, rtol=1e-4, atol=1e-5)   # linearity
    ones = ops.pool(torch.ones_like(f1), ov, append_pos=True)
    assert torch.equal(ones[:, :C], torch.ones_like(ones[:, :C]))        # mean of a constant
    assert torch.equal(ops.pool(f1, ov, append_pos=False), p1)           # deterministic
    # weighted mean of pooled rows = global mean of the feature map (checksum of checksums)
    area = ov.area[:1000].double()
    glob = (p1[:1000].double() * area[:, None]).sum(0) / area.sum()
    torch.testing.assert_close(glob, f1[0].double().mean(0), rtol=1e-4, atol=1e-5)
    # one image against the float64 oracle
    lab0 = labs[0].cpu().numpy()
    oip, oix, oct_ = so.overlap_csr(lab0, fh, fw, 1000)
    want = so.pool_count(oip, oix, oct_, f1[0].cpu().numpy(), append_pos=False)
    np.testing.assert_allclose(p1[:1000].cpu().numpy(), want.astype(np.float32), rtol=1e-5, atol=2e-6)
    # single superpixel (one 32768-cell row)
    ov1 = ops.overlap_csr(torch.zeros((1, H, W), dtype=torch.int32, device=dev()), fh, fw, [1])
    one = ops.pool(f1[:1], ov1, append_pos=True)
    torch.testing.assert_close(one[0, :C].double(), f1[0].double().mean(0), rtol=1e-4, atol=1e-5)
    assert one[0, C].item() == 511.5 and one[0, C + 1].item() == 1023.5


# ---------------------------------------------------------------------------- K3 k-means
def _golden_kmeans(golden_dir):
    g = np.load(os.path.join(golden_dir, 'kmeans_ref.npz'))
    names = sorted({k.split('__')[0] for k in g.files})
    return {n: {k.split('__')[1]: g[k] for k in g.files if k.startswith(n + '__')} for n in names}


@pytest.mark.parametrize('xdtype', ['f32', 'f64'])
def test_kmeans_groups_reference_golden(ops, golden_dir, xdtype):
    for name, c in _golden_kmeans(golden_dir).items():
        X = torch.from_numpy(c['X'].astype(np.float32 if xdtype == 'f32' else np.float64)).to(dev())
        goff = torch.tensor([0, len(c['X'])], dtype=torch.int64, device=dev())
        res = ops.kmeans_groups(X, torch.from_numpy(c['w']).to(dev()),
                                torch.from_numpy(c['init']).to(dev()), int(c['k']), goff)
        assert np.array_equal(res.assign.cpu().numpy(), c['assign']), name


def test_kmeans_large_reference_golden(ops, golden_dir):
    for name, c in _golden_kmeans(golden_dir).items():
        X = torch.from_numpy(c['X']).to(dev())
        for chunks in (1, 3):
            km = ops.KMeansLarge(X, torch.from_numpy(c['w']).to(dev()),
                                 torch.from_numpy(c['init']).to(dev()), int(c['k']),
                                 [0, len(c['X'])], chunks_per_group=chunks)
            res = km.run(poll=3)
            assert np.array_equal(res.assign.cpu().numpy(), c['assign']), (name, chunks)


def _blobs(rs, n, d, k, spread=5.0):
    cent = rs.standard_normal((k, d)) * spread
    return (cent[rs.randint(0, k, n)] + rs.standard_normal((n, d))).astype(np.float32)


@pytest.mark.parametrize('D,K', [(514, 4), (512, 4), (66, 8), (1026, 3), (7, 2)])
def test_kmeans_groups_vs_oracle_multi_group(ops, D, K):
    rs = np.random.RandomState(D * 10 + K)
    sizes = [300, 1, 57, 1000, 0, 33]
    off = np.concatenate([[0], np.cumsum(sizes)])
    X = _blobs(rs, off[-1], D, K)
    if D > 500:
        X[:, -2] = rs.uniform(0, 1023, len(X)); X[:, -1] = rs.uniform(0, 2047, len(X))
    w = rs.uniform(0, 1, len(X))
    init = np.concatenate([so.kmeans_init(K, w[a:b], rng=rs) if b > a else np.zeros(0)
                           for a, b in zip(off[:-1], off[1:])]).astype(np.int32)
    res = ops.kmeans_groups(torch.from_numpy(X).to(dev()), torch.from_numpy(w).to(dev()),
                            torch.from_numpy(init).to(dev()), K,
                            torch.from_numpy(off.astype(np.int64)).to(dev()), want_centers=True)
    got = res.assign.cpu().numpy()
    for g, (a, b) in enumerate(zip(off[:-1], off[1:])):
        if b == a:
            continue
        want, info = so.kmeans(K, X[a:b].astype(np.float64), w[a:b], init_assign=init[a:b].astype(np.float64),
                               return_info=True, verbose=False)
        assert np.array_equal(got[a:b], np.asarray(want).astype(np.int32)), (g, a, b)
        assert res.iters[g].item() == info['iters'] and res.status[g].item() == info['status']
        if info['status'] == 1 or np.isnan(info['centers']).any():
            continue
        np.testing.assert_allclose(res.centers[g].cpu().numpy(), info['centers'], rtol=1e-9, atol=1e-9)


def test_kmeans_iteration_cap_and_zero_iters(ops):
    rs = np.random.RandomState(5)
    X = _blobs(rs, 400, 20, 4, spread=1.0)
    w = rs.uniform(0, 1, 400)
    init = so.kmeans_init(4, w, rng=rs).astype(np.int32)
    for cap in (1, 2, 3):
        want, info = so.kmeans(4, X.astype(np.float64), w, n_iter=cap, init_assign=init.astype(np.float64),
                               return_info=True, verbose=False)
        goff = torch.tensor([0, 400], dtype=torch.int64, device=dev())
        res = ops.kmeans_groups(torch.from_numpy(X).to(dev()), torch.from_numpy(w).to(dev()),
                                torch.from_numpy(init).to(dev()), 4, goff, n_iter=cap)
        assert np.array_equal(res.assign.cpu().numpy(), np.asarray(want).astype(np.int32))
        assert res.status[0].item() == info['status'] and res.iters[0].item() == info['iters']
        km = ops.KMeansLarge(torch.from_numpy(X).to(dev()), torch.from_numpy(w).to(dev()),
                             torch.from_numpy(init).to(dev()), 4, [0, 400], n_iter=cap,
                             chunks_per_group=2).run(poll=1)
        assert np.array_equal(km.assign.cpu().numpy(), np.asarray(want).astype(np.int32))
        assert km.status[0].item() == info['status'] and km.iters[0].item() == info['iters']


def test_kmeans_direct_virtual_position_columns(ops):
    # direct clustering: rows are cells, columns C channels + (x, y) cell index generated in-kernel
    n, C, h, w_ = 2, 12, 9, 14
    feats = np.stack([synth.smooth_features(C, h, w_, seed=40 + i, radius=1) for i in range(n)])
    X = so.direct_features(feats)                       # float64 [n*h*w, C+2]
    prior = so.create_prior_map(h, w_, *PRIOR).reshape(1, -1).repeat(n, axis=0).reshape(-1)
    rs = np.random.RandomState(2)
    init = so.kmeans_init(4, prior, rng=rs).astype(np.int32)
    want = so.kmeans(4, X, prior, init_assign=init.astype(np.float64), verbose=False)
    cell = ops.as_cellmajor(torch.from_numpy(feats).to(dev())).reshape(n * h * w_, C)
    goff = torch.tensor([0, n * h * w_], dtype=torch.int64, device=dev())
    res = ops.kmeans_groups(cell, torch.from_numpy(prior).to(dev()), torch.from_numpy(init).to(dev()),
                            4, goff, pos_grid=(h, w_))
    assert np.array_equal(res.assign.cpu().numpy(), np.asarray(want).astype(np.int32))
    km = ops.KMeansLarge(cell, torch.from_numpy(prior).to(dev()), torch.from_numpy(init).to(dev()),
                         4, [0, n * h * w_], pos_grid=(h, w_), chunks_per_group=4).run()
    assert np.array_equal(km.assign.cpu().numpy(), np.asarray(want).astype(np.int32))
    # persistent finish kernel with the virtual columns
    kt = ops.KMeansLarge(cell, torch.from_numpy(prior).to(dev()), torch.from_numpy(init).to(dev()),
                         4, [0, n * h * w_], pos_grid=(h, w_))
    assert kt.tail
    assert np.array_equal(kt.run().assign.cpu().numpy(), np.asarray(want).astype(np.int32))
    # and the materialised matrix through the plain path gives the same answer
    res2 = ops.kmeans_groups(torch.from_numpy(X.astype(np.float32)).to(dev()),
                             torch.from_numpy(prior).to(dev()), torch.from_numpy(init).to(dev()), 4, goff)
    assert torch.equal(res2.assign, res.assign)


def test_kmeans_large_many_groups_and_chunks(ops):
    rs = np.random.RandomState(11)
    sizes = [5000, 9000, 130]
    off = np.concatenate([[0], np.cumsum(sizes)])
    X = _blobs(rs, off[-1], 34, 4)
    w = rs.uniform(0, 1, len(X))
    init = np.concatenate([so.kmeans_init(4, w[a:b], rng=rs) for a, b in zip(off[:-1], off[1:])]).astype(np.int32)
    km = ops.KMeansLarge(torch.from_numpy(X).to(dev()), torch.from_numpy(w).to(dev()),
                         torch.from_numpy(init).to(dev()), 4, off).run()
    got = km.assign.cpu().numpy()
    for g, (a, b) in enumerate(zip(off[:-1], off[1:])):
        want, info = so.kmeans(4, X[a:b].astype(np.float64), w[a:b], init_assign=init[a:b].astype(np.float64),
                               return_info=True, verbose=False)
        assert np.array_equal(got[a:b], np.asarray(want).astype(np.int32))
        assert km.iters[g].item() == info['iters'] and km.status[g].item() == info['status']


def test_kmeans_init_device_matches_host(ops):
    rs = np.random.RandomState(9)
    sizes = [1000, 37, 2, 1, 4096, 513, 4097, 30000, 70001]   # > 4096: radix-select median
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    w = rs.uniform(0, 1, off[-1])
    from superpixel_align_b200 import pipeline
    for k in (2, 4, 8):
        np.random.seed(77)
        flat, soff, m_exp = pipeline.draw_shuffles(k, sizes)
        np.random.seed(77)
        want = np.concatenate([so.kmeans_init(k, w[a:b]) for a, b in zip(off[:-1], off[1:])])
        init, m = ops.kmeans_init_device(torch.from_numpy(w).to(dev()), torch.from_numpy(off).to(dev()),
                                         torch.from_numpy(flat).to(dev()), torch.from_numpy(soff).to(dev()))
        assert np.array_equal(m.cpu().numpy(), m_exp)
        assert np.array_equal(init.cpu().numpy(), want.astype(np.int32))


# ------------------------------------------------------------------- K4 paint, K5 refine, eval
@pytest.mark.parametrize('ldt', [np.int32, np.int64])
def test_paint_matches_reference_golden(ops, golden_dir, ldt):
    g = np.load(os.path.join(golden_dir, 'weighted_kmeans_ref.npz'))
    labs, n_per = g['labs'].astype(ldt), list(g['n_per'])
    assign = so.kmeans(int(g['k']), g['anchor_features'], g['weights'],
                       init_assign=g['init'].astype(np.float64), verbose=False)
    sp_off = torch.tensor(np.concatenate([[0], np.cumsum(n_per)]), dtype=torch.int64, device=dev())
    for odt in (torch.uint8, torch.int32, torch.int64):
        cmap, mask = ops.paint(torch.from_numpy(labs).to(dev()), sp_off,
                               torch.from_numpy(np.asarray(assign).astype(np.int32)).to(dev()), out_dtype=odt)
        assert np.array_equal(cmap.cpu().numpy().astype(np.int64), g['cluster_map'].astype(np.int64))
        assert np.array_equal(mask.cpu().numpy().astype(bool), g['road'])


def test_paint_odd_size_single_image(ops):
    lab = synth.voronoi_labels(33, 47, 3, 3, image_index=1)
    table = np.arange(9, dtype=np.int32)[::-1].copy() % 4
    cmap, mask = ops.paint(torch.from_numpy(lab[None]).to(dev()), torch.tensor([0, 9], device=dev()),
                           torch.from_numpy(table).to(dev()))
    assert np.array_equal(cmap[0].cpu().numpy(), table[lab].astype(np.uint8))
    assert np.array_equal(mask[0].cpu().numpy(), (table[lab] == 0).astype(np.uint8))


def test_refine_matches_mask_loop(ops):
    labs = [synth.voronoi_labels(64, 96, 4, 6, image_index=9 + i) for i in range(2)]
    fh, fw = 8, 12
    rs = np.random.RandomState(0)
    road = rs.rand(2, fh, fw) < 0.3
    road[1] = False  # an image with no predicted road: nothing is kept
    ov = _check_overlap(ops, labs, fh, fw, prior=None)
    for thr in (0.01, 0.05, 0.2):
        overlap, road_px, keep = ops.refine(ov, torch.from_numpy(road).to(dev()), thr)
        _, refined = ops.paint(torch.from_numpy(np.stack(labs)).to(dev()), ov.sp_off, keep,
                               out_dtype=None, road_value=1)
        for i in range(2):
            oip, oix, oct_ = so.overlap_csr(labs[i], fh, fw, 24)
            o_ov, o_px, o_keep = so.refine_overlaps_csr(oip, oix, oct_, road[i], thr)
            assert np.array_equal(overlap[24 * i:24 * (i + 1)].cpu().numpy(), o_ov)
            assert road_px[i].item() == o_px
            assert np.array_equal(keep[24 * i:24 * (i + 1)].cpu().numpy().astype(bool), o_keep)
            want = so.refine_overlaps_masks(labs[i], so.upsample_nearest(road[i], 64, 96), thr)
            assert np.array_equal(refined[i].cpu().numpy(), want)


@pytest.mark.parametrize('h,w,H,W', [(128, 256, 1024, 2048), (28, 28, 224, 224), (28, 28, 1024, 2048),
                                     (224, 224, 100, 333), (1024, 2048, 224, 224), (37, 53, 37, 53)])
def test_resize_nearest_matches_cv2(ops, h, w, H, W):
    cv2 = pytest.importorskip('cv2')
    rs = np.random.RandomState(h + W)
    src = rs.randint(0, 5, size=(2, h, w)).astype(np.uint8)
    got = ops.resize_nearest_u8(torch.from_numpy(src).to(dev()), H, W).cpu().numpy()
    for i in range(2):
        assert np.array_equal(got[i], cv2.resize(src[i], (W, H), interpolation=cv2.INTER_NEAREST))


def test_confusion_matches_oracle(ops):
    rs = np.random.RandomState(3)
    gt = rs.randint(-1, 2, size=(3, 40, 50)).astype(np.int32)
    pred = rs.randint(0, 2, size=(3, 40, 50)).astype(np.uint8)
    conf = ops.confusion2(torch.from_numpy(pred).to(dev()), torch.from_numpy(gt).to(dev())).cpu().numpy()
    for i in range(3):
        assert np.array_equal(conf[i], so.confusion(pred[i], gt[i], 2))


def test_kmeans_exact_ties_and_near_ties_take_first_minimum(ops):
    # rows exactly (or within 1e-7 relative) equidistant from two centres must go through the
    # float64 fallback of the fp32 screening pass and follow NumPy's first-minimum rule
    rs = np.random.RandomState(1)
    D, K, N = 514, 4, 400
    cen = rs.standard_normal((K, D)) * 3
    cen[:, -2:] = rs.uniform(0, 2000, (K, 2))
    cen = cen.astype(np.float32).astype(np.float64)
    X = np.empty((N, D), dtype=np.float64)
    for i in range(N):
        a, b = rs.choice(K, 2, replace=False)
        mid = 0.5 * (cen[a] + cen[b])
        dirv = cen[b] - cen[a]
        orth = rs.standard_normal(D)
        orth -= orth.dot(dirv) / dirv.dot(dirv) * dirv
        eps = [0.0, 1e-7, -1e-7, 1e-4, -1e-4][i % 5]
        X[i] = mid + 0.01 * orth + eps * dirv
    X = X.astype(np.float32)
    w = rs.uniform(0, 1, N)
    dist = np.linalg.norm(X.astype(np.float64)[:, None] - cen[None], axis=2)
    want = dist.argmin(1).astype(np.int32)
    for xd in (np.float32, np.float64):
        km = ops.KMeansLarge(torch.from_numpy(X.astype(xd)).to(dev()), torch.from_numpy(w).to(dev()),
                             torch.zeros(N, dtype=torch.int32, device=dev()), K, [0, N],
                             chunks_per_group=2)
        km.centers.copy_(torch.from_numpy(cen).to(dev())[None])
        km._init_done = True
        km.step()
        assert np.array_equal(km.assign.cpu().numpy(), want)
        assert km.iters[0].item() == 1


def test_kmeans_fused_iterate_equals_three_kernel_path(ops):
    rs = np.random.RandomState(21)
    sizes = [900, 1000, 64, 2500]
    off = np.concatenate([[0], np.cumsum(sizes)])
    X = _blobs(rs, off[-1], 514, 4)
    X[:, -2] = rs.uniform(0, 1023, len(X)); X[:, -1] = rs.uniform(0, 2047, len(X))
    w = rs.uniform(0, 1, len(X))
    init = np.concatenate([so.kmeans_init(4, w[a:b], rng=rs) for a, b in zip(off[:-1], off[1:])]).astype(np.int32)
    args = (torch.from_numpy(X).to(dev()), torch.from_numpy(w).to(dev()), torch.from_numpy(init).to(dev()), 4, off)
    a = ops.KMeansLarge(*args, fused=True, incremental=False).run()
    b = ops.KMeansLarge(*args, fused=False).run(poll=3)
    assert torch.equal(a.assign, b.assign) and torch.equal(a.iters, b.iters)
    assert torch.equal(a.status, b.status) and torch.equal(a.centers, b.centers)   # bit-identical
    c = ops.KMeansLarge(*args, fused=True, incremental=True).run()                  # running sums
    assert torch.equal(a.assign, c.assign) and torch.equal(a.iters, c.iters) and torch.equal(a.status, c.status)
    torch.testing.assert_close(c.centers, a.centers, rtol=1e-11, atol=1e-11)
    for g, (lo, hi) in enumerate(zip(off[:-1], off[1:])):
        want = so.kmeans(4, X[lo:hi].astype(np.float64), w[lo:hi], init_assign=init[lo:hi].astype(np.float64), verbose=False)
        assert np.array_equal(a.assign[lo:hi].cpu().numpy(), np.asarray(want).astype(np.int32))


def test_kmeans_two_level_reduction_tree(ops):
    # groups cut into more than 32 chunks: the partial sums are added along the two-level tree
    # (runs of 32 slots by the chunk of the run that finishes last, then the run sums); the
    # fused kernel, the three-call form and the running-sum mode must agree with the oracle
    rs = np.random.RandomState(4)
    sizes = [9000, 700, 5300]
    off = np.concatenate([[0], np.cumsum(sizes)])
    X = _blobs(rs, off[-1], 34, 4, spread=1.5)
    w = rs.uniform(0, 1, len(X))
    init = np.concatenate([so.kmeans_init(4, w[a:b], rng=rs) for a, b in zip(off[:-1], off[1:])]).astype(np.int32)
    args = (torch.from_numpy(X).to(dev()), torch.from_numpy(w).to(dev()), torch.from_numpy(init).to(dev()), 4, off)
    a = ops.KMeansLarge(*args, chunks_per_group=150, incremental=False).run()
    assert a is not None and (np.diff(ops.KMeansLarge(*args, chunks_per_group=150).gco.cpu().numpy()) > 64).any()
    b = ops.KMeansLarge(*args, chunks_per_group=150, fused=False).run(poll=3)
    assert torch.equal(a.assign, b.assign) and torch.equal(a.iters, b.iters)
    assert torch.equal(a.status, b.status) and torch.equal(a.centers, b.centers)   # bit-identical
    c = ops.KMeansLarge(*args, chunks_per_group=150).run()                          # running sums + bounds
    assert torch.equal(a.assign, c.assign) and torch.equal(a.iters, c.iters) and torch.equal(a.status, c.status)
    for g, (lo, hi) in enumerate(zip(off[:-1], off[1:])):
        want, info = so.kmeans(4, X[lo:hi].astype(np.float64), w[lo:hi], init_assign=init[lo:hi].astype(np.float64),
                               return_info=True, verbose=False)
        assert np.array_equal(a.assign[lo:hi].cpu().numpy(), np.asarray(want).astype(np.int32))
        assert a.iters[g].item() == info['iters'] and a.status[g].item() == info['status']


@pytest.mark.parametrize('sizes,tail_rows', [([900, 1000, 64, 1700, 1], 2048), ([2500, 3100, 1025], 4096)])
def test_kmeans_finish_kernel_equals_per_iteration_launches(ops, sizes, tail_rows, monkeypatch):
    # one persistent CTA per group (spalign_kmeans_finish; groups above 1024 rows loop over
    # sub-chunks) against one launch per iteration, and both against the oracle
    monkeypatch.setattr(ops.KMeansLarge, 'TAIL_ROWS', tail_rows)
    rs = np.random.RandomState(33)
    off = np.concatenate([[0], np.cumsum(sizes)])
    X = _blobs(rs, off[-1], 514, 4)
    X[:, -2] = rs.uniform(0, 1023, len(X)); X[:, -1] = rs.uniform(0, 2047, len(X))
    w = rs.uniform(0, 1, len(X))
    init = np.concatenate([so.kmeans_init(4, w[a:b], rng=rs) if b - a > 1 else np.zeros(b - a)
                           for a, b in zip(off[:-1], off[1:])]).astype(np.int32)
    args = (torch.from_numpy(X).to(dev()), torch.from_numpy(w).to(dev()), torch.from_numpy(init).to(dev()), 4, off)
    km = ops.KMeansLarge(*args)
    assert km.tail
    l0 = ops.LAUNCHES
    a = km.run()
    assert ops.LAUNCHES - l0 == 3          # init sums, first full iteration, finish kernel
    b = ops.KMeansLarge(*args, tail=False).run()
    assert torch.equal(a.assign, b.assign) and torch.equal(a.iters, b.iters)
    assert torch.equal(a.status, b.status)
    # paused after 3 iterations (2 rows per warp set, 2 CTAs per SM), continued by a second launch
    c = ops.KMeansLarge(*args, tail_slice=3).run()
    assert torch.equal(a.assign, c.assign) and torch.equal(a.iters, c.iters)
    assert torch.equal(a.status, c.status)
    for g, (lo, hi) in enumerate(zip(off[:-1], off[1:])):
        if a.status[g].item() == 0:        # centres of converged groups: same up to summation order
            torch.testing.assert_close(a.centers[g], b.centers[g], rtol=1e-10, atol=1e-10)
        want, info = so.kmeans(4, X[lo:hi].astype(np.float64), w[lo:hi], init_assign=init[lo:hi].astype(np.float64),
                               return_info=True, verbose=False)
        assert np.array_equal(a.assign[lo:hi].cpu().numpy(), np.asarray(want).astype(np.int32))
        assert a.iters[g].item() == info['iters'] and a.status[g].item() == info['status']


@pytest.mark.parametrize('D,K', [(66, 8), (1026, 3), (7, 2), (512, 4), (100, 5)])
def test_kmeans_finish_kernel_other_shapes(ops, D, K):
    # kernel variants behind the finish kernel: K > 4 and rows > 1024 columns (2 rows per warp
    # set), short rows, row lengths without centroid columns, more than 3 float64 tail columns
    rs = np.random.RandomState(D + K)
    sizes = [700, 33, 1200]
    off = np.concatenate([[0], np.cumsum(sizes)])
    X = _blobs(rs, off[-1], D, K, spread=1.0)
    w = rs.uniform(0, 1, len(X))
    init = np.concatenate([so.kmeans_init(K, w[a:b], rng=rs) for a, b in zip(off[:-1], off[1:])]).astype(np.int32)
    km = ops.KMeansLarge(torch.from_numpy(X).to(dev()), torch.from_numpy(w).to(dev()),
                         torch.from_numpy(init).to(dev()), K, off)
    assert km.tail
    res = km.run()
    for g, (lo, hi) in enumerate(zip(off[:-1], off[1:])):
        want, info = so.kmeans(K, X[lo:hi].astype(np.float64), w[lo:hi], init_assign=init[lo:hi].astype(np.float64),
                               return_info=True, verbose=False)
        assert np.array_equal(res.assign[lo:hi].cpu().numpy(), np.asarray(want).astype(np.int32))
        assert res.iters[g].item() == info['iters'] and res.status[g].item() == info['status']


def test_kmeans_finish_kernel_iteration_cap(ops):
    rs = np.random.RandomState(5)
    X = _blobs(rs, 1500, 66, 4, spread=0.4)   # heavily overlapping blobs: many iterations
    w = rs.uniform(0, 1, 1500)
    init = so.kmeans_init(4, w, rng=rs).astype(np.int32)
    full = ops.KMeansLarge(torch.from_numpy(X).to(dev()), torch.from_numpy(w).to(dev()),
                           torch.from_numpy(init).to(dev()), 4, [0, 1500]).run()
    n_full = full.iters[0].item()
    assert n_full >= 4
    for cap in (1, 2, n_full - 1):
        km = ops.KMeansLarge(torch.from_numpy(X).to(dev()), torch.from_numpy(w).to(dev()),
                             torch.from_numpy(init).to(dev()), 4, [0, 1500], n_iter=cap).run()
        want, info = so.kmeans(4, X.astype(np.float64), w, n_iter=cap, init_assign=init.astype(np.float64),
                               return_info=True, verbose=False)
        assert km.iters[0].item() == cap == info['iters']
        assert km.status[0].item() == info['status']
        assert np.array_equal(km.assign.cpu().numpy(), np.asarray(want).astype(np.int32))


# ------------------------------------------------------------------- f2 bilinear overlap / pooling
@pytest.mark.parametrize('H,W,fh,fw,gy,gx,dtype', [(64, 128, 8, 16, 4, 8, np.int32), (50, 70, 7, 9, 3, 4, np.int64),
                                                  (224, 224, 28, 28, 7, 7, np.int32), (33, 47, 5, 6, 3, 3, np.int32)])
def test_bilinear_overlap_and_pooling_match_oracle(ops, H, W, fh, fw, gy, gx, dtype):
    labs = [synth.voronoi_labels(H, W, gy, gx, image_index=i, dtype=dtype) for i in range(2)]
    S = gy * gx
    C = 12
    feats = np.stack([synth.smooth_features(C, fh, fw, seed=30 + i, radius=1) for i in range(2)])
    t = torch.from_numpy(np.stack(labs)).to(dev())
    ov = ops.overlap_csr(t, fh, fw, [S, S])
    bw = ops.overlap_bilinear_csr(t, fh, fw, ov)
    nnz = bw.validate()
    ip = bw.indptr.cpu().numpy(); ix = bw.indices.cpu().numpy(); wv = bw.wvals.cpu().numpy()
    assert ip[-1] == nnz
    cell = ops.as_cellmajor(torch.from_numpy(feats).to(dev()))
    got = ops.pool_weighted(cell, ov, bw, append_pos=True).cpu().numpy()
    for i, lab in enumerate(labs):
        oip, oix, owv = so.overlap_bilinear_csr(lab, fh, fw, S)
        a, b = ip[S * i], ip[S * (i + 1)]
        assert np.array_equal(ip[S * i:S * (i + 1) + 1] - a, oip)
        assert np.array_equal(ix[a:b], oix)                       # same sparsity pattern, sorted
        np.testing.assert_allclose(wv[a:b], owv, rtol=1e-12, atol=1e-14)
        area, sy, sx = so.superpixel_stats(lab, S)
        np.testing.assert_allclose(bw.row_weight[S * i:S * (i + 1)].cpu().numpy(), area, rtol=1e-12)
        want = so.pool_dense_bilinear(lab, feats[i])              # notebook formulation
        np.testing.assert_allclose(got[S * i:S * (i + 1), :C], want.astype(np.float32), rtol=1e-5, atol=1e-6)
        assert np.array_equal(got[S * i:S * (i + 1), C], (sy / area).astype(np.float32))
    bw2 = ops.overlap_bilinear_csr(t, fh, fw, ov)
    assert torch.equal(bw.wvals[:nnz], bw2.wvals[:nnz])            # bit-reproducible


# ------------------------------------------------------------------- f2 anchor-sampled pooling
def test_anchor_pooling_matches_oracle_and_reference_golden(ops, golden_dir):
    g = np.load(os.path.join(golden_dir, 'anchors_ref.npz'))
    lab, fm = g['label'], g['feature_map']
    C, fh, fw = fm.shape
    H, W = lab.shape
    S = int(lab.max()) + 1
    t = torch.from_numpy(lab[None]).to(dev())
    ov = ops.overlap_csr(t, fh, fw, [S])
    cell = ops.as_cellmajor(torch.from_numpy(fm[None]).to(dev()))
    for n_select in (1, 10):
        anchors, n_valid = g['anchors_%d' % n_select], g['n_valid_%d' % n_select]
        got = ops.pool_anchors(cell, ov, torch.from_numpy(anchors).to(dev()),
                               torch.from_numpy(n_valid).to(dev()), H, append_pos=True).cpu().numpy()
        want_oracle = so.pool_anchors(fm, H, anchors, n_valid)         # same tie rule: all rows
        np.testing.assert_allclose(got[:, :C], want_oracle, rtol=1e-5, atol=1e-6)
        ref = g['features_%d' % n_select]                               # the reference itself
        rows_ok = g['tie_free_%d' % n_select].all(axis=1)
        np.testing.assert_allclose(got[rows_ok, :C], ref[rows_ok, :C], rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(got[:, C:], ref[:, C:], rtol=1e-6)


@pytest.mark.parametrize('H,W,fh,fw,gy,gx,dtype', [(64, 128, 8, 16, 4, 8, np.int32), (50, 70, 7, 9, 3, 4, np.int64)])
def test_anchor_sampler_draws_distinct_member_pixels(ops, H, W, fh, fw, gy, gx, dtype):
    labs = np.stack([synth.voronoi_labels(H, W, gy, gx, image_index=i, dtype=dtype) for i in range(2)])
    labs[1][labs[1] == 1] = 0          # a big superpixel and a missing id (empty row) in image 1
    S = gy * gx
    t = torch.from_numpy(labs).to(dev())
    ov = ops.overlap_csr(t, fh, fw, [S, S])
    for n_select in (10, 3):
        a, nv = ops.sample_anchors(t, ov, n_select, seed=5)
        a2, _ = ops.sample_anchors(t, ov, n_select, seed=5)
        a3, _ = ops.sample_anchors(t, ov, n_select, seed=6)
        assert torch.equal(a, a2) and not torch.equal(a, a3)
        a, nv = a.cpu().numpy(), nv.cpu().numpy()
        area = ov.area.cpu().numpy()
        assert np.array_equal(nv, np.minimum(area, n_select))
        for r in range(2 * S):
            img, s = divmod(r, S)
            pts = a[r, :nv[r]]
            assert (a[r, nv[r]:] == -1).all()
            assert all(labs[img][y, x] == s for y, x in pts)
            assert len({(int(y), int(x)) for y, x in pts}) == nv[r]
    # roughly uniform over the members: every pixel of a small superpixel is drawn sometimes
    small = int(np.argmin(np.where(ov.area.cpu().numpy()[:S] > 0, ov.area.cpu().numpy()[:S], 1 << 30)))
    hits = {}
    for seed in range(200):
        a, nv = ops.sample_anchors(t, ov, 4, seed=seed)
        for y, x in a[small, :int(nv[small])].cpu().numpy():
            hits[(int(y), int(x))] = hits.get((int(y), int(x)), 0) + 1
    n_px = int(ov.area[small].item())
    expect = 200 * min(4, n_px) / n_px
    assert len(hits) >= 0.9 * n_px and max(hits.values()) < 2.5 * expect + 10   # P(never drawn) ~ e^-4.6
