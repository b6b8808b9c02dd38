"""smoke(): one small invocation of the hot path on cuda:0, checked against the CPU oracle.
Lives under tests/ (not in the product package) because it imports the oracle, which is test
infrastructure; __graft_entry__.smoke() is its only caller."""
from __future__ import annotations

import numpy as np
import torch


def run_smoke(H=256, W=512, C=64, gy=8, gx=12, k=4, verbose=True):
    from oracle import spalign_oracle as so
    from superpixel_align_b200 import _lib, ops, pipeline, synth
    _lib.load()
    assert torch.cuda.is_available(), 'smoke() needs a CUDA device'
    dev = torch.device('cuda', 0)
    fh, fw = H // 8, W // 8
    labs = np.stack([synth.voronoi_labels(H, W, gy, gx, image_index=i) for i in range(2)])
    feats = np.stack([synth.smooth_features(C, fh, fw, seed=i) for i in range(2)])
    n_sp = [gy * gx] * 2
    lab_d = torch.from_numpy(labs).to(dev)
    feat_d = ops.as_cellmajor(torch.from_numpy(feats).to(dev))  # NCHW -> transpose kernel
    np.random.seed(1111)
    out = pipeline.run_batch(lab_d, feat_d, n_sp, fh, fw, k=k)
    torch.cuda.synchronize()
    nnz = out.overlap.validate()
    ip = out.overlap.indptr.cpu().numpy()
    ix = out.overlap.indices.cpu().numpy()[:nnz]
    ct = out.overlap.counts.cpu().numpy()[:nnz]
    np.random.seed(1111)
    off = 0
    for i in range(2):
        oip, oix, oct_ = so.overlap_csr(labs[i], fh, fw, n_sp[i])
        a, b = ip[off], ip[off + n_sp[i]]
        assert np.array_equal(ip[off:off + n_sp[i] + 1] - a, oip), 'indptr mismatch'
        assert np.array_equal(ix[a:b], oix) and np.array_equal(ct[a:b], oct_), 'CSR mismatch'
        area, sy, sx = so.superpixel_stats(labs[i], n_sp[i])
        of = so.pool_count(oip, oix, oct_, feats[i].reshape(C, -1).T, area, sy, sx, True)
        gf = out.features[off:off + n_sp[i]].cpu().numpy()
        np.testing.assert_allclose(gf, of.astype(np.float32), rtol=1e-5, atol=1e-6)
        ow = so.create_prior(labs[i], 0.75, 0.5, 0.1, 0.1)
        gw = out.weights[off:off + n_sp[i]].cpu().numpy()
        np.testing.assert_allclose(gw, ow, rtol=1e-12, atol=1e-300)
        init = so.kmeans_init(k, gw)
        oa = so.kmeans(k, gf.astype(np.float64), gw, init_assign=init, verbose=False)
        ga = out.assign[off:off + n_sp[i]].cpu().numpy()
        assert np.array_equal(np.asarray(oa).astype(np.int32), ga), 'k-means assignment mismatch'
        ocm, orm = so.weighted_kmeans_paint(labs[i][None], oa, [n_sp[i]])
        assert np.array_equal(out.cluster_map[i].cpu().numpy(), ocm[0].astype(np.uint8))
        assert np.array_equal(out.road_mask[i].cpu().numpy().astype(bool), orm[0])
        off += n_sp[i]
    if verbose:
        print('smoke: nnz=%d iters=%s status=%s launches=%d' % (
            nnz, out.iters.tolist(), out.status.tolist(), ops.LAUNCHES))
