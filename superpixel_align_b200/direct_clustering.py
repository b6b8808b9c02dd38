"""Drop-in for the hot-path functions of the reference's ``direct_clustering.py``: k-means
directly on the feature cells of a batch, rows = (n, y, x) cells, columns = C channels +
the (x, y) cell index (direct_clustering.py:286-317)."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, ops
from .batch_spalign_kmeans import (_device, _host_init, _kmeans_device, _report, _to_dev, _unwrap,
                                   kmeans, weighted_average)

__all__ = ['create_prior', 'kmeans', 'weighted_average', 'batch_weighted_kmeans',
           'cluster_cells', 'estimate_road_mask']


def create_prior(h, w, y_rel_pos=0.75, x_rel_pos=0.5, y_rel_sigma=0.1, x_rel_sigma=0.2):
    """Gaussian road prior on the h x w cell grid, float64 [h, w]
    (direct_clustering.py:188-201; the reference also evaluates this on the host)."""
    ys, xs = np.arange(h)[:, None], np.arange(w)[None, :]
    y0, x0 = int(h * y_rel_pos), int(w * x_rel_pos)
    sy, sx = h * y_rel_sigma, w * x_rel_sigma
    return np.exp(-((ys - y0) ** 2 / (2 * sy) ** 2 + (xs - x0) ** 2 / (2 * sx) ** 2))


def batch_weighted_kmeans(args, feature_maps, weights):
    """Assignment [N] for the materialised [N, C+2] matrix (direct_clustering.py:204-208)."""
    return kmeans(k=args.n_clusters, X=feature_maps, weights=weights)


def cluster_cells(feature_maps, weights, k, n_iter=1000, init_assign=None):
    """Same result as building the [n*h*w, C+2] matrix of direct_clustering.py:297-303 and
    calling kmeans(), without materialising it: the rows are read straight from the
    cell-major feature map and the (x, y) columns are generated in the kernel.

    feature_maps [n, C, h, w] (NumPy or torch, any memory format); weights [n*h*w] float64.
    Returns the assignment reshaped to [n, h, w]."""
    fm = _unwrap(feature_maps)
    as_numpy = not isinstance(fm, torch.Tensor)
    dev = _device() if as_numpy else fm.device
    fm = _to_dev(fm, dev, torch.float32)
    n, C, h, w = fm.shape
    cell = ops.as_cellmajor(fm).reshape(n * h * w, C)
    wu = _unwrap(weights)
    w_host = (wu.detach().cpu().numpy() if isinstance(wu, torch.Tensor) else np.asarray(wu))
    w_host = w_host.astype(np.float64).reshape(-1)
    init_host = _host_init(k, w_host) if init_assign is None else \
        np.asarray(init_assign, dtype=np.int32).reshape(-1)
    res = _kmeans_device(k, cell, torch.from_numpy(w_host).to(dev),
                         torch.from_numpy(init_host).to(dev), n_iter, pos_grid=(h, w))
    _report(res.status.cpu().numpy())
    out = res.assign.reshape(n, h, w)
    return out.cpu().numpy() if as_numpy else out


def estimate_road_mask(feature_maps, args):
    """Compute part of direct_clustering.estimate_road_mask (:286-317): returns
    (clustering_results [n,h,w] int32, road_masks bool [n,h,w])."""
    fm = _unwrap(feature_maps)
    n, C, h, w = fm.shape
    prior = create_prior(h, w, args.y_rel_pos, args.x_rel_pos, args.y_rel_sigma, args.x_rel_sigma)
    prior = prior.reshape(1, h * w).repeat(n, axis=0).reshape(n * h * w)
    res = cluster_cells(fm, prior, args.n_clusters)
    return res, res == 0
