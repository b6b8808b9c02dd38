"""Drop-in for the hot path of the reference's ``superpixel_overlaps.py``: direct cell
clustering followed by keeping every superpixel whose overlap with the predicted road mask
exceeds a threshold (superpixel_overlaps.py:309-369, method of arXiv:1708.06118)."""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from .batch_spalign_kmeans import (_batch_state, _device, _to_dev, _unwrap,
                                   batch_superpixel)  # superpixel_overlaps.py:292-306, same function
from .direct_clustering import (batch_weighted_kmeans, cluster_cells, create_prior, kmeans,
                                weighted_average)

__all__ = ['create_prior', 'kmeans', 'weighted_average', 'batch_weighted_kmeans',
           'batch_superpixel', 'refine_road_masks', 'estimate_road_mask']


def refine_road_masks(superpixels, road_masks, overlap_threshold=0.01):
    """refined_roadmap [N, H, W] uint8 -- the loop at superpixel_overlaps.py:359-369.

    superpixels [N, H, W] label maps; road_masks [N, h, w] CELL-level masks (the reference
    nearest-upsamples them to the label shape first, :360-362; counting pixels per
    (superpixel, cell) is the same thing without the upsample)."""
    sp = _unwrap(superpixels)
    as_numpy = not isinstance(sp, torch.Tensor)
    dev = _device() if as_numpy else sp.device
    rm = _to_dev(road_masks, dev)
    if rm.dim() == 2:
        rm = rm[None]
    h, w = rm.shape[-2:]
    st = _batch_state(superpixels, dev, h, w, None)
    _, _, keep = ops.refine(st.ov, rm, overlap_threshold)
    _, refined = ops.paint(st.labels, st.ov.sp_off, keep, out_dtype=None, want_mask=True,
                           road_value=1)
    return refined.cpu().numpy() if as_numpy else refined


def estimate_road_mask(feature_maps, superpixels, args):
    """Compute part of superpixel_overlaps.estimate_road_mask (:309-369): returns
    (refined_roadmaps [N,H,W] uint8, clustering_results [n,h,w], road_masks bool [n,h,w])."""
    fm = _unwrap(feature_maps)
    n, C, h, w = fm.shape
    prior = create_prior(h, w, args.y_rel_pos, args.x_rel_pos, args.y_rel_sigma, args.x_rel_sigma)
    prior = prior.reshape(1, h * w).repeat(n, axis=0).reshape(n * h * w)
    cres = cluster_cells(fm, prior, args.n_clusters)
    road = cres == 0
    refined = refine_road_masks(superpixels, road, args.overlap_threshold)
    return refined, cres, road
