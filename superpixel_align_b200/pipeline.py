"""Device-resident superpixel-align pipeline: label maps + cell-major features in HBM ->
cluster maps + road masks in HBM, no host synchronisation inside.

This is ``estimate_road_mask`` (batch_spalign_kmeans.py:427-457) for a batch of images with
everything the reference does on the host (S boolean masks per stage) replaced by
K1 overlap -> K2 pooling -> prior weights -> seeded init -> K3 k-means -> K4 paint-back.
The seeded init keeps the reference's stream semantics: the shuffles are drawn on the host
from ``np.random`` in batch order (they depend only on the group sizes), the median split is
done on the device.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib, ops


@dataclass
class PipelineOutput:
    cluster_map: torch.Tensor   # uint8 [n, H, W]
    road_mask: torch.Tensor     # uint8 [n, H, W]
    assign: torch.Tensor        # int32 [n_rows]
    features: torch.Tensor      # float32 [n_rows, D]
    weights: torch.Tensor       # float64 [n_rows]
    iters: torch.Tensor
    status: torch.Tensor
    init_m: torch.Tensor        # int32 [G] actual low-prior counts (tie check)
    overlap: ops.Overlap
    group_off_host: np.ndarray
    shuf_sizes: np.ndarray


def draw_shuffles(k: int, group_sizes: Sequence[int]):
    """Host side of the seeded init (batch_spalign_kmeans.py:146-148): for every group, in
    order, ``idx = arange(m) % (k-1) + 1; np.random.shuffle(idx)`` with m = N//2 + 1 (the
    number of rows with w <= upper median when the weights are distinct)."""
    chunks, sizes = [], []
    for n in group_sizes:
        m = int(n) // 2 + 1 if n > 0 else 0
        idx = np.arange(m) % (k - 1) + 1
        np.random.shuffle(idx)
        chunks.append(idx.astype(np.int32))
        sizes.append(m)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    flat = np.concatenate(chunks) if chunks else np.zeros(0, np.int32)
    return flat, off, np.asarray(sizes)


def run_batch(labels: torch.Tensor, feat_cellmajor: torch.Tensor, n_sp: Sequence[int], fh: int,
              fw: int, k: int = 4, prior=(0.75, 0.5, 0.1, 0.1), append_pos: bool = True,
              images_per_group: int = 1, n_iter: int = 1000,
              nnz_cap_per_image: Optional[int] = None, out_dtype=torch.uint8,
              timers: Optional[dict] = None, kmeans_impl: str = 'chunks') -> PipelineOutput:
    """One pass of the hot path over a batch.  ``images_per_group`` = the reference's
    ``--batchsize`` (superpixels of that many consecutive images are clustered jointly;
    1 = per-image clustering).  Groups of more than 4096 rows use the host-driven multi-CTA
    k-means (which polls a stop flag); everything else is sync-free."""
    n = labels.shape[0]
    n_sp = np.asarray(n_sp, dtype=np.int64)

    def mark(name, first):
        # optional CUDA-event brackets per stage (bench.py's roofline numbers)
        if timers is None:
            return
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        if first:
            timers[name] = [ev, None]
        else:
            timers[name][1] = ev

    mark('overlap', True)
    ov = ops.overlap_csr(labels, fh, fw, n_sp, prior=prior, nnz_cap_per_image=nnz_cap_per_image)
    mark('overlap', False)
    mark('pool', True)
    feats = ops.pool(feat_cellmajor, ov, append_pos=append_pos)
    mark('pool', False)
    weights = ov.weights()
    g_idx = np.arange(0, n + 1, images_per_group)
    if g_idx[-1] != n:
        g_idx = np.append(g_idx, n)
    group_off_host = ov.sp_off_host[g_idx]
    sizes = np.diff(group_off_host)
    dev = labels.device
    if sizes.max() <= 4096:
        flat, off, m_exp = draw_shuffles(k, sizes)
        goff = ov.sp_off if images_per_group == 1 else \
            torch.from_numpy(group_off_host).to(dev, non_blocking=True)
        flat_d = torch.from_numpy(flat).to(dev, non_blocking=True)
        off_d = torch.from_numpy(off).to(dev, non_blocking=True)
        mark('init', True)
        init, m = ops.kmeans_init_device(weights, goff, flat_d, off_d)
        mark('init', False)
        mark('kmeans', True)
        if kmeans_impl == 'groups':
            res = ops.kmeans_groups(feats, weights, init, k, goff, n_iter=n_iter)
        else:  # many CTAs per image, finished images dropped as the iterations go on
            res = ops.KMeansLarge(feats, weights, init, k, group_off_host, n_iter=n_iter).run()
        mark('kmeans', False)
    else:
        # large joint groups: the median split needs a sort of N doubles -> host init
        w_host = weights.cpu().numpy()
        init_host = np.zeros(len(w_host), dtype=np.int32)
        m_list = []
        for g in range(len(sizes)):
            a, b = group_off_host[g], group_off_host[g + 1]
            wg = w_host[a:b]
            thr = float(np.sort(wg)[len(wg) // 2])
            low = wg <= thr
            idx = np.arange(int(low.sum())) % (k - 1) + 1
            np.random.shuffle(idx)
            init_host[a:b][low] = idx
            m_list.append(int(low.sum()))
        m = torch.tensor(m_list, dtype=torch.int32)
        m_exp = np.asarray(m_list)
        mark('kmeans', True)
        res = ops.KMeansLarge(feats, weights, torch.from_numpy(init_host).to(dev), k,
                              group_off_host, n_iter=n_iter).run()
        mark('kmeans', False)
    mark('paint', True)
    cmap, mask = ops.paint(labels, ov.sp_off, res.assign, out_dtype=out_dtype)
    mark('paint', False)
    return PipelineOutput(cmap, mask, res.assign, feats, weights, res.iters, res.status, m, ov,
                          group_off_host, m_exp)
