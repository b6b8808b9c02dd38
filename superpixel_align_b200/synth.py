"""Synthetic Cityscapes-shaped inputs.

There is no dataset, no pretrained DRN and no scikit-image in this environment, so the
label maps and feature maps the hot path consumes are generated:

* ``voronoi_labels`` -- SLIC-shaped superpixels: a jittered-grid Voronoi diagram with
  contiguous ids 0..S-1 (stand-in for ``skimage.segmentation.slic`` at
  batch_spalign_kmeans.py:308-311).
* ``blob_labels`` -- felzenszwalb-shaped superpixels: heavy-tailed region sizes.
* ``smooth_features`` -- spatially smooth feature maps for CPU-sized tests.
* ``smooth_images`` / ``drn_features`` (torch) -- low-passed noise images pushed through
  a random-init DRN-C-26 (``superpixel_align_b200.drn``), layer8, stride 8, 512 channels.

NumPy versions run anywhere; the ``*_torch`` versions build the same tensors on a CUDA
device for the bench.
"""
from __future__ import annotations

import numpy as np

SLIC_GRIDS = {500: (20, 25), 1000: (25, 40), 2000: (40, 50), 4000: (50, 80)}


def _voronoi_seeds(H, W, gy, gx, seed, jitter):
    rs = np.random.RandomState(seed)
    py, px = H / gy, W / gx
    j = jitter * min(py, px)
    off = rs.uniform(-j, j, size=(gy, gx, 2))
    sy = (np.arange(gy)[:, None] + 0.5) * py + off[..., 0]
    sx = (np.arange(gx)[None, :] + 0.5) * px + off[..., 1]
    return sy, sx, py, px


def voronoi_labels(H=1024, W=2048, gy=25, gx=40, image_index=0, jitter=0.3,
                   dtype=np.int32):
    """Jittered-grid Voronoi label map, ids = gi*gx + gj, every id present."""
    sy, sx, py, px = _voronoi_seeds(H, W, gy, gx, 1111 + image_index, jitter)
    y = np.arange(H, dtype=np.float64)[:, None]
    x = np.arange(W, dtype=np.float64)[None, :]
    gi = np.minimum((y / py).astype(np.int64), gy - 1)
    gj = np.minimum((x / px).astype(np.int64), gx - 1)
    best = np.full((H, W), np.inf)
    lab = np.zeros((H, W), dtype=np.int64)
    for di in (-1, 0, 1):
        for dj in (-1, 0, 1):
            ci = np.clip(gi + di, 0, gy - 1)
            cj = np.clip(gj + dj, 0, gx - 1)
            ci, cj = np.broadcast_arrays(ci, cj)
            d = (y - sy[ci, cj]) ** 2 + (x - sx[ci, cj]) ** 2
            upd = d < best
            best = np.where(upd, d, best)
            lab = np.where(upd, ci * gx + cj, lab)
    assert len(np.unique(lab)) == gy * gx, 'voronoi generator lost a superpixel id'
    return lab.astype(dtype)


def voronoi_labels_torch(n_img, H=1024, W=2048, gy=25, gx=40, first_index=0, jitter=0.3,
                         device='cuda', dtype=None):
    """Same construction on a CUDA device: int32 [n_img, H, W]."""
    import torch
    dtype = torch.int32 if dtype is None else dtype
    out = torch.empty((n_img, H, W), dtype=dtype, device=device)
    y = torch.arange(H, dtype=torch.float64, device=device)[:, None]
    x = torch.arange(W, dtype=torch.float64, device=device)[None, :]
    for n in range(n_img):
        sy, sx, py, px = _voronoi_seeds(H, W, gy, gx, 1111 + first_index + n, jitter)
        sy_t = torch.from_numpy(sy).to(device)
        sx_t = torch.from_numpy(np.ascontiguousarray(sx)).to(device)
        gi = torch.clamp((y / py).long(), max=gy - 1)
        gj = torch.clamp((x / px).long(), max=gx - 1)
        best = torch.full((H, W), float('inf'), dtype=torch.float64, device=device)
        lab = torch.zeros((H, W), dtype=torch.int64, device=device)
        for di in (-1, 0, 1):
            for dj in (-1, 0, 1):
                ci = torch.clamp(gi + di, 0, gy - 1).expand(H, W)
                cj = torch.clamp(gj + dj, 0, gx - 1).expand(H, W)
                d = (y - sy_t[ci, cj]) ** 2 + (x - sx_t[ci, cj]) ** 2
                upd = d < best
                best = torch.where(upd, d, best)
                lab = torch.where(upd, ci * gx + cj, lab)
        out[n] = lab.to(dtype)
    return out


def blob_labels(H, W, n_regions, seed=0, dtype=np.int32):
    """Felzenszwalb-shaped map: random seeds with heavy-tailed weights (weighted Voronoi),
    relabelled to contiguous ids.  Regions may be large and irregular."""
    rs = np.random.RandomState(seed)
    sy = rs.uniform(0, H, n_regions)
    sx = rs.uniform(0, W, n_regions)
    wgt = rs.pareto(1.5, n_regions) + 0.2
    y = np.arange(H, dtype=np.float64)[:, None, None]
    x = np.arange(W, dtype=np.float64)[None, :, None]
    lab = np.zeros((H, W), dtype=np.int64)
    chunk = max(1, (1 << 22) // max(1, W * n_regions))
    for y0 in range(0, H, chunk):
        ys = y[y0:y0 + chunk]
        d = ((ys - sy) ** 2 + (x - sx) ** 2) / wgt
        lab[y0:y0 + chunk] = np.argmin(d, axis=2)
    _, inv = np.unique(lab, return_inverse=True)
    return inv.reshape(H, W).astype(dtype)


def noise_labels(H, W, n_sp, seed=0, dtype=np.int32):
    """Pathological map: iid labels per pixel (every cell holds many labels)."""
    rs = np.random.RandomState(seed)
    lab = rs.randint(0, n_sp, size=(H, W))
    lab.ravel()[:n_sp] = np.arange(n_sp)  # every id present
    return lab.astype(dtype)


def _box_filter(a, r, axes):
    for ax in axes:
        n = a.shape[ax]
        c = np.cumsum(np.concatenate([np.zeros_like(np.take(a, [0], axis=ax)), a], axis=ax), axis=ax)
        lo = np.clip(np.arange(n) - r, 0, n)
        hi = np.clip(np.arange(n) + r + 1, 0, n)
        a = (np.take(c, hi, axis=ax) - np.take(c, lo, axis=ax)) / \
            (hi - lo).reshape([-1 if i == ax else 1 for i in range(a.ndim)])
    return a


def smooth_features(C, fh, fw, seed=0, radius=3, n_modes=3):
    """Spatially smooth float32 [C, fh, fw] features with a few large-scale modes so that
    k-means has structure to find."""
    rs = np.random.RandomState(seed)
    f = _box_filter(rs.standard_normal((C, fh, fw)), radius, (1, 2)) * (2 * radius + 1)
    yy = np.linspace(0, 1, fh)[None, :, None]
    xx = np.linspace(0, 1, fw)[None, None, :]
    for m in range(n_modes):
        amp = rs.standard_normal((C, 1, 1)) * 2.0
        cy, cx, s = rs.uniform(0.2, 0.9), rs.uniform(0.1, 0.9), rs.uniform(0.15, 0.4)
        f = f + amp * np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * s * s))
    return np.ascontiguousarray(f, dtype=np.float32)


def smooth_images_torch(n_img, H=1024, W=2048, first_index=0, box=31, device='cuda'):
    """uint8-valued U[0,255] images low-passed with a ``box`` px box filter, float32
    [n_img, 3, H, W] in 0..255 (the dtype/range ``concat_examples`` hands to the model,
    batch_spalign_kmeans.py:545)."""
    import torch
    import torch.nn.functional as F
    out = torch.empty((n_img, 3, H, W), dtype=torch.float32, device=device)
    for n in range(n_img):
        g = torch.Generator(device=device)
        g.manual_seed(1111 + first_index + n)
        img = torch.randint(0, 256, (1, 3, H, W), generator=g, device=device).float()
        # coarse structure so the low-passed image is not flat grey
        coarse = torch.rand((1, 3, H // 64, W // 64), generator=g, device=device) * 255.0
        img = 0.5 * img + 0.5 * F.interpolate(coarse, size=(H, W), mode='bilinear',
                                              align_corners=False)
        img = F.avg_pool2d(img, box, stride=1, padding=box // 2, count_include_pad=False)
        out[n] = img[0]
    return out


def drn_features_torch(model, imgs, channels_last=True):
    """DRN layer8 features for a float32 0..255 batch, normalised like
    models/drn.py:304-325 (``/255``, ImageNet mean/std).  Returns [n, 512, H/8, W/8]
    float32, physically channels_last (= cell-major) when ``channels_last``."""
    import torch
    mean = torch.tensor([0.485, 0.456, 0.406], device=imgs.device).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225], device=imgs.device).view(1, 3, 1, 1)
    with torch.no_grad():
        x = (imgs / 255.0 - mean) / std
        if channels_last:
            x = x.contiguous(memory_format=torch.channels_last)
        f = model(x)
        if channels_last:
            f = f.contiguous(memory_format=torch.channels_last)
        else:
            f = f.contiguous()
    return f.float()
