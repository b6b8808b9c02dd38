"""CPU oracle for the superpixel-align hot path.  TEST INFRASTRUCTURE ONLY.

This file is a NumPy restatement of the reference algorithm (pfnet-research/
superpixel-align).  It is the *checker*: only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The
product (``superpixel_align_b200``) never imports anything from ``oracle/`` and has
no CPU fallback.

Pinning status (see DESIGN.md "Oracle"):
  * ``kmeans``, ``kmeans_init``, ``weighted_average``, ``create_prior``,
    ``create_prior_map``, ``weighted_kmeans_paint`` are pinned against the reference's
    own functions executed unmodified (``oracle/ref_extract.py`` AST-extracts them from
    ``/root/reference`` in the authoring container; ``oracle/gen_golden.py`` froze their
    outputs into ``tests/golden/*.npz``).
  * ``overlap_csr`` / ``pool_count`` have NO in-tree reference source
    (``notebooks/Efficient_Superpixel_Align.ipynb`` is listed in
    ``.MISSING_LARGE_BLOBS``): **parity unpinned by the reference** for these two; the
    contract is the one fixed in SURVEY.md section 8 (a1, a2).  They are cross-checked
    against the reference's mask-based loops (``superpixel_overlaps.py:359-369`` and
    ``scipy.ndimage.center_of_mass`` as used at ``batch_spalign_kmeans.py:229``).
  * ``confusion`` / ``road_iou`` restate chainercv 0.7/0.8 semantics from memory
    (chainercv is not in the tree): parity unpinned.

All line numbers refer to files under ``/root/reference``.
"""
from __future__ import annotations

import numpy as np

# --------------------------------------------------------------------------------------
# a1. overlap (count) matrix
# --------------------------------------------------------------------------------------


def cell_of_pixel(n_pix: int, n_cell: int) -> np.ndarray:
    """Nearest-neighbour pixel -> cell map, ``min(floor(p * n_cell / n_pix), n_cell-1)``.

    Identical to cv2.resize(..., INTER_NEAREST) source-index selection, which is how the
    reference upsamples a cell mask to pixel resolution (superpixel_overlaps.py:360-362).
    """
    p = np.arange(n_pix, dtype=np.int64)
    return np.minimum((p * n_cell) // n_pix, n_cell - 1).astype(np.int64)


def overlap_csr(label: np.ndarray, fh: int, fw: int, n_sp: int | None = None):
    """CSR overlap matrix M[s, c] = #pixels with label s that fall in feature cell c.

    Contract: rows = labels 0..S-1, columns ascending cell id
    (cell = cy * fw + cx), int32 counts.  Returns (indptr[S+1], indices[nnz], counts[nnz])
    as int32 arrays.
    """
    label = np.asarray(label)
    H, W = label.shape
    if n_sp is None:
        n_sp = int(label.max()) + 1
    cy = cell_of_pixel(H, fh)
    cx = cell_of_pixel(W, fw)
    cell = cy[:, None] * fw + cx[None, :]
    nc = fh * fw
    key = label.astype(np.int64).ravel() * nc + cell.ravel()
    uk, cnt = np.unique(key, return_counts=True)
    rows = uk // nc
    cols = uk % nc
    indptr = np.zeros(n_sp + 1, dtype=np.int64)
    np.add.at(indptr, rows + 1, 1)
    indptr = np.cumsum(indptr)
    return indptr.astype(np.int32), cols.astype(np.int32), cnt.astype(np.int32)


def superpixel_stats(label: np.ndarray, n_sp: int | None = None):
    """Per-superpixel pixel count, sum of y and sum of x (exact integers).

    centroid = (sum_y/area, sum_x/area) equals scipy.ndimage.center_of_mass(mask) used
    at batch_spalign_kmeans.py:229.
    """
    label = np.asarray(label)
    H, W = label.shape
    if n_sp is None:
        n_sp = int(label.max()) + 1
    flat = label.ravel().astype(np.int64)
    area = np.bincount(flat, minlength=n_sp).astype(np.int64)
    yy = np.repeat(np.arange(H, dtype=np.int64), W)
    xx = np.tile(np.arange(W, dtype=np.int64), H)
    # bincount with integer-valued float64 weights is exact below 2**53
    sum_y = np.bincount(flat, weights=yy, minlength=n_sp).astype(np.int64)
    sum_x = np.bincount(flat, weights=xx, minlength=n_sp).astype(np.int64)
    return area, sum_y, sum_x


# --------------------------------------------------------------------------------------
# a3. prior
# --------------------------------------------------------------------------------------


def create_prior_map(h, w, y_rel_pos=0.75, x_rel_pos=0.5, y_rel_sigma=0.1, x_rel_sigma=0.2):
    """Gaussian road prior on an h x w grid (direct_clustering.py:188-201 and the
    pixel-level map inside batch_spalign_kmeans.py:111-122).  Note ``(2*sigma)**2``."""
    xcoord, ycoord = np.meshgrid(np.arange(w), np.arange(h))
    ymean, xmean = int(h * y_rel_pos), int(w * x_rel_pos)
    y_sigma = h * y_rel_sigma
    x_sigma = w * x_rel_sigma
    return np.exp(-((ycoord - ymean) ** 2 / (2 * y_sigma) ** 2
                    + (xcoord - xmean) ** 2 / (2 * x_sigma) ** 2))


def prior_axes(h, w, y_rel_pos=0.75, x_rel_pos=0.5, y_rel_sigma=0.1, x_rel_sigma=0.2):
    """Separable factors gy[h], gx[w] with create_prior_map == outer(gy, gx) up to 1 ulp."""
    ymean, xmean = int(h * y_rel_pos), int(w * x_rel_pos)
    y_sigma = h * y_rel_sigma
    x_sigma = w * x_rel_sigma
    gy = np.exp(-((np.arange(h) - ymean) ** 2 / (2 * y_sigma) ** 2))
    gx = np.exp(-((np.arange(w) - xmean) ** 2 / (2 * x_sigma) ** 2))
    return gy, gx


def create_prior(superpixels, y_rel_pos=0.75, x_rel_pos=0.5, y_rel_sigma=0.1,
                 x_rel_sigma=0.2):
    """Mean prior weight per superpixel, sorted-label order
    (batch_spalign_kmeans.py:111-129).  float64 [S]."""
    superpixels = np.asarray(superpixels)
    h, w = superpixels.shape
    weights = create_prior_map(h, w, y_rel_pos, x_rel_pos, y_rel_sigma, x_rel_sigma)
    ids, inv = np.unique(superpixels, return_inverse=True)
    inv = inv.ravel()
    s = np.bincount(inv, weights=weights.ravel(), minlength=len(ids))
    n = np.bincount(inv, minlength=len(ids))
    return s / n


# --------------------------------------------------------------------------------------
# a2. count pooling
# --------------------------------------------------------------------------------------


def pool_count(indptr, indices, counts, feat_cellmajor, area=None, sum_y=None, sum_x=None,
               append_pos=True):
    """feat[s, :C] = sum_c M[s,c] * F[c,:] / area[s]; optional (sum_y/area, sum_x/area).

    ``feat_cellmajor`` is [Nc, C] (= feature_map[C, fh, fw].reshape(C, -1).T).  float64 out.
    Contract fixed in SURVEY 8 a2 (count pooling = mean over all member pixels of the
    nearest-upsampled feature map); centroid columns follow batch_spalign_kmeans.py:229,270.
    """
    indptr = np.asarray(indptr, dtype=np.int64)
    indices = np.asarray(indices, dtype=np.int64)
    counts = np.asarray(counts, dtype=np.float64)
    F = np.asarray(feat_cellmajor, dtype=np.float64)
    S = len(indptr) - 1
    C = F.shape[1]
    out = np.zeros((S, C + (2 if append_pos else 0)), dtype=np.float64)
    rowsum = np.zeros(S, dtype=np.float64)
    for s in range(S):
        a, b = indptr[s], indptr[s + 1]
        if b > a:
            out[s, :C] = counts[a:b] @ F[indices[a:b]]
            rowsum[s] = counts[a:b].sum()
    if area is None:
        area = rowsum
    area = np.asarray(area, dtype=np.float64)
    with np.errstate(invalid='ignore', divide='ignore'):
        out[:, :C] /= area[:, None]
        if append_pos:
            out[:, C] = np.asarray(sum_y, dtype=np.float64) / area
            out[:, C + 1] = np.asarray(sum_x, dtype=np.float64) / area
    return out


def pool_dense_nearest(label, feature_map, append_pos=True):
    """Independent formulation of count pooling: nearest-upsample the [C, fh, fw] map to
    H x W and take the mean over each superpixel's pixels.  O(H*W*C) -- small inputs only.
    Used to cross-check ``overlap_csr`` + ``pool_count``."""
    label = np.asarray(label)
    H, W = label.shape
    C, fh, fw = feature_map.shape
    cy = cell_of_pixel(H, fh)
    cx = cell_of_pixel(W, fw)
    up = np.asarray(feature_map, dtype=np.float64)[:, cy][:, :, cx]  # [C, H, W]
    S = int(label.max()) + 1
    out = np.zeros((S, C + (2 if append_pos else 0)))
    for s in range(S):
        m = label == s
        out[s, :C] = up[:, m].mean(axis=1)
        if append_pos:
            ys, xs = np.nonzero(m)
            out[s, C] = ys.mean()
            out[s, C + 1] = xs.mean()
    return out


# --------------------------------------------------------------------------------------
# a4, a5. prior-weighted k-means
# --------------------------------------------------------------------------------------


def weighted_average(a, b, axis=0):
    """batch_spalign_kmeans.py:132-133."""
    return (a * b[:, None]).sum(axis=axis) / b.sum(axis=axis)


def kmeans_init(k, weights, rng=None):
    """Initial assignment (batch_spalign_kmeans.py:141-149).

    Rows whose prior weight is above the upper median go to cluster 0; the rest get
    1..k-1 round-robin, shuffled with the (legacy, process-global) NumPy stream.
    Returns float64 [N] exactly like the reference's ``assign`` array.
    """
    rng = np.random if rng is None else rng
    weights = np.asarray(weights)
    n = weights.shape[0]
    assign = np.zeros((n,))
    thr = float(np.sort(weights)[n // 2])
    cond = weights <= thr
    idx = np.arange(int(cond.sum())) % (k - 1) + 1
    rng.shuffle(idx)
    assign[cond] = idx
    return assign


def kmeans(k, X, weights, n_iter=1000, init_assign=None, rng=None, return_info=False,
           verbose=True):
    """Prior-weighted k-means, restating batch_spalign_kmeans.py:136-183.

    ``init_assign`` (optional) replaces the seeded init so the CUDA path and the oracle
    start identically.  status: 0 = assignments stopped changing, 1 = stopped on an empty
    cluster, 2 = iteration cap.
    """
    X = np.asarray(X)
    weights = np.asarray(weights)
    weights_other = 1 - weights
    if init_assign is None:
        assign = kmeans_init(k, weights, rng)
    else:
        assign = np.asarray(init_assign)
    with np.errstate(invalid='ignore', divide='ignore'):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            centers = np.stack([X[assign == i].mean(axis=0) for i in np.arange(k)])
        status, iters = 2, 0
        for _ in range(n_iter):
            iters += 1
            distances = np.linalg.norm(X[:, None, :] - centers[None, :, :], axis=2)
            new_assign = np.argmin(distances, axis=1).astype(np.int32)
            if np.all(new_assign == assign):
                status = 0
                break
            assign = new_assign
            mask = assign == 0
            centers[0] = weighted_average(X[mask], weights[mask], axis=0)
            for j in range(1, k):
                mask = assign == j
                centers[j] = weighted_average(X[mask], weights_other[mask], axis=0)
            done = False
            for j in range(k):
                if (assign == j).sum() == 0:
                    if verbose:
                        print(('Terminate KMeans iteration due to {}-th cluster is '
                               'empty').format(j))
                    done = True
                    break
            if done:
                status = 1
                break
    if return_info:
        return assign, dict(iters=iters, status=status, centers=centers)
    return assign


# --------------------------------------------------------------------------------------
# a6. paint-back
# --------------------------------------------------------------------------------------


def weighted_kmeans_paint(superpixels, result, n_superpixels_per_image):
    """Paint cluster ids back to pixels (batch_spalign_kmeans.py:191-207), as a gather.

    The reference uses the enumerate index as the label (labels assumed 0..n_i-1) and
    leaves pixels whose label is >= n_i at 0 (zeros_like)."""
    superpixels = np.asarray(superpixels)
    out = np.zeros_like(superpixels)
    i = 0
    for img_idx, n_sp in enumerate(n_superpixels_per_image):
        table = np.asarray(result[i:i + n_sp]).astype(superpixels.dtype)
        lab = superpixels[img_idx]
        ok = (lab >= 0) & (lab < n_sp)
        out[img_idx][ok] = table[lab[ok]]
        i += n_sp
    return out, out == 0


# --------------------------------------------------------------------------------------
# a7. overlap refine (superpixel_overlaps.py:359-369)
# --------------------------------------------------------------------------------------


def refine_overlaps_csr(indptr, indices, counts, road_cell, thr):
    """keep[s] = road_px > 0 and (sum_c M[s,c]*road[c]) / road_px > thr, with road_px the
    number of road pixels of the nearest-upsampled cell mask (= sum_s overlap[s])."""
    indptr = np.asarray(indptr, dtype=np.int64)
    road = np.asarray(road_cell).ravel().astype(np.int64)
    contrib = np.asarray(counts, dtype=np.int64) * road[np.asarray(indices, dtype=np.int64)]
    csum = np.concatenate([[0], np.cumsum(contrib)])
    overlap = csum[indptr[1:]] - csum[indptr[:-1]]
    road_px = int(overlap.sum())
    keep = np.zeros(len(overlap), dtype=bool)
    if road_px > 0:
        keep = (overlap.astype(np.float64) / float(road_px)) > thr
    return overlap, road_px, keep


def refine_overlaps_masks(superpixel, road_mask, thr):
    """Literal restatement of the reference loop on pixel masks (small inputs only)."""
    superpixel = np.asarray(superpixel)
    road_mask = np.asarray(road_mask).astype(np.uint8)
    refined = np.zeros_like(road_mask)
    n_road = float(np.sum(road_mask))
    for idx in np.unique(superpixel):
        sp_mask = superpixel == idx
        overlap = float(np.sum(np.asarray(sp_mask, dtype=np.int32) * road_mask))
        if n_road > 0 and (overlap / float(n_road)) > thr:
            refined[sp_mask] = 1
    return refined


def upsample_nearest(mask_cells, H, W):
    fh, fw = mask_cells.shape
    return np.asarray(mask_cells)[cell_of_pixel(H, fh)][:, cell_of_pixel(W, fw)]


# --------------------------------------------------------------------------------------
# a8. direct feature matrix (direct_clustering.py:297-303)
# --------------------------------------------------------------------------------------


def direct_features(feature_maps):
    """[n, C, h, w] -> float64 [n*h*w, C+2]; last two columns are (x, y) CELL indices."""
    feature_maps = np.asarray(feature_maps)
    n, c, h, w = feature_maps.shape
    xy = np.stack(np.meshgrid(np.arange(w), np.arange(h))).reshape(2, -1)[None].repeat(n, axis=0)
    xy = xy.transpose(0, 2, 1).reshape(-1, 2).astype(np.int32)
    fm = feature_maps.transpose(0, 2, 3, 1).reshape(n * h * w, c)
    return np.concatenate([fm, xy], axis=1)


# --------------------------------------------------------------------------------------
# evaluation (chainercv semantics restated; batch_spalign_kmeans.py:398-405)
# --------------------------------------------------------------------------------------


def confusion(pred, gt, n_class=2):
    pred = np.asarray(pred).ravel().astype(np.int64)
    gt = np.asarray(gt).ravel().astype(np.int64)
    m = gt >= 0
    return np.bincount(n_class * gt[m] + pred[m], minlength=n_class ** 2).reshape(n_class, n_class)


def road_iou(pred_road, gt):
    """gt: -1 void, 1 road, 0 other.  Returns (iou_road, precision, recall, TP, FP, FN)."""
    conf = confusion(pred_road, gt, 2)
    tp, fp, fn = conf[1, 1], conf[0, 1], conf[1, 0]
    with np.errstate(invalid='ignore', divide='ignore'):
        iou = tp / float(tp + fp + fn)
        prec = tp / float(tp + fp)
        rec = tp / float(tp + fn)
    return iou, prec, rec, int(tp), int(fp), int(fn)


# --------------------------------------------------------------------------------------
# whole path on CPU (cpu_baseline "port")
# --------------------------------------------------------------------------------------


def spalign_image_cpu(label, feat_cellmajor, fh, fw, k=4, prior=(0.75, 0.5, 0.1, 0.1),
                      append_pos=True, init_assign=None, rng=None):
    """One image through the count-matrix path on the CPU: overlap CSR -> pooling ->
    prior -> weighted k-means -> paint.  Returns dict of all intermediates."""
    label = np.asarray(label)
    n_sp = int(label.max()) + 1
    indptr, indices, counts = overlap_csr(label, fh, fw, n_sp)
    area, sy, sx = superpixel_stats(label, n_sp)
    feat = pool_count(indptr, indices, counts, feat_cellmajor, area, sy, sx, append_pos)
    w = create_prior(label, *prior)
    assign, info = kmeans(k, feat, w, init_assign=init_assign, rng=rng, return_info=True,
                          verbose=False)
    cmap, road = weighted_kmeans_paint(label[None], assign, [n_sp])
    return dict(indptr=indptr, indices=indices, counts=counts, area=area, sum_y=sy, sum_x=sx,
                features=feat, weights=w, assign=assign, info=info, cluster_map=cmap[0],
                road_mask=road[0])


# --------------------------------------------------------------------------------------
# f2. bilinear-weight overlap matrix (dense pooling of notebooks/Superpixel_Align.ipynb cell 4)
# --------------------------------------------------------------------------------------
# The notebook resizes the feature map to the image size with chainer.functions.resize_images
# and takes the mean over each superpixel.  chainer (v3/v4, not in the tree: parity unpinned)
# samples with corner alignment: u = linspace(0, n_in - 1, n_out), u0 = clip(floor(u), 0,
# n_in - 2), u1 = u0 + 1, weights (u1 - u) and (u - u0).


def bilinear_axis(n_out: int, n_in: int):
    """(i0 [n_out] int, w0 [n_out], w1 [n_out]): output index o reads w0*in[i0] + w1*in[i0+1]."""
    assert n_in >= 2, 'bilinear resize needs at least 2 input samples per axis'
    u = np.linspace(0, n_in - 1, num=n_out)
    i0 = np.clip(np.floor(u).astype(np.int64), 0, n_in - 2)
    return i0, (i0 + 1) - u, u - i0


def resize_bilinear(feature_map, H, W):
    """[C, fh, fw] -> float64 [C, H, W], chainer F.resize_images semantics."""
    F = np.asarray(feature_map, dtype=np.float64)
    C, fh, fw = F.shape
    v0, wv0, wv1 = bilinear_axis(H, fh)
    u0, wu0, wu1 = bilinear_axis(W, fw)
    rows = wv0[None, :, None] * F[:, v0, :] + wv1[None, :, None] * F[:, v0 + 1, :]   # [C, H, fw]
    return wu0[None, None, :] * rows[:, :, u0] + wu1[None, None, :] * rows[:, :, u0 + 1]


def pool_dense_bilinear(label, feature_map):
    """Mean of the bilinearly resized map over each superpixel (notebook cell 4). Small inputs."""
    label = np.asarray(label)
    H, W = label.shape
    up = resize_bilinear(feature_map, H, W)
    S = int(label.max()) + 1
    return np.stack([up[:, label == s].mean(axis=1) for s in range(S)])


def overlap_bilinear_csr(label, fh, fw, n_sp=None):
    """CSR of Wb[s, c] = sum over pixels p of superpixel s of the bilinear weight of cell c at p.
    Every pixel contributes its four neighbours (zero weights included), columns ascending.
    Returns (indptr int32, indices int32, wvals float64); row sums equal the superpixel areas."""
    label = np.asarray(label)
    H, W = label.shape
    if n_sp is None:
        n_sp = int(label.max()) + 1
    v0, wv0, wv1 = bilinear_axis(H, fh)
    u0, wu0, wu1 = bilinear_axis(W, fw)
    nc = fh * fw
    lab = label.astype(np.int64)
    keys, vals = [], []
    for dv, wv in ((0, wv0), (1, wv1)):
        for du, wu in ((0, wu0), (1, wu1)):
            cell = (v0[:, None] + dv) * fw + (u0[None, :] + du)
            keys.append((lab * nc + cell).ravel())
            vals.append((wv[:, None] * wu[None, :]).ravel())
    keys = np.concatenate(keys)
    vals = np.concatenate(vals)
    uk, inv = np.unique(keys, return_inverse=True)
    wsum = np.bincount(inv, weights=vals, minlength=len(uk))
    rows = uk // nc
    indptr = np.zeros(n_sp + 1, dtype=np.int64)
    np.add.at(indptr, rows + 1, 1)
    return np.cumsum(indptr).astype(np.int32), (uk % nc).astype(np.int32), wsum


def pool_weighted(indptr, indices, wvals, feat_cellmajor, area):
    """feat[s] = sum_c W[s,c] * F[c] / area[s] (float64)."""
    F = np.asarray(feat_cellmajor, dtype=np.float64)
    S = len(indptr) - 1
    out = np.zeros((S, F.shape[1]))
    for s in range(S):
        a, b = indptr[s], indptr[s + 1]
        out[s] = np.asarray(wvals[a:b]) @ F[indices[a:b]]
    return out / np.asarray(area, dtype=np.float64)[:, None]


# --------------------------------------------------------------------------------------
# f3. SLIC superpixels (batch_superpixel, batch_spalign_kmeans.py:299-313:
#     skimage.segmentation.slic(img.transpose(1, 2, 0), n_segments)).
# --------------------------------------------------------------------------------------
# scikit-image 0.13.1 is not installed and not in the tree: PARITY UNPINNED.  This restates the
# published algorithm of that version (segmentation/slic_superpixels.py + _slic.pyx: regular
# seed grid, k-means in (y, x, L, a, b) restricted to 2*step windows, ties to the lower segment
# id, max_iter = 10, then 4-connectivity enforcement with min_size = 0.5 * H * W / n_segments) with
# two choices that make a GPU implementation bit-reproducible against it:
#   * colours are quantised to multiples of 2^-12 after the Lab conversion and the division by
#     the compactness, centre sums are integers (exact, order independent);
#   * a component smaller than min_size joins the segment of the pixel LEFT of its first
#     (raster-order) pixel, else the one ABOVE it (skimage takes the last labelled neighbour its
#     flood fill met; both are "an adjacent segment visited earlier").  Final ids are numbered in
#     raster order of each segment's first pixel, as skimage numbers them.
SLIC_Q = 4096.0


def rgb2lab(rgb):
    """skimage.color.rgb2lab (D65, 2 degree observer) on float input as given (no rescale: the
    reference hands 0..255 floats to slic, so the conversion runs on 0..255 -- kept)."""
    arr = np.asarray(rgb, dtype=np.float64)
    mask = arr > 0.04045
    lin = np.where(mask, np.power((arr + 0.055) / 1.055, 2.4), arr / 12.92)
    m = np.array([[0.412453, 0.357580, 0.180423],
                  [0.212671, 0.715160, 0.072169],
                  [0.019334, 0.119193, 0.950227]])
    xyz = lin @ m.T
    xyz = xyz / np.array([0.95047, 1.0, 1.08883])
    mask = xyz > 0.008856
    f = np.where(mask, np.cbrt(xyz), 7.787 * xyz + 16.0 / 116.0)
    L = 116.0 * f[..., 1] - 16.0
    a = 500.0 * (f[..., 0] - f[..., 1])
    b = 200.0 * (f[..., 1] - f[..., 2])
    return np.stack([L, a, b], axis=-1)


def slic_grid(H, W, n_segments):
    """skimage.util.regular_grid for a (1, H, W) volume: (start_y, step_y, start_x, step_x)."""
    step = (H * W / float(n_segments)) ** 0.5
    if H < step:                          # a dimension shorter than the step keeps all of it
        sy, sx = float(H), (H * W / float(n_segments)) / H
    elif W < step:
        sy, sx = (H * W / float(n_segments)) / W, float(W)
    else:
        sy = sx = step
    y0, x0 = int(sy // 2), int(sx // 2)
    return y0, max(1, int(round(sy))), x0, max(1, int(round(sx)))


def slic_quantise(img_chw, compactness=10.0, convert2lab=True):
    """[3, H, W] float image -> int32 [H, W, 3] colours in units of 2^-12 (Lab / compactness)."""
    hwc = np.asarray(img_chw, dtype=np.float32).transpose(1, 2, 0).astype(np.float64)
    col = rgb2lab(hwc) if convert2lab else hwc
    return np.rint(col * (1.0 / compactness) * SLIC_Q).astype(np.int64)


def slic(img_chw, n_segments=100, compactness=10.0, max_iter=10, convert2lab=True,
         enforce_connectivity=True, min_size_factor=0.5, return_raw=False):
    """Label map int32 [H, W] with contiguous ids 0..S-1 (every id present)."""
    q = slic_quantise(img_chw, compactness, convert2lab)
    H, W, _ = q.shape
    y0, sy, x0, sx = slic_grid(H, W, n_segments)
    gy, gx = np.arange(y0, H, sy), np.arange(x0, W, sx)
    ny, nx = len(gy), len(gx)
    step = float(max(sy, sx))
    cy = np.repeat(gy, nx).astype(np.float64)
    cx = np.tile(gx, ny).astype(np.float64)
    cc = q[np.repeat(gy, nx), np.tile(gx, ny)].astype(np.float64) / SLIC_Q
    n_seg = ny * nx
    wsp = 1.0 / (step * step)
    col = q.astype(np.float64) / SLIC_Q
    nearest = np.zeros((H, W), dtype=np.int64)
    yy, xx = np.mgrid[0:H, 0:W]
    for _ in range(max_iter):
        dist = np.full((H, W), np.inf)
        new = nearest.copy()
        for k in range(n_seg):           # ascending k, strict '<': the lower id wins ties
            if not np.isfinite(cy[k]):
                continue
            ya = int(max(cy[k] - 2 * sy, 0)); yb = int(min(cy[k] + 2 * sy + 1, H))
            xa = int(max(cx[k] - 2 * sx, 0)); xb = int(min(cx[k] + 2 * sx + 1, W))
            dy = cy[k] - yy[ya:yb, xa:xb]
            dx = cx[k] - xx[ya:yb, xa:xb]
            d = (dy * dy + dx * dx) * wsp
            dc = col[ya:yb, xa:xb] - cc[k]
            d = d + ((dc[..., 0] * dc[..., 0] + dc[..., 1] * dc[..., 1]) + dc[..., 2] * dc[..., 2])
            upd = d < dist[ya:yb, xa:xb]
            dist[ya:yb, xa:xb][upd] = d[upd]
            new[ya:yb, xa:xb][upd] = k
        changed = not np.array_equal(new, nearest)
        nearest = new
        if not changed:
            break
        cnt = np.bincount(nearest.ravel(), minlength=n_seg).astype(np.float64)
        sum_y = np.bincount(nearest.ravel(), weights=yy.ravel(), minlength=n_seg)
        sum_x = np.bincount(nearest.ravel(), weights=xx.ravel(), minlength=n_seg)
        has = cnt > 0
        cy = np.where(has, sum_y / np.where(has, cnt, 1), cy)   # an empty segment keeps its centre
        cx = np.where(has, sum_x / np.where(has, cnt, 1), cx)
        for c in range(3):
            sc = np.zeros(n_seg, dtype=np.int64)
            np.add.at(sc, nearest.ravel(), q[..., c].ravel())
            cc[:, c] = np.where(has, (sc.astype(np.float64) / SLIC_Q) / np.where(has, cnt, 1), cc[:, c])
    if return_raw or not enforce_connectivity:
        if return_raw:
            return nearest.astype(np.int32)
        _, inv = np.unique(nearest, return_inverse=True)
        return inv.reshape(H, W).astype(np.int32)
    return slic_enforce_connectivity(nearest, int(min_size_factor * H * W / float(n_segments)))


def slic_enforce_connectivity(nearest, min_size):
    """4-connected components of the cluster map; components below min_size join the segment left
    of (else above) their first pixel; ids renumbered in raster order of first pixels."""
    from scipy import ndimage
    H, W = nearest.shape
    comp = np.zeros((H, W), dtype=np.int64)
    n_comp = 0
    for v in np.unique(nearest):
        lab, n = ndimage.label(nearest == v)
        comp[lab > 0] = lab[lab > 0] + n_comp
        n_comp += n
    # root = first (minimum) linear index of each component
    lin = np.arange(H * W).reshape(H, W)
    root = ndimage.minimum(lin, comp, index=np.arange(1, n_comp + 1)).astype(np.int64)
    size = np.bincount(comp.ravel(), minlength=n_comp + 1)[1:]
    root_map = root[comp - 1]                      # per pixel: root index of its component
    target = {}
    order = np.argsort(root)
    for ci in order:                               # ascending first pixel: targets are final
        r = int(root[ci])
        if size[ci] >= min_size:
            continue
        y, x = divmod(r, W)
        if x > 0:
            t = int(root_map[y, x - 1])
        elif y > 0:
            t = int(root_map[y - 1, x])
        else:
            continue
        target[r] = target.get(t, t)
    if target:
        keys = np.array(sorted(target))
        vals = np.array([target[k] for k in keys])
        idx = np.searchsorted(keys, root_map)
        idx_c = np.clip(idx, 0, len(keys) - 1)
        hit = keys[idx_c] == root_map
        root_map = np.where(hit, vals[idx_c], root_map)
    _, inv = np.unique(root_map, return_inverse=True)
    return inv.reshape(H, W).astype(np.int32)


# --------------------------------------------------------------------------------------
# f2. anchor-sampled superpixel align (the reference's own pooling, batch_spalign_kmeans.py:234-274)
# --------------------------------------------------------------------------------------
# Deterministic part restated: given the anchor pixels of a superpixel, each anchor is mapped to
# feature coordinates p = pixel * (fh / H) + 0.5 (the HEIGHT ratio for both axes, :215), clipped
# to [0, f - 0.5] (:237-240); the n_neighbor = 4 nearest cell centres (i + 0.5, j + 0.5) are
# found (:244-246); their bounding box gives the four corner cells and the bilinear weights
# (:247-266).  The reference's argsort is unstable: ties between equidistant centres (common,
# because the coordinates are multiples of 1/8) are broken HERE by the lower flat cell index
# (stable sort) -- the contract of the CUDA path.  Which pixels are anchors is random in the
# reference (random.shuffle of the member list, :232); callers pass them in.


def anchor_cells_weights(py, px, fh, fw):
    """One anchor in feature coordinates -> ((c11, c12, c21, c22), (w11, w12, w21, w22)) with the
    division by the box area folded into the weights (:257-266)."""
    yy, xx = np.meshgrid(np.arange(fh), np.arange(fw))          # the reference's (transposed) grid
    flat = (np.stack([yy, xx]).transpose(1, 2, 0) + 0.5).reshape(-1, 2)
    cell = (flat[:, 0] - 0.5).astype(np.int64) * fw + (flat[:, 1] - 0.5).astype(np.int64)
    d2 = ((flat - np.array([py, px])[None, :]) ** 2).sum(axis=1)
    # stable order by (distance, flat cell index)
    order = np.lexsort((cell, d2))[:4]
    nb = flat[order]
    max_y, max_x = nb.max(axis=0)
    min_y, min_x = nb.min(axis=0)
    area = (max_x - min_x) * (max_y - min_y)
    cells = (int(min_y) * fw + int(min_x), int(max_y) * fw + int(min_x),
             int(min_y) * fw + int(max_x), int(max_y) * fw + int(max_x))
    w = ((max_x - px) * (max_y - py) / area, (max_x - px) * (py - min_y) / area,
         (px - min_x) * (max_y - py) / area, (px - min_x) * (py - min_y) / area)
    return cells, w


def anchor_to_feature_coords(y, x, H, fh, fw):
    r = float(fh) / H
    py = min(max(y * r + 0.5, 0.0), fh - 1 + 0.5)
    px = min(max(x * r + 0.5, 0.0), fw - 1 + 0.5)
    return py, px


def pool_anchors(feature_map, H, anchors, n_valid):
    """feature_map [C, fh, fw]; anchors int [S, n, 2] (y, x) pixels, n_valid [S] -> [S, C]
    float64: mean over the anchors of the bilinear sample (:242-274)."""
    C, fh, fw = feature_map.shape
    F = feature_map.reshape(C, -1).astype(np.float64)
    out = np.zeros((anchors.shape[0], C))
    for s in range(anchors.shape[0]):
        acc = np.zeros(C)
        for a in range(int(n_valid[s])):
            py, px = anchor_to_feature_coords(int(anchors[s, a, 0]), int(anchors[s, a, 1]), H, fh, fw)
            cells, w = anchor_cells_weights(py, px, fh, fw)
            for c, wt in zip(cells, w):
                acc += wt * F[:, c]
        out[s] = acc / max(int(n_valid[s]), 1)
    return out


def replay_reference_anchors(label, n_select=10):
    """The anchors the reference's superpixel_align draws for this label map from the CURRENT
    state of Python's ``random`` (consumed exactly as :226-234 consumes it)."""
    import random
    ids = np.sort(np.unique(label))
    anchors = np.full((len(ids), n_select, 2), -1, dtype=np.int64)
    n_valid = np.zeros(len(ids), dtype=np.int64)
    for i, idx in enumerate(ids):
        y, x = np.where(label == idx)
        coords = list(zip(y.tolist(), x.tolist()))
        random.shuffle(coords)
        sel = coords[:n_select]
        anchors[i, :len(sel)] = np.asarray(sel)
        n_valid[i] = len(sel)
    return anchors, n_valid


# ---------------------------------------------------------------------------------------
# Felzenszwalb-Huttenlocher graph segmentation as scikit-image 0.13 runs it
# (skimage/segmentation/_felzenszwalb_cy.pyx; called at batch_spalign_kmeans.py:303-307 with
# img / 255., scale=300, sigma=0.8, min_size=20).  scikit-image is not in the reference tree:
# PARITY UNPINNED; this restatement is the contract of spalign_felzenszwalb.  One choice is
# pinned where skimage leaves it open: edges of equal cost keep their index order (stable sort;
# skimage uses np.argsort's default, whose tie order is unspecified).

def felz_weights(sigma, truncate=4.0):
    """scipy.ndimage 1-D Gaussian (order 0): (radius, weights[0..radius] by distance)."""
    lw = int(truncate * float(sigma) + 0.5)
    sd = float(sigma) * float(sigma)
    w = [1.0]
    tot = 1.0
    for ii in range(1, lw + 1):
        tmp = np.exp(-0.5 * float(ii * ii) / sd)
        w.append(float(tmp))
        tot += 2.0 * float(tmp)
    return lw, np.asarray(w, dtype=np.float64) / tot


def _felz_reflect(i, n):
    while i < 0 or i >= n:
        i = -i - 1 if i < 0 else 2 * n - 1 - i
    return i


def felz_blur(img, sigma):
    """gaussian_filter(img [H, W, C] float64, sigma=[sigma, sigma, 0]), mode 'reflect': axis 0
    then axis 1, each output = x*w0 + sum over distance d = radius..1 of (x[-d] + x[+d])*w[d]
    (the symmetric branch of scipy's correlate1d, plain multiplies and adds)."""
    lw, w = felz_weights(sigma)
    out = np.asarray(img, dtype=np.float64)
    for axis in (0, 1):
        n = out.shape[axis]
        x = np.moveaxis(out, axis, 0)
        idx = lambda off: np.asarray([_felz_reflect(i + off, n) for i in range(n)])
        acc = x * w[0]
        for d in range(lw, 0, -1):
            acc = acc + (x[idx(-d)] + x[idx(d)]) * w[d]
        out = np.moveaxis(acc, 0, axis)
    return out


def felz_edges(H, W):
    """Edge endpoints in skimage's order: right, down, down-right, up-right."""
    seg = np.arange(H * W).reshape(H, W)
    right = np.c_[seg[:, 1:].ravel(), seg[:, :W - 1].ravel()]
    down = np.c_[seg[1:, :].ravel(), seg[:H - 1, :].ravel()]
    dright = np.c_[seg[1:, 1:].ravel(), seg[:H - 1, :W - 1].ravel()]
    uright = np.c_[seg[:H - 1, 1:].ravel(), seg[1:, :W - 1].ravel()]
    return np.vstack([right, down, dright, uright])


def felz_costs(sm):
    H, W = sm.shape[:2]

    def cost(a, b):
        d = a - b
        return np.sqrt((d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2])
    return np.hstack([cost(sm[:, 1:], sm[:, :W - 1]).ravel(), cost(sm[1:], sm[:H - 1]).ravel(),
                      cost(sm[1:, 1:], sm[:H - 1, :W - 1]).ravel(),
                      cost(sm[1:, :W - 1], sm[:H - 1, 1:]).ravel()])


def felzenszwalb(img_chw, scale=300.0, sigma=0.8, min_size=20):
    """[3, H, W] float32 image with values in 0..1 -> int32 [H, W] labels 0..S-1, numbered in the
    order of np.unique over the union-find roots (a root is the smallest pixel index of its set)."""
    img = np.asarray(img_chw, dtype=np.float32).transpose(1, 2, 0).astype(np.float64)
    H, W, _ = img.shape
    scale = float(scale) / 255.0
    sm = felz_blur(img, sigma) if sigma > 0 else img
    costs = felz_costs(sm)
    edges = felz_edges(H, W)
    order = np.argsort(costs, kind='stable')
    parent = np.arange(H * W)
    size = np.ones(H * W, dtype=np.int64)
    cint = np.zeros(H * W)

    def find(i):
        while parent[i] != i:
            parent[i] = parent[parent[i]]
            i = parent[i]
        return i
    for e in order:
        a, b = find(edges[e, 0]), find(edges[e, 1])
        if a == b:
            continue
        c = costs[e]
        if c < min(cint[a] + scale / size[a], cint[b] + scale / size[b]):
            r, o = (a, b) if a < b else (b, a)
            parent[o] = r
            size[r] = size[a] + size[b]
            cint[r] = c
    for e in order:
        a, b = find(edges[e, 0]), find(edges[e, 1])
        if a == b:
            continue
        if size[a] < min_size or size[b] < min_size:
            r, o = (a, b) if a < b else (b, a)
            parent[o] = r
            size[r] = size[a] + size[b]
    roots = np.asarray([find(i) for i in range(H * W)])
    return np.unique(roots, return_inverse=True)[1].reshape(H, W).astype(np.int32)
