"""Thin torch-tensor wrappers over the C ABI (include/spalign.h).

torch is the carrier only (device memory, streams): every function here enqueues hand-written
sm_100a kernels from libspalign_b200.so on the current CUDA stream and returns device tensors.
Nothing in this module synchronises unless its docstring says so.
"""
from __future__ import annotations

import os

import math
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import check

LAUNCHES = 0  # kernels launched by this library (bench.py reports it)

# kernel launches per entry point (kept in sync with csrc/*.cu)
_LAUNCH_COST = dict(label_max=1, overlap_csr=7, overlap_bilinear_csr=7, pool_weighted=1, pool=1, nchw_to_cellmajor=1, kmeans_groups=1,
                    kmeans_sweep=1, kmeans_finish=1, kmeans_reduce=1, kmeans_update=1, kmeans_init=1, paint=1,
                    refine=2, confusion2=1, slic=0, sample_anchors=1, anchor_weights=1,
                    resize_nearest=1)


def _count(name):
    global LAUNCHES
    LAUNCHES += _LAUNCH_COST[name]


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


def _label_code(t):
    if t.dtype == torch.int32:
        return _lib.I32
    if t.dtype == torch.int64:
        return _lib.I64
    raise TypeError('label maps must be int32 or int64, got %s' % t.dtype)


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.SpalignError('spalign ops need CUDA tensors (there is no CPU fallback)')


# --------------------------------------------------------------------------------------
def label_max(labels: torch.Tensor) -> torch.Tensor:
    """Per-image maximum label, int32 [n_img] (n_superpixels = max + 1)."""
    _require_cuda(labels)
    labels = labels.contiguous()
    n, H, W = labels.shape
    out = torch.empty(n, dtype=torch.int32, device=labels.device)
    check(_lib.load().spalign_label_max(_ptr(labels), _label_code(labels), n, H, W, _ptr(out),
                                        _stream()), 'label_max')
    _count('label_max')
    return out


def prior_axes(H, W, y_rel_pos, x_rel_pos, y_rel_sigma, x_rel_sigma):
    """Separable factors of the Gaussian road prior (batch_spalign_kmeans.py:111-122):
    w(y, x) = gy[y] * gx[x].  Host float64 (H + W exps)."""
    ymean, xmean = int(H * y_rel_pos), int(W * x_rel_pos)
    ys, xs = H * y_rel_sigma, W * x_rel_sigma
    gy = np.exp(-((np.arange(H) - ymean) ** 2 / (2 * ys) ** 2))
    gx = np.exp(-((np.arange(W) - xmean) ** 2 / (2 * xs) ** 2))
    return gy, gx


_CONST_CACHE = {}


def _cached_const(key, make):
    """Small per-device cache of read-only device constants (prior factors, row offsets) so
    repeated batches of the same shape do not pay a pageable host->device copy each."""
    t = _CONST_CACHE.get(key)
    if t is None:
        if len(_CONST_CACHE) > 64:
            _CONST_CACHE.clear()
        t = _CONST_CACHE[key] = make()
    return t


@dataclass
class Overlap:
    """CSR overlap matrix of a batch + per-superpixel statistics (device tensors)."""
    indptr: torch.Tensor      # int32 [n_rows+1]
    indices: torch.Tensor     # int32 [nnz_cap] (first nnz valid), cell id within the image
    counts: torch.Tensor      # int32 [nnz_cap]
    area: torch.Tensor        # int32 [n_rows]
    sum_y: torch.Tensor       # int64 [n_rows]
    sum_x: torch.Tensor       # int64 [n_rows]
    sum_prior: Optional[torch.Tensor]  # float64 [n_rows]
    nnz_flags: torch.Tensor   # int64 [4]
    sp_off: torch.Tensor      # int64 [n_img+1] device
    sp_off_host: np.ndarray   # int64 [n_img+1]
    n_img: int
    H: int
    W: int
    fh: int
    fw: int

    @property
    def n_rows(self) -> int:
        return int(self.sp_off_host[-1])

    @property
    def max_rows(self) -> int:
        return int(np.diff(self.sp_off_host).max())

    def validate(self):
        """Synchronises.  Raises on label-range / capacity problems; returns nnz."""
        nnz, flags, hw, _ = self.nnz_flags.tolist()
        if flags & _lib.F_NNZ_OVERFLOW:
            raise OverflowError('overlap CSR capacity exceeded (nnz=%d, per-image high water %d)'
                                % (nnz, hw))
        if flags & _lib.F_LABEL_RANGE:
            raise ValueError('label map holds ids outside [0, n_superpixels)')
        return nnz

    @property
    def has_empty_rows(self) -> bool:
        return bool(self.nnz_flags[1].item() & _lib.F_EMPTY_ROW)

    def weights(self) -> torch.Tensor:
        """Mean prior per superpixel, float64 (create_prior, batch_spalign_kmeans.py:124-127)."""
        return self.sum_prior / self.area.to(torch.float64)


def overlap_csr(labels: torch.Tensor, fh: int, fw: int, n_sp: Sequence[int],
                prior: Optional[Sequence[float]] = None, nnz_cap_per_image: Optional[int] = None,
                retry: bool = False) -> Overlap:
    """K1.  labels [n_img, H, W] int32/int64 CUDA; n_sp = superpixels per image (host ints).

    ``prior`` = (y_rel_pos, x_rel_pos, y_rel_sigma, x_rel_sigma) also accumulates the prior.
    With ``retry`` the call synchronises and re-runs once with a larger capacity on overflow.
    """
    _require_cuda(labels)
    labels = labels.contiguous()
    n, H, W = labels.shape
    dev = labels.device
    n_sp = np.asarray(n_sp, dtype=np.int64).reshape(-1)
    assert len(n_sp) == n, 'one superpixel count per image'
    sp_off_host = np.concatenate([[0], np.cumsum(n_sp)]).astype(np.int64)
    n_rows = int(sp_off_host[-1])
    sp_off = _cached_const(('sp_off', dev, sp_off_host.tobytes()),
                           lambda: torch.from_numpy(sp_off_host).to(dev))
    ncell = fh * fw
    if nnz_cap_per_image is None:
        nnz_cap_per_image = min(H * W, 3 * ncell + 2 * int(n_sp.max()) + 1024)
    cap = int(nnz_cap_per_image) * n
    gy = gx = None
    if prior is not None:
        def make():
            gy_h, gx_h = prior_axes(H, W, *prior)
            return torch.from_numpy(gy_h).to(dev), torch.from_numpy(gx_h).to(dev)
        gy, gx = _cached_const(('prior', dev, H, W, tuple(float(p) for p in prior)), make)
    lib = _lib.load()
    ws_bytes = lib.spalign_overlap_workspace_bytes(n, H, W, fh, fw, n_rows, cap)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    i32 = dict(dtype=torch.int32, device=dev)
    i64 = dict(dtype=torch.int64, device=dev)
    ov = Overlap(
        indptr=torch.empty(n_rows + 1, **i32), indices=torch.empty(cap, **i32),
        counts=torch.empty(cap, **i32), area=torch.empty(n_rows, **i32),
        sum_y=torch.empty(n_rows, **i64), sum_x=torch.empty(n_rows, **i64),
        sum_prior=torch.empty(n_rows, dtype=torch.float64, device=dev) if prior is not None else None,
        nnz_flags=torch.empty(4, **i64), sp_off=sp_off, sp_off_host=sp_off_host, n_img=n, H=H,
        W=W, fh=fh, fw=fw)
    check(lib.spalign_overlap_csr(
        _ptr(labels), _label_code(labels), n, H, W, fh, fw, _ptr(sp_off), n_rows, _ptr(gy),
        _ptr(gx), cap, _ptr(ov.indptr), _ptr(ov.indices), _ptr(ov.counts), _ptr(ov.area),
        _ptr(ov.sum_y), _ptr(ov.sum_x), _ptr(ov.sum_prior), _ptr(ov.nnz_flags), _ptr(ws),
        ws_bytes, _stream()), 'overlap_csr')
    _count('overlap_csr')
    ov._keepalive = (ws, gy, gx, labels)  # stream-ordered reuse is safe, but keep it simple
    if retry:
        nnz, flags, hw, _ = ov.nnz_flags.tolist()
        if flags & _lib.F_NNZ_OVERFLOW:
            bigger = min(H * W, max(2 * nnz_cap_per_image, int(hw) + 1024, int(nnz) // n + 1024))
            if bigger <= nnz_cap_per_image:
                raise OverflowError('overlap CSR capacity exceeded at the maximum capacity')
            return overlap_csr(labels, fh, fw, n_sp, prior, bigger, retry=True)
    return ov


def bilinear_axis(n_out: int, n_in: int):
    """Corner-aligned bilinear sampling tables of chainer.functions.resize_images along one
    axis: (i0 int32 [n_out], w0, w1 float64 [n_out], start int32 [n_in+1]) with
    out[o] = w0[o]*in[i0[o]] + w1[o]*in[i0[o]+1] and start[c] = first o with i0[o] >= c."""
    assert n_in >= 2, 'bilinear pooling needs at least 2 feature cells per axis'
    u = np.linspace(0, n_in - 1, num=n_out)
    i0 = np.clip(np.floor(u).astype(np.int64), 0, n_in - 2)
    start = np.searchsorted(i0, np.arange(n_in + 1), side='left').astype(np.int32)
    return i0.astype(np.int32), (i0 + 1) - u, u - i0, start


@dataclass
class BilinearOverlap:
    """CSR bilinear-weight overlap matrix of a batch (device tensors)."""
    indptr: torch.Tensor       # int32 [n_rows+1]
    indices: torch.Tensor      # int32 [nnz_cap]
    wvals: torch.Tensor        # float64 [nnz_cap]
    row_weight: torch.Tensor   # float64 [n_rows] (= area up to rounding)
    nnz_flags: torch.Tensor    # int64 [4]

    def validate(self):
        nnz, flags, hw, _ = self.nnz_flags.tolist()
        if flags & _lib.F_NNZ_OVERFLOW:
            raise OverflowError('bilinear overlap CSR capacity exceeded (nnz=%d)' % nnz)
        if flags & _lib.F_LABEL_RANGE:
            raise ValueError('label map holds ids outside [0, n_superpixels)')
        return nnz


def overlap_bilinear_csr(labels: torch.Tensor, fh: int, fw: int, ov: Overlap,
                         nnz_cap_per_image: Optional[int] = None) -> BilinearOverlap:
    """K1b.  Bilinear-weight overlap matrix for the same batch / row numbering as ``ov``."""
    _require_cuda(labels)
    labels = labels.contiguous()
    n, H, W = labels.shape
    dev = labels.device
    ncell = fh * fw
    if nnz_cap_per_image is None:
        nnz_cap_per_image = min(4 * H * W, 6 * ncell + 8 * ov.max_rows + 1024)
    cap = int(nnz_cap_per_image) * n

    def make():
        iy0, wy0, wy1, ys = bilinear_axis(H, fh)
        ix0, wx0, wx1, xs = bilinear_axis(W, fw)
        return tuple(torch.from_numpy(np.ascontiguousarray(a)).to(dev)
                     for a in (iy0, wy0, wy1, ys, ix0, wx0, wx1, xs))
    tabs = _cached_const(('bilinear', dev, H, W, fh, fw), make)
    lib = _lib.load()
    ws_bytes = lib.spalign_overlap_bilinear_workspace_bytes(n, H, W, fh, fw, ov.n_rows, cap)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    out = BilinearOverlap(
        indptr=torch.empty(ov.n_rows + 1, dtype=torch.int32, device=dev),
        indices=torch.empty(cap, dtype=torch.int32, device=dev),
        wvals=torch.empty(cap, dtype=torch.float64, device=dev),
        row_weight=torch.empty(ov.n_rows, dtype=torch.float64, device=dev),
        nnz_flags=torch.empty(4, dtype=torch.int64, device=dev))
    check(lib.spalign_overlap_bilinear_csr(
        _ptr(labels), _label_code(labels), n, H, W, fh, fw, _ptr(ov.sp_off), ov.n_rows,
        *[_ptr(t) for t in tabs], cap, _ptr(out.indptr), _ptr(out.indices), _ptr(out.wvals),
        _ptr(out.row_weight), _ptr(out.nnz_flags), _ptr(ws), ws_bytes, _stream()),
        'overlap_bilinear_csr')
    _count('overlap_bilinear_csr')
    return out


def pool_weighted(feat_cellmajor: torch.Tensor, ov: Overlap, bw: BilinearOverlap,
                  append_pos: bool = True) -> torch.Tensor:
    """K2 with the bilinear weights: [n_rows, C(+2)] float32; the centroid columns and the
    normalisation (area) come from the count matrix ``ov``."""
    _require_cuda(feat_cellmajor)
    n, ncell, C = feat_cellmajor.shape
    assert n == ov.n_img and ncell == ov.fh * ov.fw and feat_cellmajor.is_contiguous()
    D = C + (2 if append_pos else 0)
    ld = padded_ld(D)
    out = torch.empty((ov.n_rows, ld), dtype=torch.float32, device=feat_cellmajor.device)
    check(_lib.load().spalign_pool_weighted(
        _ptr(feat_cellmajor), n, C, ov.fh, ov.fw, _ptr(ov.sp_off), ov.n_rows, ov.max_rows,
        _ptr(bw.indptr), _ptr(bw.indices), _ptr(bw.wvals), _ptr(ov.area), _ptr(ov.sum_y),
        _ptr(ov.sum_x), int(append_pos), _ptr(out), ld, _stream()), 'pool_weighted')
    _count('pool_weighted')
    return out[:, :D]


def sample_anchors(labels: torch.Tensor, ov: Overlap, n_select: int = 10, seed: int = 1111):
    """f2: n_select distinct member pixels per superpixel, uniform over the members (the
    reference shuffles with Python's ``random``, batch_spalign_kmeans.py:232; this stream is a
    statistical equivalent).  Returns (anchors int32 [n_rows, n_select, 2], n_valid int32)."""
    _require_cuda(labels)
    labels = labels.contiguous()
    n, H, W = labels.shape
    dev = labels.device
    anchors = torch.empty((ov.n_rows, n_select, 2), dtype=torch.int32, device=dev)
    n_valid = torch.empty(ov.n_rows, dtype=torch.int32, device=dev)
    check(_lib.load().spalign_sample_anchors(
        _ptr(labels), _label_code(labels), n, H, W, ov.fh, ov.fw, _ptr(ov.sp_off), ov.n_rows,
        _ptr(ov.indptr), _ptr(ov.indices), _ptr(ov.counts), _ptr(ov.area), int(n_select),
        int(seed) & 0xffffffffffffffff, _ptr(anchors), _ptr(n_valid), _stream()), 'sample_anchors')
    _count('sample_anchors')
    return anchors, n_valid


def pool_anchors(feat_cellmajor: torch.Tensor, ov: Overlap, anchors: torch.Tensor,
                 n_valid: torch.Tensor, H: int, append_pos: bool = True) -> torch.Tensor:
    """f2: the reference's anchor-sampled descriptors (batch_spalign_kmeans.py:234-274) for
    given anchors: [n_rows, C(+2)] float32 (centroid columns from the count matrix ``ov``)."""
    _require_cuda(feat_cellmajor, anchors, n_valid)
    n, ncell, C = feat_cellmajor.shape
    assert n == ov.n_img and ncell == ov.fh * ov.fw and feat_cellmajor.is_contiguous()
    dev = feat_cellmajor.device
    anchors = anchors.to(torch.int32).contiguous()
    n_valid = n_valid.to(torch.int32).contiguous()
    R, n_select = anchors.shape[0], anchors.shape[1]
    assert R == ov.n_rows
    indptr = torch.empty(R + 1, dtype=torch.int32, device=dev)
    indices = torch.empty(R * 4 * n_select, dtype=torch.int32, device=dev)
    wvals = torch.empty(R * 4 * n_select, dtype=torch.float64, device=dev)
    lib = _lib.load()
    check(lib.spalign_anchor_weights(_ptr(anchors), _ptr(n_valid), R, n_select, int(H), ov.fh, ov.fw,
                                     _ptr(indptr), _ptr(indices), _ptr(wvals), _stream()),
          'anchor_weights')
    _count('anchor_weights')
    D = C + (2 if append_pos else 0)
    ld = padded_ld(D)
    out = torch.empty((R, ld), dtype=torch.float32, device=dev)
    ones = torch.ones(R, dtype=torch.int32, device=dev)     # the weights carry the 1 / n_valid
    check(lib.spalign_pool_weighted(
        _ptr(feat_cellmajor), n, C, ov.fh, ov.fw, _ptr(ov.sp_off), R, ov.max_rows, _ptr(indptr),
        _ptr(indices), _ptr(wvals), _ptr(ones), _ptr(ov.sum_y), _ptr(ov.sum_x), 0, _ptr(out), ld,
        _stream()), 'pool_weighted')
    _count('pool_weighted')
    if append_pos:
        area = ov.area.to(torch.float64)
        out[:, C] = (ov.sum_y.to(torch.float64) / area).float()
        out[:, C + 1] = (ov.sum_x.to(torch.float64) / area).float()
    return out[:, :D]


# --------------------------------------------------------------------------------------
def as_cellmajor(feature_maps: torch.Tensor) -> torch.Tensor:
    """[n, C, fh, fw] (any memory format) -> [n, fh*fw, C] contiguous float32.

    channels_last input is a zero-copy view; NCHW input goes through the transpose kernel."""
    _require_cuda(feature_maps)
    assert feature_maps.dim() == 4
    n, C, fh, fw = feature_maps.shape
    if feature_maps.dtype != torch.float32:
        feature_maps = feature_maps.float()
    nhwc = feature_maps.permute(0, 2, 3, 1)
    if nhwc.is_contiguous():
        return nhwc.reshape(n, fh * fw, C)
    src = feature_maps.contiguous()
    dst = torch.empty((n, fh * fw, C), dtype=torch.float32, device=src.device)
    check(_lib.load().spalign_nchw_to_cellmajor(_ptr(src), _ptr(dst), n, C, fh * fw, _stream()),
          'nchw_to_cellmajor')
    _count('nchw_to_cellmajor')
    return dst


def padded_ld(d: int) -> int:
    """Row stride (floats) used for descriptor matrices: multiple of 4 floats (16 bytes)."""
    return (d + 3) // 4 * 4


def pool(feat_cellmajor: torch.Tensor, ov: Overlap, append_pos: bool = True,
         out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """K2.  feat [n_img, fh*fw, C] float32 cell-major -> descriptors [n_rows, C(+2)] float32
    (a view of a [n_rows, ld] buffer with ld = padded_ld(C+2); pad columns are zero)."""
    _require_cuda(feat_cellmajor)
    n, ncell, C = feat_cellmajor.shape
    assert n == ov.n_img and ncell == ov.fh * ov.fw and feat_cellmajor.is_contiguous()
    D = C + (2 if append_pos else 0)
    ld = padded_ld(D)
    if out is None:
        out = torch.empty((ov.n_rows, ld), dtype=torch.float32, device=feat_cellmajor.device)
    assert out.shape == (ov.n_rows, ld) and out.is_contiguous()
    check(_lib.load().spalign_pool(
        _ptr(feat_cellmajor), n, C, ov.fh, ov.fw, _ptr(ov.sp_off), ov.n_rows, ov.max_rows,
        _ptr(ov.indptr), _ptr(ov.indices), _ptr(ov.counts), _ptr(ov.area), _ptr(ov.sum_y),
        _ptr(ov.sum_x), int(append_pos), _ptr(out), ld, _stream()), 'pool')
    _count('pool')
    return out[:, :D]


# --------------------------------------------------------------------------------------
def _x_code(X):
    if X.dtype == torch.float32:
        return _lib.F32, 4
    if X.dtype == torch.float64:
        return _lib.F64, 8
    raise TypeError('k-means rows must be float32 or float64')


def as_kmeans_rows(X: torch.Tensor) -> torch.Tensor:
    """Return a [N, Dr] view whose row stride is a multiple of 16 bytes and base 16-byte
    aligned (copying into a padded buffer when needed)."""
    _require_cuda(X)
    assert X.dim() == 2
    code, es = _x_code(X)
    per16 = 16 // es
    ok = (X.stride(1) == 1 and X.stride(0) % per16 == 0 and X.data_ptr() % 16 == 0
          and X.stride(0) >= (X.shape[1] + per16 - 1) // per16 * per16)
    if ok:
        return X
    ld = (X.shape[1] + per16 - 1) // per16 * per16
    buf = torch.zeros((X.shape[0], ld), dtype=X.dtype, device=X.device)
    buf[:, :X.shape[1]] = X
    return buf[:, :X.shape[1]]


@dataclass
class KMeansResult:
    assign: torch.Tensor   # int32 [N]
    iters: torch.Tensor    # int32 [G]
    status: torch.Tensor   # int32 [G]
    centers: Optional[torch.Tensor]  # float64 [G, K, D]


def kmeans_groups(X: torch.Tensor, w: torch.Tensor, init_assign: torch.Tensor, K: int,
                  group_off: torch.Tensor, n_iter: int = 1000, pos_grid=None,
                  want_centers: bool = False) -> KMeansResult:
    """K3, one persistent CTA per group: the whole iteration loop runs on the device.

    X [N, Dr] float32/float64 (row stride multiple of 16 B), w [N] float64, init_assign [N]
    int32 (copied), group_off [G+1] int64 device.  ``pos_grid=(fh, fw)`` appends the virtual
    (x, y) cell-index columns of direct_clustering.py:297-303."""
    _require_cuda(X, w, init_assign, group_off)
    X = as_kmeans_rows(X)
    code, _ = _x_code(X)
    N, Dr = X.shape
    D = Dr + (2 if pos_grid else 0)
    G = group_off.numel() - 1
    dev = X.device
    assign = init_assign.to(torch.int32).clone().contiguous()
    w = w.to(torch.float64).contiguous()
    iters = torch.empty(G, dtype=torch.int32, device=dev)
    status = torch.empty(G, dtype=torch.int32, device=dev)
    centers = torch.empty((G, K, D), dtype=torch.float64, device=dev) if want_centers else None
    pos_w = pos_grid[1] if pos_grid else 0
    pos_period = pos_grid[0] * pos_grid[1] if pos_grid else 0
    check(_lib.load().spalign_kmeans_groups(
        _ptr(X), code, X.stride(0), 1 if pos_grid else 0, pos_w, pos_period, _ptr(w), D, K,
        n_iter, _ptr(group_off), G, _ptr(assign), _ptr(centers), _ptr(iters), _ptr(status), None,
        0, _stream()), 'kmeans_groups')
    _count('kmeans_groups')
    return KMeansResult(assign, iters, status, centers)


class KMeansLarge:
    """K3 with many CTAs per problem: sweeps over row chunks driven from the host, one step per
    iteration, for any number of independent groups at once.

    ``step()`` enqueues sweep -> reduce -> [all-reduce hook] -> update (3 launches).  ``run()``
    loops, reads the stop flags every ``poll`` iterations (one small D2H) and re-cuts the chunk
    list over the groups that are still running, so late iterations of a few slow groups still
    spread over the whole GPU.  Used for batches of per-image problems, joint batches, direct
    cell clustering and (with ``allreduce``) the multi-GPU global clustering.
    """

    TILE = 16           # rows per shared-memory tile of the main kernel variant
    TARGET_CHUNKS = 4 * 148
    TAIL_ROWS = 2048    # groups up to this size finish in a persistent CTA each

    def __init__(self, X, w, init_assign, K, group_off_host, n_iter=1000, pos_grid=None,
                 pos_row0=0, chunks_per_group=None, allreduce=None, fused=True,
                 incremental=True, bounds=True, tail=True, tail_after=1, tail_slice=0,
                 comm=None):
        _require_cuda(X, w, init_assign)
        self.X = as_kmeans_rows(X)
        self.code, _ = _x_code(self.X)
        self.N, Dr = self.X.shape
        self.K, self.n_iter = K, n_iter
        self.D = Dr + (2 if pos_grid else 0)
        self.pos_mode = 1 if pos_grid else 0
        self.pos_w = pos_grid[1] if pos_grid else 0
        self.pos_period = pos_grid[0] * pos_grid[1] if pos_grid else 0
        self.pos_row0 = pos_row0
        self.allreduce = allreduce
        self.comm = comm      # dist_kmeans.PeerComm: the exchange runs inside the iterate kernel
        if comm is not None:
            assert allreduce is None and fused and len(group_off_host) == 2, \
                'peer-memory exchange: one global group, fused iterate kernel'
        self.chunks_per_group = chunks_per_group
        self.dev = dev = self.X.device
        self.w = w.to(torch.float64).contiguous()
        self.assign = init_assign.to(torch.int32).clone().contiguous()
        self.goff = np.asarray(group_off_host, dtype=np.int64)
        self.G = len(self.goff) - 1
        self.pv = K * (self.D + 2) + 1
        self._set_chunks(np.arange(self.G))
        self.partials = torch.zeros((self.n_chunks, self.pv), dtype=torch.float64, device=dev)
        self.totals = torch.zeros((self.G, self.pv), dtype=torch.float64, device=dev)
        self.centers = torch.zeros((self.G, K, self.D), dtype=torch.float64, device=dev)
        self.iters = torch.zeros(self.G, dtype=torch.int32, device=dev)
        self.status = torch.full((self.G,), _lib.KM_RUNNING, dtype=torch.int32, device=dev)
        self.counters = torch.zeros(self.G + self.n_chunks, dtype=torch.int32, device=dev)
        self.ub = self.lb = self.cdelta = None
        if bounds and incremental and fused and self.code == _lib.F32:
            self.ub = torch.empty(self.N, dtype=torch.float32, device=dev)
            self.lb = torch.empty(self.N, dtype=torch.float32, device=dev)
            self.cdelta = torch.zeros((self.G, K), dtype=torch.float64, device=dev)
        self.fused = fused
        self.incremental = incremental
        self._full_done = False
        # many small groups: after the first full iteration one persistent CTA per group runs
        # the remaining iterations on its own (spalign_kmeans_finish), no per-iteration launch
        sizes = np.diff(self.goff)
        max_rows = int(sizes.max()) if self.G else 0
        self.tail = (tail and self.ub is not None and allreduce is None and comm is None and
                     chunks_per_group is None and
                     (max_rows <= self.TAIL_ROWS or
                      (max_rows <= 8 * self.TAIL_ROWS and self.G >= 148)))
        self.tail_after = tail_after
        self.tail_slice = tail_slice
        if self.tail:
            self.goff_dev = torch.from_numpy(self.goff).to(dev, non_blocking=True)
        self._lib = _lib.load()
        self._init_done = False

    def _set_chunks(self, active):
        """Cut the rows of the ``active`` groups into chunks (vectorised; host side)."""
        active = np.asarray(active, dtype=np.int64)
        r0, r1 = self.goff[active], self.goff[active + 1]
        n = r1 - r0
        if self.chunks_per_group:
            per = np.maximum(1, -(-n // self.chunks_per_group))
            per = -(-per // self.TILE) * self.TILE
            nch = np.maximum(1, -(-n // per))
            rep = np.repeat(np.arange(len(active)), nch)
            j = np.arange(int(nch.sum())) - np.repeat(np.cumsum(nch) - nch, nch)
            rb = r0[rep] + j * per[rep]
            re = np.minimum(rb + per[rep], r1[rep])
            size_rank = np.zeros(len(rb), dtype=np.int64)
        else:
            # uniform chunks; per-CTA fixed cost (centres, screening set-up, partials) argues
            # for large chunks, load balance for small ones: 32..1024 rows by active volume
            # (1024 = the limit of the bounds pass; large problems get large chunks and with
            # them fewer partial-sum vectors for the fused reduction to add up)
            # many small groups (per-image clustering) keep the 256-row cap tuned in round 1
            total = int(n.sum())
            cap = int(os.environ.get('SPALIGN_KM_CHUNK_ROWS', 0)) or \
                (64 * self.TILE if int(n.max()) > 4096 else 16 * self.TILE)
            rpc = max(2 * self.TILE, min(cap, total // self.TARGET_CHUNKS))
            rpc = -(-rpc // self.TILE) * self.TILE
            per = np.full(len(active), rpc, dtype=np.int64)
            nch = np.maximum(1, -(-n // per))
            rep = np.repeat(np.arange(len(active)), nch)
            j = np.arange(int(nch.sum())) - np.repeat(np.cumsum(nch) - nch, nch)
            rb = r0[rep] + j * per[rep]
            re = np.minimum(rb + per[rep], r1[rep])
            size_rank = np.zeros(len(rb), dtype=np.int64)
        tot = int(nch.sum())
        # slot = group-contiguous, row-ordered position (what the fixed-order reduction walks);
        # launch order = large chunks first (stable)
        slot = np.arange(tot, dtype=np.int64)
        order = np.argsort(size_rank, kind='stable')
        chunks = np.stack([active[rep], rb, np.maximum(re, rb), slot], axis=1).astype(np.int64)[order]
        counts = np.zeros(self.G, dtype=np.int64)
        counts[active] = nch
        gco = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
        self.n_chunks = tot
        if getattr(self, 'partials', None) is not None and tot > self.partials.shape[0]:
            self.partials = torch.zeros((tot, self.pv), dtype=torch.float64, device=self.dev)
            self.counters = torch.zeros(self.G + tot, dtype=torch.int32, device=self.dev)
        self.chunks = torch.from_numpy(chunks).to(self.dev, non_blocking=True)
        self.gco = torch.from_numpy(gco).to(self.dev, non_blocking=True)

    def _sweep(self, mode):
        if self.n_chunks == 0:
            return
        if self.allreduce is None and self.fused:
            # one launch per iteration: the last chunk of every group reduces and updates.
            # After the first full iteration the centroid sums are maintained incrementally
            # (mode 2: only rows that changed cluster are subtracted / added).
            if mode == 1 and self.incremental:
                if self._full_done:
                    mode = 2
                self._full_done = True
            args = (_ptr(self.X), self.code, self.X.stride(0), self.pos_mode, self.pos_w,
                    self.pos_period, self.pos_row0, _ptr(self.w), self.D, self.K, _ptr(self.chunks),
                    self.n_chunks, _ptr(self.gco), mode, self.n_iter, _ptr(self.assign),
                    _ptr(self.partials), _ptr(self.totals), _ptr(self.centers), _ptr(self.iters),
                    _ptr(self.status), _ptr(self.counters), _ptr(self.ub),
                    _ptr(self.lb), _ptr(self.cdelta))
            if self.comm is not None:
                check(self._lib.spalign_kmeans_iterate_dist(*args, self.comm.handle, _stream()),
                      'kmeans_iterate_dist')
            else:
                check(self._lib.spalign_kmeans_iterate(*args, _stream()), 'kmeans_iterate')
            _count('kmeans_sweep')
            return
        check(self._lib.spalign_kmeans_sweep(
            _ptr(self.X), self.code, self.X.stride(0), self.pos_mode, self.pos_w,
            self.pos_period, self.pos_row0, _ptr(self.w), self.D, self.K, _ptr(self.chunks),
            self.n_chunks, _ptr(self.centers), mode, _ptr(self.assign), _ptr(self.status),
            _ptr(self.partials), _stream()), 'kmeans_sweep')
        _count('kmeans_sweep')
        check(self._lib.spalign_kmeans_reduce(_ptr(self.partials), _ptr(self.gco), self.G, self.D,
                                              self.K, _ptr(self.totals), _stream()),
              'kmeans_reduce')
        _count('kmeans_reduce')
        if self.allreduce is not None:
            self.allreduce(self.totals)
        check(self._lib.spalign_kmeans_update(_ptr(self.totals), self.G, self.D, self.K, mode,
                                              self.n_iter, _ptr(self.centers), _ptr(self.iters),
                                              _ptr(self.status), _stream()), 'kmeans_update')
        _count('kmeans_update')

    def init_centers(self):
        self._sweep(0)
        self._init_done = True

    def step(self):
        if not self._init_done:
            self.init_centers()
        self._sweep(1)

    def can_split(self) -> bool:
        """True when the iterations after the first run in the per-group finish kernel, i.e. when
        ``prepare()`` + ``finish_groups()`` over any partition of the groups equals ``run()``."""
        return bool(self.tail and self.n_iter > 0)

    def prepare(self):
        """Seed centres (one sweep) and the first full iteration(s): leaves running sums, centres,
        centre drift and Hamerly bounds for ``finish_groups``."""
        if not self._init_done:
            self.init_centers()
        for _ in range(min(self.tail_after, self.n_iter)):
            self._sweep(1)

    def finish_groups(self, g0: int, g1: int):
        """Remaining iterations of groups [g0, g1): one persistent CTA per group on the current
        stream.  Groups are independent, so disjoint ranges may run on different streams (the
        paint-back of a finished range can then start while slower ranges still iterate)."""
        # default: one launch, one CTA per SM (4 rows per warp set).  tail_slice > 0 first runs
        # that many iterations with every group resident (2 CTAs per SM, 2 rows per set), then
        # the rest -- measured no faster at 300 groups (profiles/README.md)
        plan = [(0, 4)]
        if self.tail_slice > 0:
            plan = [(self.tail_slice, 2), (0, 4)]
        for slice_iters, rows in plan:
            check(self._lib.spalign_kmeans_finish(
                _ptr(self.X), self.code, self.X.stride(0), self.pos_mode, self.pos_w,
                self.pos_period, self.pos_row0, _ptr(self.w), self.D, self.K,
                _ptr(self.goff_dev[g0:]), g1 - g0, self.n_iter, _ptr(self.assign),
                _ptr(self.totals[g0:]), _ptr(self.centers[g0:]), _ptr(self.iters[g0:]),
                _ptr(self.status[g0:]), _ptr(self.ub), _ptr(self.lb), _ptr(self.cdelta[g0:]),
                slice_iters, rows, _stream()), 'kmeans_finish')
            _count('kmeans_finish')

    def result(self) -> KMeansResult:
        return KMeansResult(self.assign, self.iters, self.status, self.centers)

    def run(self, poll: int = 4, first_poll: int = 6, blocking: bool = False,
            lookahead: int = 2) -> KMeansResult:
        """Iterate until every group has stopped.  The stop flags are read back asynchronously
        (pinned buffer + event) after ``first_poll`` iterations and then every ``poll``; sweeps
        keep being enqueued meanwhile (at most ``lookahead`` of them; CTAs of finished groups
        exit at once), so the GPU does not idle during the round trip.  When a read-back lands, finished groups are dropped from the chunk
        list.  ``blocking=True`` synchronises at every poll instead (deterministic launch
        count, used by tests)."""
        if self.can_split():
            self.prepare()
            self.finish_groups(0, self.G)
            return self.result()
        if not self._init_done:
            self.init_centers()
        if self.allreduce is not None:
            # every rank must enqueue the same number of collectives: poll at fixed iterations
            blocking = True
        done = 0
        nxt = min(first_poll, self.n_iter)
        pending = None
        ahead_until = 0
        can_recut = self.allreduce is None and self.chunks_per_group is None
        if not blocking and getattr(self, '_pin', None) is None:
            self._pin = torch.empty(self.G, dtype=torch.int32).pin_memory()
        while done < self.n_iter:
            self._sweep(1)
            done += 1
            if blocking:
                if done >= nxt:
                    running = (self.status == _lib.KM_RUNNING).cpu().numpy()   # sync
                    if not running.any():
                        break
                    if can_recut:
                        self._set_chunks(np.nonzero(running)[0])
                    nxt = min(done + poll, self.n_iter)
                continue
            if pending is None and done >= nxt:
                self._pin.copy_(self.status, non_blocking=True)
                pending = torch.cuda.Event()
                pending.record()
                ahead_until = done + lookahead
            if pending is not None and (pending.query() or done >= ahead_until
                                        or done >= self.n_iter):
                pending.synchronize()
                running = self._pin.numpy() == _lib.KM_RUNNING
                pending = None
                if not running.any():
                    break
                if can_recut:
                    self._set_chunks(np.nonzero(running)[0])
                nxt = done + poll
        if self.n_iter == 0:
            self.status.fill_(_lib.KM_ITER_CAP)
        return KMeansResult(self.assign, self.iters, self.status, self.centers)


def kmeans_debug_stats(reset: bool = False):
    """(rows screened in fp32, rows sent to the exact float64 pass) since the last reset.
    Synchronises the device."""
    import ctypes
    buf = (ctypes.c_int64 * 2)()
    check(_lib.load().spalign_kmeans_debug_stats(buf, int(reset)), 'kmeans_debug_stats')
    return int(buf[0]), int(buf[1])


def kmeans_init_device(w: torch.Tensor, group_off: torch.Tensor, shuffled: torch.Tensor,
                       shuf_off: torch.Tensor):
    """Seeded init on the device (any group size).  Returns (assign int32 [N], m int32 [G]);
    m[g] != len(shuffled_g) means prior-weight ties changed the split size: the shuffle drawn
    for the expected size does not match the reference's stream for that group."""
    _require_cuda(w, group_off, shuffled, shuf_off)
    G = group_off.numel() - 1
    assign = torch.empty(w.numel(), dtype=torch.int32, device=w.device)
    m = torch.empty(G, dtype=torch.int32, device=w.device)
    check(_lib.load().spalign_kmeans_init(_ptr(w), _ptr(group_off), G, _ptr(shuffled),
                                          _ptr(shuf_off), _ptr(assign), _ptr(m), _stream()),
          'kmeans_init')
    _count('kmeans_init')
    return assign, m


# --------------------------------------------------------------------------------------
def felzenszwalb(images: torch.Tensor, scale: float = 300.0, sigma: float = 0.8,
                 min_size: int = 20, chunk: int = 64):
    """f3: Felzenszwalb-Huttenlocher label maps on the device (the reference's default
    superpixel method).  images [n, 3, H, W] float32 CUDA (values in 0..1) -> (labels int32
    [n, H, W] with ids 0..S-1, n_labels int32 [n]).  Contract and oracle:
    oracle/spalign_oracle.py:felzenszwalb (parity unpinned: scikit-image is not in the reference
    tree).  ``chunk`` images share one workspace (their merge passes run side by side)."""
    global LAUNCHES
    _require_cuda(images)
    assert images.dim() == 4 and images.shape[1] == 3
    images = images.float().contiguous()
    n, _, H, W = images.shape
    dev = images.device
    lib = _lib.load()
    labels = torch.empty((n, H, W), dtype=torch.int32, device=dev)
    n_labels = torch.empty(n, dtype=torch.int32, device=dev)
    chunk = max(1, min(chunk, n))
    ws_bytes = lib.spalign_felzenszwalb_workspace_bytes(chunk, H, W)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    for i in range(0, n, chunk):
        m = min(chunk, n - i)
        check(lib.spalign_felzenszwalb(_ptr(images[i:]), m, H, W, float(scale), float(sigma),
                                       int(min_size), _ptr(labels[i:]), _ptr(n_labels[i:]),
                                       _ptr(ws), ws_bytes, _stream()), 'felzenszwalb')
        LAUNCHES += 6 + 3 * m   # blur x2, costs, merge, relabel; radix sort passes per image
    return labels, n_labels


# --------------------------------------------------------------------------------------
def slic(images: torch.Tensor, n_segments: int = 100, compactness: float = 10.0, max_iter: int = 10,
         convert2lab: bool = True, enforce_connectivity: bool = True, min_size_factor: float = 0.5,
         chunk: int = 16):
    """f3: SLIC label maps on the device.  images [n, 3, H, W] float32 CUDA (values in 0..1) ->
    (labels int32 [n, H, W] with contiguous ids, n_labels int32 [n]).  Contract and oracle:
    oracle/spalign_oracle.py:slic (parity unpinned: scikit-image is not in the reference tree)."""
    global LAUNCHES
    _require_cuda(images)
    assert images.dim() == 4 and images.shape[1] == 3
    images = images.float().contiguous()
    n, _, H, W = images.shape
    dev = images.device
    lib = _lib.load()
    labels = torch.empty((n, H, W), dtype=torch.int32, device=dev)
    n_labels = torch.empty(n, dtype=torch.int32, device=dev)
    ws_bytes = lib.spalign_slic_workspace_bytes(min(chunk, n), H, W, int(n_segments))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    for i in range(0, n, chunk):
        m = min(chunk, n - i)
        check(lib.spalign_slic(_ptr(images[i:i + m]), m, H, W, int(n_segments), float(compactness),
                               int(max_iter), int(convert2lab), int(enforce_connectivity),
                               float(min_size_factor), _ptr(labels[i:i + m]), _ptr(n_labels[i:i + m]),
                               _ptr(ws), ws_bytes, _stream()), 'slic')
        LAUNCHES += 3 + 5 * int(max_iter) + (30 if enforce_connectivity else 3) + 3
    return labels, n_labels


_OUT_CODES = {torch.uint8: _lib.U8, torch.int32: _lib.I32, torch.int64: _lib.I64}


def paint(labels: torch.Tensor, sp_off: torch.Tensor, table: torch.Tensor,
          out_dtype: Optional[torch.dtype] = torch.uint8, want_mask: bool = True,
          road_value: int = 0, out=None):
    """K4: cluster_map[p] = table[sp_off[img] + label[p]], road_mask = (cluster_map == road_value).
    Returns (cluster_map or None, road_mask uint8 or None).  ``out`` = (cluster_map, road_mask)
    contiguous [n, H, W] tensors to write into (e.g. slices of a larger batch)."""
    _require_cuda(labels, sp_off, table)
    labels = labels.contiguous()
    n, H, W = labels.shape
    dev = labels.device
    table = table.to(torch.int32).contiguous()
    if out is not None:
        cmap, mask = out
        for o in (cmap, mask):
            if o is not None and (not o.is_contiguous() or tuple(o.shape) != (n, H, W)):
                raise ValueError('paint: out tensors must be contiguous [n, H, W]')
        if mask is not None and mask.dtype != torch.uint8:
            raise ValueError('paint: road mask must be uint8')
        out_dtype = cmap.dtype if cmap is not None else None
    else:
        cmap = torch.empty((n, H, W), dtype=out_dtype, device=dev) if out_dtype is not None else None
        mask = torch.empty((n, H, W), dtype=torch.uint8, device=dev) if want_mask else None
    check(_lib.load().spalign_paint(
        _ptr(labels), _label_code(labels), n, H, W, _ptr(sp_off), _ptr(table), _ptr(cmap),
        _OUT_CODES[out_dtype] if out_dtype is not None else _lib.U8, _ptr(mask), road_value,
        _stream()), 'paint')
    _count('paint')
    return cmap, mask


def refine(ov: Overlap, road_cell: torch.Tensor, thr: float):
    """K5: per-superpixel overlap with a cell-level road mask [n_img, fh, fw] (bool/uint8).
    Returns (overlap int64 [n_rows], road_px int64 [n_img], keep int32 [n_rows])."""
    _require_cuda(road_cell)
    dev = road_cell.device
    rc = road_cell.to(torch.uint8).contiguous().reshape(ov.n_img, -1)
    overlap = torch.empty(ov.n_rows, dtype=torch.int64, device=dev)
    road_px = torch.empty(ov.n_img, dtype=torch.int64, device=dev)
    keep = torch.empty(ov.n_rows, dtype=torch.int32, device=dev)
    check(_lib.load().spalign_refine(
        _ptr(ov.sp_off), ov.n_img, ov.n_rows, ov.fh * ov.fw, ov.max_rows, _ptr(ov.indptr),
        _ptr(ov.indices), _ptr(ov.counts), _ptr(rc), float(thr), _ptr(overlap), _ptr(road_px),
        _ptr(keep), _stream()), 'refine')
    _count('refine')
    return overlap, road_px, keep


def resize_nearest_u8(maps: torch.Tensor, H: int, W: int) -> torch.Tensor:
    """uint8 / bool maps [n, h, w] -> [n, H, W] with cv2.INTER_NEAREST index arithmetic (the
    resize to the label shape at batch_spalign_kmeans.py:470-477)."""
    _require_cuda(maps)
    m = maps.to(torch.uint8).contiguous()
    n, h, w = m.shape
    if (h, w) == (H, W):
        return m
    out = torch.empty((n, H, W), dtype=torch.uint8, device=m.device)
    check(_lib.load().spalign_resize_nearest_u8(_ptr(m), n, h, w, _ptr(out), int(H), int(W),
                                                _stream()), 'resize_nearest_u8')
    _count('resize_nearest')
    return out


def confusion2(pred: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
    """2-class confusion per image: conf[img, gt, pred] over pixels with gt >= 0
    (chainercv semantics used at batch_spalign_kmeans.py:398-402).  int64 [n_img, 2, 2]."""
    _require_cuda(pred, gt)
    n = pred.shape[0]
    pred = pred.to(torch.uint8).contiguous().reshape(n, -1)
    gt = gt.to(torch.int32).contiguous().reshape(n, -1)
    conf = torch.empty((n, 2, 2), dtype=torch.int64, device=pred.device)
    check(_lib.load().spalign_confusion2(_ptr(pred), _ptr(gt), n, pred.shape[1], _ptr(conf),
                                         _stream()), 'confusion2')
    _count('confusion2')
    return conf
