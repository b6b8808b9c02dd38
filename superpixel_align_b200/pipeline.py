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

    def check(self):
        """Synchronises.  Raises on an overflowed / out-of-range overlap matrix (whose rows were
        then left EMPTY by K1, so nothing downstream read unwritten storage) and on label ids
        with gaps; returns the number of groups whose prior weights were tied at the median, for
        which the seeded init differs from the reference's ``np.random`` stream."""
        self.overlap.validate()
        if self.overlap.has_empty_rows:
            raise ValueError('label ids are not contiguous 0..S-1 (rows without pixels): use '
                             'batch_spalign_kmeans.prepare_batch, which relabels like the reference')
        ties = int((self.init_m.cpu().numpy() != np.asarray(self.shuf_sizes)).sum())
        if ties:
            import warnings
            warnings.warn('%d group(s) with prior weights tied at the median: their seeded init '
                          'does not follow the reference stream' % ties)
        return ties


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
              timers: Optional[dict] = None, kmeans_impl: str = 'chunks',
              out=None, paint_overlap: int = 0) -> PipelineOutput:
    """One pass of the hot path over a batch.  ``images_per_group`` = the reference's
    ``--batchsize`` (superpixels of that many consecutive images are clustered jointly;
    1 = per-image clustering).  Groups of more than 2048 rows use the multi-CTA k-means whose
    driver polls a stop flag asynchronously; everything else is free of host synchronisation.
    Data-dependent conditions (overlap capacity, label range, empty rows, tied prior weights)
    are left in device words: call ``.check()`` on the result at the next synchronisation.
    ``paint_overlap`` = p > 1 cuts the groups into p ranges whose k-means finish kernels run on
    separate streams; the paint-back of a range starts when that range has stopped, under the
    tail of the slower ranges (same results; the 'kmeans' timer then includes the paint-back)."""
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
    flat, off, m_exp = draw_shuffles(k, sizes)
    goff = ov.sp_off if images_per_group == 1 else \
        torch.from_numpy(group_off_host).to(dev, non_blocking=True)
    flat_d = torch.from_numpy(flat).to(dev, non_blocking=True)
    off_d = torch.from_numpy(off).to(dev, non_blocking=True)
    mark('init', True)
    init, m = ops.kmeans_init_device(weights, goff, flat_d, off_d)
    mark('init', False)
    mark('kmeans', True)
    km = None
    if kmeans_impl == 'groups':
        res = ops.kmeans_groups(feats, weights, init, k, goff, n_iter=n_iter)
    else:  # many CTAs per group; small groups finish in one persistent CTA each
        km = ops.KMeansLarge(feats, weights, init, k, group_off_host, n_iter=n_iter)
    n_groups = len(group_off_host) - 1
    if km is not None and paint_overlap > 1 and km.can_split() and n_groups >= 2 * paint_overlap:
        # The finish kernel ends with its slowest group while most SMs are already idle.  Ranges
        # of groups finish on separate (high-priority) streams; the paint-back of a range starts
        # on a low-priority stream as soon as that range is done and runs on the SMs the
        # remaining groups do not occupy.  Same kernels, same results.
        km.prepare()
        cmap, mask = out if out is not None else (
            torch.empty(labels.shape, dtype=out_dtype, device=dev),
            torch.empty(labels.shape, dtype=torch.uint8, device=dev))
        cur = torch.cuda.current_stream()
        fork = torch.cuda.Event()
        fork.record(cur)
        hi, lo = _paint_streams(dev, paint_overlap)
        cuts = np.linspace(0, n_groups, paint_overlap + 1).astype(np.int64)
        for j in range(paint_overlap):
            g0, g1 = int(cuts[j]), int(cuts[j + 1])
            i0, i1 = int(g_idx[g0]), int(g_idx[g1])
            hi[j].wait_event(fork)
            with torch.cuda.stream(hi[j]):
                km.finish_groups(g0, g1)
                done = torch.cuda.Event()
                done.record(hi[j])
            lo[j].wait_event(done)
            with torch.cuda.stream(lo[j]):
                ops.paint(labels[i0:i1], ov.sp_off[i0:i1 + 1], km.assign,
                          out=(cmap[i0:i1], mask[i0:i1]))
                painted = torch.cuda.Event()
                painted.record(lo[j])
            cur.wait_event(painted)
        res = km.result()
        mark('kmeans', False)   # includes the overlapped paint-back
    else:
        if km is not None:
            res = km.run()
        mark('kmeans', False)
        mark('paint', True)
        cmap, mask = ops.paint(labels, ov.sp_off, res.assign, out_dtype=out_dtype, out=out)
        mark('paint', False)
    return PipelineOutput(cmap, mask, res.assign, feats, weights, res.iters, res.status, m, ov,
                          group_off_host, m_exp)


_SIDE_STREAMS = {}
_PAINT_STREAMS = {}


def _paint_streams(dev, n):
    """(high-priority streams for the k-means finish ranges, default-priority ones for their
    paint-back), cached per device."""
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), n)
    if key not in _PAINT_STREAMS:
        _PAINT_STREAMS[key] = ([torch.cuda.Stream(device=dev, priority=-1) for _ in range(n)],
                               [torch.cuda.Stream(device=dev) for _ in range(n)])
    return _PAINT_STREAMS[key]


def _side_streams(dev, n):
    key = (dev.index if dev.index is not None else torch.cuda.current_device(), n)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = [torch.cuda.Stream(device=dev) for _ in range(n)]
    return _SIDE_STREAMS[key]


@dataclass
class OverlappedOutput:
    cluster_map: torch.Tensor   # uint8 [n, H, W]
    road_mask: torch.Tensor     # uint8 [n, H, W]
    parts: list                 # PipelineOutput per sub-batch, in image order

    @property
    def iters(self):
        return torch.cat([p.iters for p in self.parts])

    @property
    def status(self):
        return torch.cat([p.status for p in self.parts])

    @property
    def assign(self):
        return torch.cat([p.assign for p in self.parts])


def run_batch_overlapped(labels: torch.Tensor, feat_cellmajor: torch.Tensor,
                         n_sp: Sequence[int], fh: int, fw: int, sub_batch: int = 60,
                         n_streams: int = 2, out_dtype=torch.uint8, timers: Optional[list] = None,
                         **kw) -> OverlappedOutput:
    """``run_batch`` over sub-batches of ``sub_batch`` images that alternate between
    ``n_streams`` CUDA streams.  The stages of one sub-batch have very different appetites --
    K2/K4 stream HBM at full bandwidth, the k-means tail is a latency-bound handful of CTAs
    waiting for its slowest images -- so running two sub-batches side by side lets the pooling
    of one fill the SMs and the memory system that the k-means tail of the other leaves idle.
    Same results as one ``run_batch`` call (the seeded init consumes ``np.random`` in image
    order either way); the outputs land in one [n, H, W] pair.  The calling stream waits for
    all side streams before this returns (no host synchronisation)."""
    n, H, W = labels.shape
    dev = labels.device
    n_sp = np.asarray(n_sp, dtype=np.int64)
    cmap = torch.empty((n, H, W), dtype=out_dtype, device=dev)
    mask = torch.empty((n, H, W), dtype=torch.uint8, device=dev)
    main = torch.cuda.current_stream()
    streams = _side_streams(dev, n_streams)
    start = torch.cuda.Event()
    start.record(main)
    parts = []
    for j, i0 in enumerate(range(0, n, sub_batch)):
        i1 = min(n, i0 + sub_batch)
        st = streams[j % n_streams]
        with torch.cuda.stream(st):
            if j < n_streams:
                st.wait_event(start)
            tm = {} if timers is not None else None
            parts.append(run_batch(labels[i0:i1], feat_cellmajor[i0:i1], n_sp[i0:i1], fh, fw,
                                   timers=tm, out=(cmap[i0:i1], mask[i0:i1]), **kw))
            if timers is not None:
                timers.append(tm)
    for st in streams:
        ev = torch.cuda.Event()
        ev.record(st)
        main.wait_event(ev)
    for t in (cmap, mask):
        for st in streams:
            t.record_stream(st)
    return OverlappedOutput(cmap, mask, parts)


def _raise_on_overlap_flags(words, index):
    """The K1 condition words of sub-batch ``index`` (copied back with its result): the same
    errors ``Overlap.validate()`` / ``PipelineOutput.check()`` raise.  On a capacity overflow K1
    left an EMPTY matrix, so the maps handed back would be meaningless, not unsafe."""
    nnz, flags, hw, _ = [int(v) for v in words.tolist()]
    if flags & _lib.F_NNZ_OVERFLOW:
        raise OverflowError('sub-batch %d: overlap CSR capacity exceeded (nnz=%d, per-image high '
                            'water %d); pass a larger nnz_cap_per_image' % (index, nnz, hw))
    if flags & _lib.F_LABEL_RANGE:
        raise ValueError('sub-batch %d: label outside [0, n_sp)' % index)
    if flags & _lib.F_EMPTY_ROW:
        raise ValueError('sub-batch %d: label ids are not contiguous 0..S-1 (rows without pixels)'
                         % index)


class HostPipeline:
    """End-to-end path for inputs that live in HOST memory (pinned): label maps and cell-major
    feature maps go host -> device on a copy stream while the previous sub-batch is being
    processed on the compute stream, and the resulting cluster maps / road masks come back
    device -> host on the copy stream.  Two device slots and two pinned result slots are
    reused, so nothing is allocated in steady state.

        hp = HostPipeline(H, W, fh, fw, C, sub_batch=6)
        hp.process([(labels_cpu, feats_cpu, n_sp), ...], on_result=lambda i, cmap, mask: ...)

    ``labels_cpu`` [b, H, W] int32 and ``feats_cpu`` [b, fh*fw, C] float32 are pinned CPU
    tensors with b <= sub_batch; ``on_result`` receives pinned uint8 CPU tensors that are only
    valid during the callback.
    """

    def __init__(self, H, W, fh, fw, C, sub_batch=6, k=4, prior=(0.75, 0.5, 0.1, 0.1),
                 append_pos=True, device=None, label_dtype=torch.int32):
        self.dev = torch.device('cuda', torch.cuda.current_device()) if device is None else device
        self.shape = (H, W, fh, fw, C)
        self.k, self.prior, self.append_pos, self.B = k, prior, append_pos, sub_batch
        d = self.dev
        self.lab = [torch.empty((sub_batch, H, W), dtype=label_dtype, device=d) for _ in range(2)]
        self.feat = [torch.empty((sub_batch, fh * fw, C), dtype=torch.float32, device=d)
                     for _ in range(2)]
        self.out_c = [torch.empty((sub_batch, H, W), dtype=torch.uint8).pin_memory() for _ in range(2)]
        self.out_m = [torch.empty((sub_batch, H, W), dtype=torch.uint8).pin_memory() for _ in range(2)]
        self.flags = [torch.zeros(4, dtype=torch.int64).pin_memory() for _ in range(2)]
        self.copy = torch.cuda.Stream(device=d)
        self.comp = torch.cuda.Stream(device=d)
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def _upload(self, slot, labels_cpu, feats_cpu, free_event):
        b = labels_cpu.shape[0]
        with torch.cuda.stream(self.copy):
            if free_event is not None:
                self.copy.wait_event(free_event)      # slot's previous occupant is done
            self.lab[slot][:b].copy_(labels_cpu, non_blocking=True)
            self.feat[slot][:b].copy_(feats_cpu, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy)
        self.h2d_bytes += labels_cpu.numel() * labels_cpu.element_size() + \
            feats_cpu.numel() * feats_cpu.element_size()
        return ev

    def process(self, batches, on_result=None):
        H, W, fh, fw, C = self.shape
        batches = list(batches)
        n = len(batches)
        if n == 0:
            return
        comp_done = [None, None]     # compute finished reading device slot s
        d2h_done = [None, None]      # pinned result slot s has been consumed / is free
        up = self._upload(0, batches[0][0], batches[0][1], None)
        pending = None               # (index, slot, b, event) of the D2H in flight
        for i in range(n):
            s = i & 1
            labels_cpu, feats_cpu, n_sp = batches[i]
            b = labels_cpu.shape[0]
            nxt = None
            if i + 1 < n:                # prefetch the next sub-batch while this one computes
                nxt = self._upload(s ^ 1, batches[i + 1][0], batches[i + 1][1], comp_done[s ^ 1])
            with torch.cuda.stream(self.comp):
                self.comp.wait_event(up)
                out = run_batch(self.lab[s][:b], self.feat[s][:b], n_sp, fh, fw, k=self.k,
                                prior=self.prior, append_pos=self.append_pos)
                ev_c = torch.cuda.Event()
                ev_c.record(self.comp)
            comp_done[s] = ev_c
            if pending is not None:      # hand the previous result to the caller
                pi, ps, pb, pev = pending
                pev.synchronize()
                _raise_on_overlap_flags(self.flags[ps], pi)
                if on_result is not None:
                    on_result(pi, self.out_c[ps][:pb], self.out_m[ps][:pb])
            with torch.cuda.stream(self.copy):
                self.copy.wait_event(ev_c)
                self.out_c[s][:b].copy_(out.cluster_map, non_blocking=True)
                self.out_m[s][:b].copy_(out.road_mask, non_blocking=True)
                self.flags[s].copy_(out.overlap.nnz_flags, non_blocking=True)
                ev_d = torch.cuda.Event()
                ev_d.record(self.copy)
            out.cluster_map.record_stream(self.copy)
            out.road_mask.record_stream(self.copy)
            self.d2h_bytes += 2 * b * H * W
            pending = (i, s, b, ev_d)
            up = nxt
        pi, ps, pb, pev = pending
        pev.synchronize()
        _raise_on_overlap_flags(self.flags[ps], pi)
        if on_result is not None:
            on_result(pi, self.out_c[ps][:pb], self.out_m[ps][:pb])


def bind_to_gpu_numa_node(device_index: int) -> dict:
    """Pin this process (and with it the first-touch placement of the pinned pools it allocates
    next) to the CPUs NVML reports as local to the GPU.  On hosts that expose a single NUMA node
    this changes nothing; it is the software half of keeping eight host->device streams from
    meeting on one memory controller.  Returns what was done, for the bench record."""
    import os
    info = {'device': int(device_index), 'bound': False}
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        cpus = [64 * i + b for i, wd in enumerate(words) for b in range(64) if (wd >> b) & 1]
        cpus = [c for c in cpus if c < n_cpu]
        try:
            info['numa_node'] = int(pynvml.nvmlDeviceGetNumaNodeId(h))
        except Exception:
            pass
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info.update(bound=True, n_cpus=len(allowed), cpu_first=allowed[0], cpu_last=allowed[-1])
    except Exception as e:  # pragma: no cover
        info['error'] = repr(e)[:120]
    return info


class ImagePipeline:
    """The reference's real boundary (estimate_road_mask, batch_spalign_kmeans.py:427-457) for
    inputs in HOST memory: uint8 RGB images [b, 3, H, W] and label maps [b, H, W] (uint16 when
    S <= 65535, else int32) go host -> device, the DRN backbone (PyTorch / cuDNN, the input
    producer) runs on the device, its stride-8 map feeds K1..K4 without ever leaving HBM, and the
    uint8 cluster maps / road masks come back.  Copies run on a second stream, double buffered.

        ip = ImagePipeline(model, H, W, sub_batch=4)
        ip.process([(images_u8_cpu, labels_cpu, n_sp), ...], on_result=lambda i, cmap, mask: ...)
    """

    MEAN = (0.485, 0.456, 0.406)
    STD = (0.229, 0.224, 0.225)

    def __init__(self, model, H, W, sub_batch=4, k=4, prior=(0.75, 0.5, 0.1, 0.1), append_pos=True,
                 device=None, label_dtype=torch.int16, backbone_dtype=torch.float32):
        self.dev = torch.device('cuda', torch.cuda.current_device()) if device is None else device
        self.model, self.H, self.W, self.B = model, H, W, sub_batch
        self.k, self.prior, self.append_pos = k, prior, append_pos
        self.backbone_dtype = backbone_dtype
        d = self.dev
        self.img = [torch.empty((sub_batch, 3, H, W), dtype=torch.uint8, device=d) for _ in range(2)]
        self.lab = [torch.empty((sub_batch, H, W), dtype=label_dtype, device=d) for _ in range(2)]
        self.out_c = [torch.empty((sub_batch, H, W), dtype=torch.uint8).pin_memory() for _ in range(2)]
        self.out_m = [torch.empty((sub_batch, H, W), dtype=torch.uint8).pin_memory() for _ in range(2)]
        self.flags = [torch.zeros(4, dtype=torch.int64).pin_memory() for _ in range(2)]
        self.copy = torch.cuda.Stream(device=d)
        self.comp = torch.cuda.Stream(device=d)
        self.mean = torch.tensor(self.MEAN, device=d).view(1, 3, 1, 1) * 255.0
        self.istd = 1.0 / (torch.tensor(self.STD, device=d).view(1, 3, 1, 1) * 255.0)
        self.h2d_bytes = self.d2h_bytes = 0
        self.backbone_ms = []        # (event, event) pairs around the backbone, read by the bench

    def _upload(self, slot, imgs_cpu, labels_cpu, free_event):
        b = imgs_cpu.shape[0]
        with torch.cuda.stream(self.copy):
            if free_event is not None:
                self.copy.wait_event(free_event)
            self.img[slot][:b].copy_(imgs_cpu, non_blocking=True)
            self.lab[slot][:b].copy_(labels_cpu, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy)
        self.h2d_bytes += imgs_cpu.numel() * imgs_cpu.element_size() + \
            labels_cpu.numel() * labels_cpu.element_size()
        return ev

    def features(self, imgs_u8):
        """uint8 [b,3,H,W] on the device -> cell-major float32 [b, fh*fw, C] (layer8 of the DRN,
        normalised like models/drn.py:304-325)."""
        with torch.no_grad():
            x = (imgs_u8.float() - self.mean) * self.istd
            x = x.contiguous(memory_format=torch.channels_last)
            if self.backbone_dtype != torch.float32:
                with torch.autocast('cuda', dtype=self.backbone_dtype):
                    f = self.model(x)
            else:
                f = self.model(x)
            f = f.float().contiguous(memory_format=torch.channels_last)
        n, C, fh, fw = f.shape
        return f.permute(0, 2, 3, 1).reshape(n, fh * fw, C), fh, fw

    def process(self, batches, on_result=None, time_backbone=False):
        batches = list(batches)
        n = len(batches)
        if n == 0:
            return
        H, W = self.H, self.W
        comp_done = [None, None]
        up = self._upload(0, batches[0][0], batches[0][1], None)
        pending = None
        for i in range(n):
            s = i & 1
            imgs_cpu, labels_cpu, n_sp = batches[i]
            b = imgs_cpu.shape[0]
            nxt = None
            if i + 1 < n:
                nxt = self._upload(s ^ 1, batches[i + 1][0], batches[i + 1][1], comp_done[s ^ 1])
            with torch.cuda.stream(self.comp):
                self.comp.wait_event(up)
                if time_backbone:
                    e0 = torch.cuda.Event(enable_timing=True)
                    e0.record(self.comp)
                feat, fh, fw = self.features(self.img[s][:b])
                if time_backbone:
                    e1 = torch.cuda.Event(enable_timing=True)
                    e1.record(self.comp)
                    self.backbone_ms.append((e0, e1, b))
                labels = self.lab[s][:b].to(torch.int32)     # uint16 on the wire, int32 for K1/K4
                if self.lab[s].dtype == torch.int16:
                    labels = labels & 0xffff
                out = run_batch(labels, feat, n_sp, fh, fw, k=self.k, prior=self.prior,
                                append_pos=self.append_pos)
                ev_c = torch.cuda.Event()
                ev_c.record(self.comp)
            comp_done[s] = ev_c
            if pending is not None:
                pi, ps, pb, pev = pending
                pev.synchronize()
                _raise_on_overlap_flags(self.flags[ps], pi)
                if on_result is not None:
                    on_result(pi, self.out_c[ps][:pb], self.out_m[ps][:pb])
            with torch.cuda.stream(self.copy):
                self.copy.wait_event(ev_c)
                self.out_c[s][:b].copy_(out.cluster_map, non_blocking=True)
                self.out_m[s][:b].copy_(out.road_mask, non_blocking=True)
                self.flags[s].copy_(out.overlap.nnz_flags, non_blocking=True)
                ev_d = torch.cuda.Event()
                ev_d.record(self.copy)
            out.cluster_map.record_stream(self.copy)
            out.road_mask.record_stream(self.copy)
            self.d2h_bytes += 2 * b * H * W
            pending = (i, s, b, ev_d)
            up = nxt
        pi, ps, pb, pev = pending
        pev.synchronize()
        _raise_on_overlap_flags(self.flags[ps], pi)
        if on_result is not None:
            on_result(pi, self.out_c[ps][:pb], self.out_m[ps][:pb])
