#!/usr/bin/env python
"""Benchmark of the superpixel-align hot path (BASELINE.json metric: images/sec at 1024x2048).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config 1|3]
                    [--images M] [--verify V] [--quick]

One "step" = one pass of the hot path (K1 overlap CSR -> K2 pooling -> prior -> seeded init ->
K3 per-image k-means -> K4 paint-back) over a batch of synthetic Cityscapes-shaped images that
is already resident in HBM.

  --config 1 (default) BASELINE.json configs[1]: 300 images of 1024x2048 per GPU, SLIC-shaped
             ~1000 superpixels, DRN-C-26 layer8 (stride 8, 512 channels) random-init features.
             Under torchrun every rank owns its range of a 300*N image set (the reference's shell
             rule, utils/create_val_labels.sh:38-52), no collective on the data path: weak scaling.
  --config 3 BASELINE.json configs[2]: 500 images in total, split over the N ranks by the same
             rule (strong scaling; the last shard is the short one).

The line also carries, each in its own key: `e2e` (the reference's real boundary: uint8 images +
label maps in HOST memory -> DRN on the device -> K1..K4 -> masks back in host memory),
`e2e_features_precomputed` (host fp32 features over PCIe, round 1's e2e), `dropin_numpy` (the
reference-named batch_* functions, NumPy in / NumPy out), `s_sweep` (configs[3]: S = 500 / 1000 /
2000 / 4000), `joint30` (the reference's default --batchsize 30 joint clustering), `global_kmeans`
(configs[4]: one clustering over all ranks' descriptors, exchange inside the kernel over NVLink
vs NCCL), `verify` (the whole batch against the CPU oracle) and `cpu_baseline`.

Prints ONE JSON line (rank 0).  `--impl reference` times the CPU restatement of the reference
path (oracle/, kind "port") on the host cores instead.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, FH, FW, C = 1024, 2048, 128, 256, 512
GY, GX = 25, 40            # 1000 superpixels
K = 4
PRIOR = (0.75, 0.5, 0.1, 0.1)
PAINT_OVERLAP = 10         # image ranges whose paint-back runs under the k-means tail
S_GRIDS = {500: (20, 25), 1000: (25, 40), 2000: (40, 50), 4000: (50, 80)}


def shard_range(n_data, n_shards, rank):
    """utils/create_val_labels.sh:38-52: step = n/N + 1; [i, min(i+step, n))."""
    step = n_data // n_shards + 1
    lo = min(rank * step, n_data)
    return lo, min(lo + step, n_data)


# ------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix='.csv')
            os.close(fd)
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                 '--format=csv,noheader,nounits', '-lms', '20'],
                stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def count(self):
        try:
            with open(self.path) as fp:
                return sum(1 for _ in fp)
        except Exception:
            return 0

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(',')]
                if len(f) < 7:
                    continue
                try:
                    sm.append(float(f[0]))
                    mx.append(float(f[1]))
                except ValueError:
                    continue
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
                                    'sw_power_cap'), f[3:7]):
                    if v.lower().startswith('active'):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), samples=len(sm))
        out['reasons'] = sorted(reasons)
        return out


# ------------------------------------------------------------- CPU side: oracle port / reference
_SHARED = {}     # arrays inherited by forked workers (no pickling of 75 MB per image)


def _cpu_one(job):
    """One image through the oracle port of the count-matrix path (NumPy/SciPy, float64)."""
    idx, seed, init = job
    from oracle import spalign_oracle as so
    lab, feat = _SHARED['labels'][idx], _SHARED['feats'][idx]
    rs = np.random.RandomState(seed)
    t0 = time.time()
    r = so.spalign_image_cpu(lab, feat, FH, FW, k=K, prior=PRIOR, append_pos=True, rng=rs,
                             init_assign=init)
    dt = time.time() - t0
    if init is None:
        return dt
    return dt, np.asarray(r['assign']).astype(np.int32), r['features'].astype(np.float32), \
        np.packbits(r['road_mask'])


def _cpu_one_given_descriptors(job):
    """Reference kmeans() semantics (oracle) on the descriptors the GPU produced."""
    idx, X, w, init = job
    from oracle import spalign_oracle as so
    a, info = so.kmeans(K, X.astype(np.float64), w, init_assign=init.astype(np.float64),
                        return_info=True, verbose=False)
    return np.asarray(a).astype(np.int32), int(info['iters']), int(info['status'])


def _ref_one(job):
    """One image through the reference's OWN functions, unmodified (oracle/_ref copies of the
    scripts, AST-extracted): superpixel_align (10 anchors) + create_prior + weighted_kmeans."""
    idx, seed = job
    import contextlib
    import io
    import random
    from oracle import ref_extract
    ref = ref_extract.load('batch_spalign_kmeans.py', seed=seed)
    random.seed(seed)
    lab = _SHARED['labels'][idx].astype(np.int64)
    fmap = np.ascontiguousarray(_SHARED['feats'][idx].T).reshape(C, FH, FW)
    img = np.zeros((3, H, W), dtype=np.float32)
    t0 = time.time()
    with contextlib.redirect_stdout(io.StringIO()):      # the reference prints; stdout is one JSON line
        f = ref.superpixel_align(img, fmap, lab, 10, 4, True)
        w = ref.create_prior(lab, *PRIOR)
        ref.weighted_kmeans(lab[None], f, w, K, [len(np.unique(lab))])
    return time.time() - t0


class CpuPool:
    """Forked worker pool over shared host arrays, one image per task."""

    def __init__(self, labels, feats, procs):
        import multiprocessing as mp
        _SHARED['labels'], _SHARED['feats'] = labels, feats
        self.n, self.procs = len(labels), procs
        self.pool = mp.get_context('fork').Pool(procs)

    def run_port(self, n_images):
        jobs = [(i % self.n, i, None) for i in range(n_images)]
        t0 = time.time()
        per = self.pool.map(_cpu_one, jobs, chunksize=1)
        dt = time.time() - t0
        return n_images / dt, float(np.mean(per)), dt

    def run_reference_functions(self):
        """images/s of the reference's own functions (kind "reference"), one image per worker
        process in parallel; ~40 s per image per core, so one round only."""
        t0 = time.time()
        per = self.pool.map(_ref_one, [(i % self.n, 1111 + i) for i in range(self.procs)],
                            chunksize=1)
        return self.procs / (time.time() - t0), float(np.mean(per))

    def close(self):
        self.pool.close()
        self.pool.join()


def host_sample(n, use_cuda):
    """`n` label maps + cell-major feature maps on the host (same generators as the GPU arm)."""
    from superpixel_align_b200 import synth
    labels = [synth.voronoi_labels(H, W, GY, GX, image_index=i) for i in range(n)]
    feats = None
    if use_cuda:
        try:
            import torch
            from superpixel_align_b200 import drn
            dev = torch.device('cuda', 0)
            model = drn.drn_c_26(device=dev)
            feats = []
            for i in range(n):
                img = synth.smooth_images_torch(1, H, W, first_index=i, device=dev)
                f = synth.drn_features_torch(model, img)           # [1,512,128,256] channels_last
                feats.append(f.permute(0, 2, 3, 1).reshape(FH * FW, C).cpu().numpy())
            del model
            torch.cuda.empty_cache()
        except Exception as e:  # pragma: no cover
            print('bench: DRN features unavailable (%s); using smooth noise' % e, file=sys.stderr)
            feats = None
    if feats is None:
        feats = [synth.smooth_features(C, FH, FW, seed=i).reshape(C, -1).T.copy() for i in range(n)]
    return labels, feats


def reference_functions_available():
    from oracle import ref_extract
    return ref_extract.available()


REF_SAMPLE = ('%d images, one per worker process: the reference\'s own superpixel_align (10 anchors) '
              '+ create_prior + weighted_kmeans run unmodified (oracle/_ref copies of the scripts, '
              'AST-extracted, NumPy float64), %.1f s per image per core')


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 32))
    n_pool = min(4, procs)
    try:
        import torch
        use_cuda = torch.cuda.is_available()
    except Exception:
        use_cuda = False
    labels, feats = host_sample(n_pool, use_cuda)
    per_step = procs  # one image per worker per step: a bounded sample of the batch
    port = CpuPool(labels, feats, procs)
    port.run_port(procs)        # warm the workers (imports, page-in)
    for _ in range(args.warmup):
        port.run_port(per_step)
    dt, tot = 0.0, 0
    for _ in range(args.steps):
        dt += port.run_port(per_step)[2]
        tot += per_step
    ref_line = None
    if reference_functions_available() and not args.quick:
        v, per = port.run_reference_functions()
        ref_line = {'value': v, 'unit': 'images/s', 'cores': procs, 'kind': 'reference',
                    'sample': REF_SAMPLE % (procs, per)}
    port.close()
    value = tot / dt
    world = int(os.environ.get('WORLD_SIZE', '1'))
    line = {
        'impl': 'reference', 'metric': 'images/sec (1024x2048 hot path)', 'value': value,
        'unit': 'images/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1000.0 * dt / max(1, args.steps), 'higher_is_better': True,
        'scaling': 'strong' if args.config == 3 else 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic', 'config': workload_config(args, world),
        'cpu_baseline': {'value': value, 'unit': 'images/s', 'cores': procs, 'kind': 'port',
                         'sample': '%d images per step (one per worker process) of the batch, '
                                   'oracle/spalign_oracle.py count-matrix path, float64' % per_step},
        'reference_functions': ref_line,
        'e2e': {'value': value, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def workload_config(args, n_gpus):
    if args.config == 3:
        return {'workload': 'configs[2] val-shaped set: 500 synthetic 1024x2048 images in total, '
                            'split over the GPUs by the reference rule step = 500/N + 1 '
                            '(utils/create_val_labels.sh:38-52), per-image clustering, no collective',
                'images_total': 500, 'n_gpus': n_gpus, 'superpixels': GY * GX, 'channels': C,
                'shard_sizes': [shard_range(500, n_gpus, r)[1] - shard_range(500, n_gpus, r)[0]
                                for r in range(n_gpus)],
                'l2_policy': 'inputs far larger than the 126 MB L2; no flush needed'}
    n_images = args.images
    return {'workload': 'configs[1] random300-shaped batch: %d synthetic 1024x2048 images per GPU, '
                        'SLIC-shaped 1000 superpixels (jittered Voronoi), DRN-C-26 layer8 512-ch '
                        'stride-8 features (random init), per-image prior-weighted k-means K=4' % n_images,
            'images_per_gpu': n_images, 'n_gpus': n_gpus, 'superpixels': GY * GX, 'channels': C,
            'clustering': 'per-image (batchsize 1)', 'append_pos': True,
            'l2_policy': 'inputs (%.1f GB per GPU) larger than the 126 MB L2; no flush needed'
                         % (n_images * (FH * FW * C * 4 + H * W * 4) / 1e9),
            'sharding': 'reference rule step = n/N + 1 over a 300*N image set, no collective'}


# ------------------------------------------------------------------------------- GPU arm
class Ctx:
    pass


def timed_steps(ctx, fn, steps, warmup=3):
    """ms per step of fn() (CUDA events on the current stream), max over ranks."""
    import torch
    import torch.distributed as dist
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if ctx.world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if ctx.world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=ctx.dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def stage_means(stage_ev):
    out = {}
    for name in ('overlap', 'pool', 'init', 'kmeans', 'paint'):
        v = [tm[name][0].elapsed_time(tm[name][1]) for tm in stage_ev if name in tm]
        if v:
            out[name + '_ms'] = float(np.mean(v))
    return out


def run_ours(args):
    import torch
    import torch.distributed as dist
    from superpixel_align_b200 import _lib, drn, ops, pipeline, synth

    ctx = Ctx()
    world = ctx.world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = ctx.rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device (the product has no CPU fallback)')
    torch.cuda.set_device(local)
    dev = ctx.dev = torch.device('cuda', local)
    numa = pipeline.bind_to_gpu_numa_node(local)      # before any pinned allocation
    if world > 1:
        if os.environ.get('NCCL_DEBUG', 'VERSION').upper() == 'VERSION':
            os.environ['NCCL_DEBUG'] = 'WARN'   # keep stdout to the one JSON line
        dist.init_process_group('nccl', device_id=dev)
    _lib.load()

    if args.config == 3:
        lo, hi = shard_range(500, world, rank)
    else:
        lo, hi = shard_range(args.images * world, world, rank)
    n_img = hi - lo
    n_sp = [GY * GX] * n_img

    # ---- inputs, resident in HBM ----
    t_setup = time.time()
    labels = synth.voronoi_labels_torch(max(n_img, 1), H, W, GY, GX, first_index=lo, device=dev)[:n_img]
    # inference form of the input producer: BatchNorm folded, conv + bias (+ shortcut) + ReLU
    # as single cuDNN calls
    model = drn.drn_c_26(device=dev, fold_bn=True, fused=True)
    feats = torch.empty((n_img, FH * FW, C), dtype=torch.float32, device=dev)
    bs = 2
    for i in range(0, n_img, bs):
        m = min(bs, n_img - i)
        imgs = synth.smooth_images_torch(m, H, W, first_index=lo + i, device=dev)
        f = synth.drn_features_torch(model, imgs)
        feats[i:i + m] = f.permute(0, 2, 3, 1).reshape(m, FH * FW, C)
    torch.cuda.synchronize()
    t_setup = time.time() - t_setup

    def step(timers=None, paint_overlap=PAINT_OVERLAP):
        np.random.seed(1111)
        return pipeline.run_batch(labels, feats, n_sp, FH, FW, k=K, prior=PRIOR, append_pos=True,
                                  images_per_group=1, timers=timers, paint_overlap=paint_overlap)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()          # runs through warm-up, the timed steps and a short tail of the same load
    out = None
    if n_img > 0:
        for _ in range(max(args.warmup, 3)):
            out = step()
        torch.cuda.synchronize()
        nnz = out.overlap.validate()
        tie_groups = out.check()
        iters = out.iters.cpu().numpy()
        status = out.status.cpu().numpy()
    else:
        nnz, tie_groups, iters, status = 0, 0, np.zeros(1), np.zeros(1)

    barrier()
    launches0 = ops.LAUNCHES
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_ev = []
    ev0.record()
    for _ in range(args.steps):
        if n_img > 0:
            step()
    ev1.record()
    barrier()
    launches = ops.LAUNCHES - launches0
    ms = ev0.elapsed_time(ev1)
    # stage attribution (K1 / K2 / K3 / K4 launch times for the roofline block): the same steps
    # once more with everything on one stream -- in the timed steps above the paint-back of
    # finished image ranges runs under the k-means tail, so the stages do not add up to a step
    for _ in range(args.steps):
        timers = {}
        if n_img > 0:
            step(timers, paint_overlap=0)
        stage_ev.append(timers)
    torch.cuda.synchronize()
    # the timed region lasts tens of milliseconds: keep the same load running (untimed) until
    # nvidia-smi has delivered enough samples taken under it
    t_tail = time.time()
    while n_img > 0 and sampler.count() < 12 and time.time() - t_tail < 2.0:
        step()
        torch.cuda.synchronize()
    clocks = sampler.stop()
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    cnt = torch.tensor([float(n_img)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    ms_max = float(t.item())
    total_images = float(cnt.item())
    value = total_images * args.steps / (ms_max / 1000.0)

    screened, exact_rows = ops.kmeans_debug_stats(reset=True)
    stages = stage_means(stage_ev)

    # ---- roofline of the dominant kernel (K2 pooling; K1 alongside) ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak = float(peaks.get('hbm_gbs', 6650.0))
    peak_src = 'measured (MEASURED_PEAKS.json)' if 'hbm_gbs' in peaks else 'fallback (B200_PROFILING.md)'
    S = GY * GX
    ld = ops.padded_ld(C + 2)
    pool_bytes = n_img * (FH * FW * C * 4 + (S + 1) * 4 + S * (4 + 8 + 8) + S * ld * 4) + nnz * 8
    k1_bytes = n_img * (H * W * 4 + (S + 1) * 4 + S * (4 + 8 + 8 + 8)) + nnz * 8
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json')))
        traffic = tj.get('pool_rows_kernel')
        if traffic is not None:
            traffic = traffic * n_img  # stored per image
            traffic_src = 'profiles/traffic.json (ncu --set full capture of pool_rows_kernel, ' \
                          'dram bytes per image x images per launch; not measured in this run)'
    except Exception:
        pass
    roofline = None
    if 'pool_ms' in stages:
        ach = pool_bytes / (stages['pool_ms'] * 1e-3) / 1e9
        roofline = {'bound': 'hbm', 'kernel': 'pool_rows_kernel (K2 CSR SpMM pooling)',
                    'achieved': ach, 'peak': peak, 'unit': 'GB/s', 'frac': ach / peak,
                    'peak_source': peak_src, 'traffic': traffic, 'traffic_source': traffic_src,
                    'algorithmic_bytes_per_launch': pool_bytes,
                    'launch_ms': stages['pool_ms']}
        if 'overlap_ms' in stages:
            k1 = k1_bytes / (stages['overlap_ms'] * 1e-3) / 1e9
            roofline['k1_overlap'] = {'achieved': k1, 'frac': k1 / peak,
                                      'algorithmic_bytes_per_launch': k1_bytes,
                                      'launch_ms': stages['overlap_ms'], 'kernels': 7}
            both = (pool_bytes + k1_bytes) / ((stages['pool_ms'] + stages['overlap_ms']) * 1e-3) / 1e9
            roofline['k1_plus_k2'] = {'achieved': both, 'frac': both / peak,
                                      'frac_of_nominal_8TBs': both / 8000.0}
        if 'paint_ms' in stages:
            pb = n_img * (H * W * 4 + 2 * H * W)
            roofline['k4_paint'] = {'achieved': pb / (stages['paint_ms'] * 1e-3) / 1e9,
                                    'frac': pb / (stages['paint_ms'] * 1e-3) / 1e9 / peak}

    extras = {}
    full = args.config == 1 and not args.quick and n_img > 0
    # ---- configs[3]: superpixel-count sweep; the reference's joint batch-30 mode ----
    if full:
        extras['s_sweep'] = run_s_sweep(ctx, args, feats, peak)
        extras['joint30'] = run_joint30(ctx, args, labels, feats, n_sp)
    # ---- configs[4]: dataset-wide clustering of the pooled descriptors ----
    if full and not args.no_global_kmeans:
        extras['global_kmeans'] = run_global_kmeans(args, dev, out, world, rank, n_img)
    # ---- the whole batch against the CPU oracle (rank 0) ----
    if full and rank == 0 and args.verify != 0:
        extras['verify'] = run_verify(args, out, labels, feats)
    if world > 1:
        dist.barrier()
    # ---- end to end, host buffers ----
    e2e = e2e_feat = dropin = None
    hl = hf = None
    if args.config == 1 and n_img > 0:
        out = None
        torch.cuda.empty_cache()
        e2e_feat = run_e2e_features(args, ctx, labels, feats)
        if not args.quick and rank == 0:
            dropin = run_dropin_numpy(args, ctx, labels, feats)
        if rank == 0 and world == 1 and not args.no_cpu_baseline:
            n_host = min(4, n_img)
            hl = [labels[i].cpu().numpy() for i in range(n_host)]
            hf = [feats[i].cpu().numpy() for i in range(n_host)]
        feats = None
        torch.cuda.empty_cache()
        if world > 1:
            dist.barrier()
        e2e = run_e2e_images(args, ctx, model, labels, lo)
        e2e['numa'] = numa
    model = None

    # ---- CPU baseline: oracle port (+ the reference's own functions) on the host cores ----
    cpu = None
    if hl is not None:
        cores = os.cpu_count() or 1
        procs = max(1, min(cores, 32))
        pool = CpuPool(hl, hf, procs)
        pool.run_port(procs)
        n_s = 2 * procs
        v, per, _ = pool.run_port(n_s)
        cpu = {'value': v, 'unit': 'images/s', 'cores': procs, 'kind': 'port',
               'sample': '%d of the 300 images, %d worker processes, oracle/spalign_oracle.py '
                         '(NumPy/SciPy float64 count-matrix path), %.2f s per image per core'
                         % (n_s, procs, per)}
        if reference_functions_available() and not args.quick:
            rv, rper = pool.run_reference_functions()
            cpu['reference_functions'] = {'value': rv, 'unit': 'images/s', 'cores': procs,
                                          'kind': 'reference', 'sample': REF_SAMPLE % (procs, rper)}
        pool.close()

    if rank == 0:
        line = {
            'metric': 'images/sec (1024x2048 hot path)', 'value': value, 'unit': 'images/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': ms_max / args.steps, 'higher_is_better': True,
            'scaling': 'strong' if args.config == 3 else 'weak',
            'vs_baseline': None, 'dtype': 'f32 pooling / f64 k-means / i32 overlap',
            'data': 'synthetic', 'config': workload_config(args, world),
            'roofline': roofline, 'cpu_baseline': cpu, 'e2e': e2e,
            'e2e_features_precomputed': e2e_feat, 'dropin_numpy': dropin,
            'gpu_launches': launches, 'clocks': clocks, 'stages_ms_per_step': stages,
            'stages_note': 'stage times from %d untimed one-stream steps after the timed region; '
                           'the timed steps run K4 of finished image ranges under the K3 tail '
                           '(paint_overlap=%d), so ms_per_step < sum of stages' % (args.steps, PAINT_OVERLAP),
            'kmeans': {'iters_mean': float(iters.mean()), 'iters_max': int(iters.max()),
                       'status_counts': {str(int(s)): int((status == s).sum()) for s in np.unique(status)},
                       'init_tie_groups': tie_groups,
                       'rows_screened_fp32': screened, 'rows_exact_f64': exact_rows},
            'nnz_per_image': nnz / max(n_img, 1), 'setup_s': t_setup,
            'us_per_image': 1000.0 * ms_max / args.steps / max(n_img, 1),
        }
        line.update(extras)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------- configs[3] sweep
def run_s_sweep(ctx, args, feats, peak):
    """Throughput and K1/K2 roofline fractions per superpixel count (same features: their bytes
    do not depend on S)."""
    import torch
    from superpixel_align_b200 import ops, pipeline, synth
    n_img = min(feats.shape[0], args.sweep_images)
    res = {}
    for S, (gy, gx) in S_GRIDS.items():
        labels = synth.voronoi_labels_torch(n_img, H, W, gy, gx, first_index=0, device=ctx.dev)
        n_sp = [gy * gx] * n_img
        stage_ev = []

        def fn():
            np.random.seed(1111)
            tm = {}
            o = pipeline.run_batch(labels, feats[:n_img], n_sp, FH, FW, k=K, prior=PRIOR, timers=tm)
            stage_ev.append(tm)
            return o
        out = fn()
        torch.cuda.synchronize()
        nnz = out.overlap.validate()
        ms = timed_steps(ctx, fn, 3, warmup=2)
        st = stage_means(stage_ev[-3:])
        ld = ops.padded_ld(C + 2)
        pool_bytes = n_img * (FH * FW * C * 4 + (S + 1) * 4 + S * 20 + S * ld * 4) + nnz * 8
        k1_bytes = n_img * (H * W * 4 + (S + 1) * 4 + S * 28) + nnz * 8
        res[str(S)] = {
            'images': n_img, 'images_per_s': n_img / ms * 1e3, 'ms_per_step': ms,
            'nnz_per_image': nnz / n_img, 'stages_ms': st,
            'k1_frac': k1_bytes / (st['overlap_ms'] * 1e-3) / 1e9 / peak,
            'k2_frac': pool_bytes / (st['pool_ms'] * 1e-3) / 1e9 / peak,
            'iters_mean': float(out.iters.float().mean().item()),
            'iters_max': int(out.iters.max().item())}
        del labels, out
    return res


def run_joint30(ctx, args, labels, feats, n_sp):
    """The reference's default mode (--batchsize 30, batch_spalign_kmeans.py:69): the superpixels
    of 30 consecutive images are clustered jointly (30 000 rows per problem)."""
    import torch
    from superpixel_align_b200 import pipeline
    n_img = (labels.shape[0] // 30) * 30
    if n_img == 0:
        return None
    stage_ev = []

    def fn():
        np.random.seed(1111)
        tm = {}
        o = pipeline.run_batch(labels[:n_img], feats[:n_img], n_sp[:n_img], FH, FW, k=K, prior=PRIOR,
                               images_per_group=30, timers=tm)
        stage_ev.append(tm)
        return o
    out = fn()
    torch.cuda.synchronize()
    ties = out.check()
    ms = timed_steps(ctx, fn, 3, warmup=2)
    return {'images': n_img, 'groups': n_img // 30, 'rows_per_group': int(sum(n_sp[:30])),
            'images_per_s': n_img / ms * 1e3, 'ms_per_step': ms, 'stages_ms': stage_means(stage_ev[-3:]),
            'iters': out.iters.cpu().tolist(), 'status': out.status.cpu().tolist(),
            'init_tie_groups': ties}


# ------------------------------------------------------------------------------ configs[4]
def run_global_kmeans(args, dev, out, world, rank, n_img):
    """BASELINE configs[4]: ONE prior-weighted k-means over the pooled descriptors of all ranks
    (rows sharded contiguously, rank r holds its 300 images' superpixels).  Timed with CUDA
    events around the device work of a whole run (init sums + all iterations), max over ranks:
    'peer' = exchange inside the iterate kernel over NVLink peer memory (one launch per
    iteration), 'nccl' = sweep/reduce/all_reduce/update (the baseline), 'local' = this rank's rows
    alone without any exchange.  A reduced problem (2 images per rank) is checked against the
    CPU oracle's kmeans() on the concatenated matrix in the same run."""
    import torch
    import torch.distributed as dist
    from superpixel_align_b200 import dist_kmeans, ops
    X, w = out.features, out.weights
    N_local, D = X.shape
    pv = K * (D + 2) + 1
    res = {'rows_per_gpu': int(N_local), 'rows_total': int(N_local) * world, 'columns': int(D),
           'K': K, 'exchange_doubles_per_iteration': pv}

    def host_init(wh):
        n = len(wh)
        init = np.zeros(n, dtype=np.int32)
        thr = float(np.sort(wh)[n // 2])
        low = wh <= thr
        idx = np.arange(int(low.sum())) % (K - 1) + 1
        np.random.shuffle(idx)
        init[low] = idx
        return init

    def timed(fn, reps=3):
        best = None
        for _ in range(reps):
            ev = []
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            r = fn(ev)
            torch.cuda.synchronize()
            ms = ev[0].elapsed_time(ev[1])
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            if best is None or ms < best[0]:
                best = (ms, r)
        ms, r = best
        it = int(r.iters[0].item())
        return {'ms': ms, 'iterations': it, 'us_per_iteration': 1e3 * ms / max(1, it),
                'status': int(r.status[0].item())}

    # this rank's rows alone (what N = 1 pays per iteration for the same rows per GPU)
    np.random.seed(1111)
    init_l = torch.from_numpy(host_init(w.cpu().numpy())).to(dev)

    def local(ev):
        km = ops.KMeansLarge(X, w, init_l, K, [0, N_local])
        return dist_kmeans._timed_run(km, 4, ev)
    res['local'] = timed(local)
    if world == 1:
        return res
    np.random.seed(1111)
    init_g = dist_kmeans.distributed_init(w.cpu().numpy(), K)
    row0 = rank * N_local
    comm = dist_kmeans.PeerComm(pv)
    res['peer'] = timed(lambda ev: dist_kmeans.global_kmeans(
        X, w, K, init_local=init_g, exchange='peer', comm=comm, row0=row0, events=ev))
    res['nccl'] = timed(lambda ev: dist_kmeans.global_kmeans(
        X, w, K, init_local=init_g, exchange='nccl', row0=row0, events=ev))
    # the all-reduce alone: pv float64 values, back to back
    buf = torch.zeros(pv, dtype=torch.float64, device=dev)
    for _ in range(5):
        dist.all_reduce(buf)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    dist.barrier()
    e0.record()
    for _ in range(50):
        dist.all_reduce(buf)
    e1.record()
    torch.cuda.synchronize()
    res['nccl_allreduce_us'] = 1e3 * e0.elapsed_time(e1) / 50
    res['nccl']['pct_in_allreduce'] = 100.0 * res['nccl_allreduce_us'] / res['nccl']['us_per_iteration']
    res['peer_vs_local_us_per_iteration'] = res['peer']['us_per_iteration'] / res['local']['us_per_iteration']
    # reduced problem against the oracle (rank 0 runs the reference kmeans on the whole matrix)
    n_small = min(n_img, 2) * (N_local // n_img)
    Xs, ws = X[:n_small].contiguous(), w[:n_small].contiguous()
    np.random.seed(1111)
    init_s = dist_kmeans.distributed_init(ws.cpu().numpy(), K)
    r = dist_kmeans.global_kmeans(Xs, ws, K, init_local=init_s, exchange='peer', comm=comm,
                                  row0=rank * n_small)
    gX = [torch.empty_like(Xs) for _ in range(world)]
    gw = [torch.empty_like(ws) for _ in range(world)]
    ga = [torch.empty_like(r.assign) for _ in range(world)]
    gi = [torch.empty(n_small, dtype=torch.int32, device=dev) for _ in range(world)]
    dist.all_gather(gX, Xs)
    dist.all_gather(gw, ws)
    dist.all_gather(ga, r.assign)
    dist.all_gather(gi, torch.from_numpy(np.asarray(init_s, dtype=np.int32)).to(dev))
    if rank == 0:
        from oracle import spalign_oracle as so     # checker only
        want, info = so.kmeans(K, torch.cat(gX).cpu().numpy().astype(np.float64),
                               torch.cat(gw).cpu().numpy(),
                               init_assign=torch.cat(gi).cpu().numpy().astype(np.float64),
                               return_info=True, verbose=False)
        got = torch.cat(ga).cpu().numpy()
        res['check_vs_oracle'] = {
            'rows': int(n_small) * world,
            'identical_assignments': bool(np.array_equal(got, np.asarray(want).astype(np.int32))),
            'iterations': [int(r.iters[0].item()), int(info['iters'])],
            'status': [int(r.status[0].item()), int(info['status'])]}
    dist.barrier()
    comm.close()
    return res


# ----------------------------------------------------------------------------------- verify
def synthetic_road_gt():
    """Ground truth for the road-IoU acceptance line: a trapezoid 'road' in the lower half, void
    (-1) band at the bottom like the ego-vehicle region (labels: -1 void, 1 road, 0 other)."""
    yy, xx = np.mgrid[0:H, 0:W]
    half = (yy - H * 0.45) / (H * 0.55) * (W * 0.45) + W * 0.05
    gt = ((yy > H * 0.45) & (np.abs(xx - W * 0.5) < half)).astype(np.int32)
    gt[int(H * 0.94):] = -1
    return gt


def run_verify(args, out, labels, feats):
    """Acceptance lines of BASELINE.json over the whole batch (rank 0's images) against the CPU
    oracle: (1) k-means assignments identical to the reference semantics under the same init on
    the descriptors the GPU pooled; (2) end to end (the oracle's own float64 CSR + pooling + prior
    + k-means + paint from the raw inputs): assignment flips <= 0.1 %, road-IoU delta <= 1e-3 per
    image against a synthetic ground truth, pooled features rtol."""
    import multiprocessing as mp
    from oracle import spalign_oracle as so     # checker only
    n_all = labels.shape[0]
    n = n_all if args.verify < 0 else min(args.verify, n_all)
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 32))
    t0 = time.time()
    w_all = out.weights.cpu().numpy()
    a_all = out.assign.cpu().numpy()
    it_all = out.iters.cpu().numpy()
    st_all = out.status.cpu().numpy()
    off = out.group_off_host
    # replay the seeded stream in image order (draw_shuffles consumed it the same way)
    np.random.seed(1111)
    inits = []
    for g in range(n_all):
        wg = w_all[off[g]:off[g + 1]]
        inits.append(so.kmeans_init(K, wg).astype(np.int32))
    gt = synthetic_road_gt()
    res = {'images': int(n), 'procs': procs}
    # (1) reference k-means semantics on the GPU's descriptors
    X_all = out.features.cpu().numpy()
    jobs = [(g, X_all[off[g]:off[g + 1]], w_all[off[g]:off[g + 1]], inits[g]) for g in range(n)]
    with mp.get_context('fork').Pool(procs) as pool:
        got = pool.map(_cpu_one_given_descriptors, jobs, chunksize=4)
    bad_rows = sum(int((a != a_all[off[g]:off[g + 1]]).sum()) for g, (a, _, _) in enumerate(got))
    res['kmeans_on_gpu_descriptors'] = {
        'rows': int(off[n]), 'assignment_mismatches': bad_rows,
        'iteration_count_mismatches': int(sum(it != it_all[g] for g, (_, it, _) in enumerate(got))),
        'status_mismatches': int(sum(s != st_all[g] for g, (_, _, s) in enumerate(got)))}
    # (2) end to end from the raw inputs, in chunks (67 MB of features per image on the host)
    flips = rows = 0
    max_diou = max_feat_rel = 0.0
    images_identical = 0
    chunk = 48
    mask_d = out.road_mask
    for c0 in range(0, n, chunk):
        c1 = min(n, c0 + chunk)
        hl = [labels[i].cpu().numpy() for i in range(c0, c1)]
        hf = [feats[i].cpu().numpy() for i in range(c0, c1)]
        pool = CpuPool(hl, hf, min(procs, c1 - c0))
        r = pool.pool.map(_cpu_one, [(i - c0, 0, inits[i]) for i in range(c0, c1)], chunksize=1)
        pool.close()
        for i, (_, oa, of, oroad) in zip(range(c0, c1), r):
            ga = a_all[off[i]:off[i + 1]]
            d = int((oa != ga).sum())
            flips += d
            rows += len(ga)
            images_identical += int(d == 0)
            gf = X_all[off[i]:off[i + 1]]
            max_feat_rel = max(max_feat_rel, float(np.max(np.abs(gf - of) / (np.abs(of) + 1e-3))))
            groad = mask_d[i].cpu().numpy().astype(bool)
            oroad = np.unpackbits(oroad)[:H * W].reshape(H, W).astype(bool)
            iou_g = so.road_iou(groad, gt)[0]
            iou_o = so.road_iou(oroad, gt)[0]
            max_diou = max(max_diou, abs(float(iou_g) - float(iou_o)))
    res['end_to_end'] = {
        'rows': rows, 'assignment_flips': flips, 'flip_fraction': flips / max(rows, 1),
        'images_with_identical_assignments': images_identical,
        'max_road_iou_delta': max_diou, 'max_pooled_feature_rel_err': max_feat_rel,
        'pass': bool(flips / max(rows, 1) <= 1e-3 and max_diou <= 1e-3)}
    res['seconds'] = time.time() - t0
    return res


# -------------------------------------------------------------------------- end to end legs
def run_e2e_images(args, ctx, model, labels, first_index):
    """`e2e`: the reference's boundary (estimate_road_mask, batch_spalign_kmeans.py:427-457)
    through superpixel_align_b200.pipeline.ImagePipeline: every step copies uint8 RGB images
    and uint16 label maps from pinned HOST memory to the device, runs the DRN backbone
    (PyTorch/cuDNN, fp32 with TF32 convolutions, channels_last) and K1..K4 on the device, and
    copies the uint8 cluster maps and road masks back to pinned host memory.  The features never
    cross PCIe, as in the reference (:431-435)."""
    import torch
    import torch.distributed as dist
    from superpixel_align_b200 import pipeline, synth
    dev, world = ctx.dev, ctx.world
    n_img = labels.shape[0]
    pool_n = min(args.host_pool, n_img)
    sub = min(args.e2e_sub_batch, pool_n)
    h_img = torch.empty((pool_n, 3, H, W), dtype=torch.uint8).pin_memory()
    for i in range(pool_n):
        im = synth.smooth_images_torch(1, H, W, first_index=first_index + i, device=dev)
        h_img[i].copy_(im[0].round().clamp(0, 255).to(torch.uint8))
    h_lab = labels[:pool_n].to(torch.int16).cpu().pin_memory()   # ids < 65536: uint16 on the wire
    n_e2e = max(sub, (min(args.e2e_images, n_img) // sub) * sub)
    batches = []
    for i in range(0, n_e2e, sub):
        j = i % (pool_n - sub + 1) if pool_n > sub else 0
        batches.append((h_img[j:j + sub], h_lab[j:j + sub], [GY * GX] * sub))
    ip = pipeline.ImagePipeline(model, H, W, sub_batch=sub, k=K, prior=PRIOR, device=dev)
    sink = {'road_px': 0}

    def on_result(i, cmap, mask):
        sink['road_px'] += int(mask[0, ::64, ::64].sum())     # touch the host copy

    np.random.seed(1111)
    ip.process(batches[:3], on_result)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ip.h2d_bytes = ip.d2h_bytes = 0
    steps = max(1, min(args.steps, 3))
    t0 = time.time()
    for _ in range(steps):
        ip.process(batches, on_result, time_backbone=True)
    torch.cuda.synchronize()
    dt = time.time() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    bb = [e0.elapsed_time(e1) / b for e0, e1, b in ip.backbone_ms]
    return {'value': world * n_e2e * steps / dt, 'unit': 'images/s',
            'h2d_bytes_per_step': ip.h2d_bytes // steps, 'd2h_bytes_per_step': ip.d2h_bytes // steps,
            'images_per_step': n_e2e, 'steps': steps, 'sub_batch': sub,
            'backbone_ms_per_image': float(np.mean(bb)) if bb else None,
            'backbone': 'DRN-C-26 (random init; inference form: BatchNorm folded, conv+bias(+shortcut)+ReLU '
                        'fused cuDNN calls), fp32 / TF32 convolutions, channels_last',
            'h2d_GBps_per_gpu': ip.h2d_bytes / dt / 1e9,
            'api': 'superpixel_align_b200.pipeline.ImagePipeline.process: pinned host uint8 '
                   '[b,3,H,W] images + uint16 label maps in, DRN + K1..K4 on the device, uint8 '
                   'cluster maps + road masks out to pinned host memory; double-buffered '
                   'copy/compute streams; per-image clustering'}


def run_e2e_features(args, ctx, labels, feats):
    """`e2e_features_precomputed` (round 1's e2e): the hot path alone through
    pipeline.HostPipeline with the fp32 feature maps in HOST memory -- 75.5 MB per image over
    PCIe, so this leg measures the host->device path, not the kernels."""
    import torch
    import torch.distributed as dist
    from superpixel_align_b200 import pipeline
    dev, world = ctx.dev, ctx.world
    n_img = labels.shape[0]
    pool_n = min(args.host_pool, n_img)
    sub = min(8, pool_n)
    h_lab = labels[:pool_n].cpu().pin_memory()
    h_feat = feats[:pool_n].cpu().pin_memory()
    n_e2e = max(sub, (min(96, n_img) // sub) * sub)
    batches = []
    for i in range(0, n_e2e, sub):
        j = i % (pool_n - sub + 1) if pool_n > sub else 0
        batches.append((h_lab[j:j + sub], h_feat[j:j + sub], [GY * GX] * sub))
    hp = pipeline.HostPipeline(H, W, FH, FW, C, sub_batch=sub, k=K, prior=PRIOR, device=dev)
    sink = {'road_px': 0}

    def on_result(i, cmap, mask):
        sink['road_px'] += int(mask[0, ::64, ::64].sum())

    np.random.seed(1111)
    hp.process(batches[:3], on_result)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    hp.h2d_bytes = hp.d2h_bytes = 0
    steps = max(1, min(args.steps, 3))
    t0 = time.time()
    for _ in range(steps):
        hp.process(batches, on_result)
    torch.cuda.synchronize()
    dt = time.time() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    return {'value': world * n_e2e * steps / dt, 'unit': 'images/s',
            'h2d_bytes_per_step': hp.h2d_bytes // steps, 'd2h_bytes_per_step': hp.d2h_bytes // steps,
            'images_per_step': n_e2e, 'sub_batch': sub, 'h2d_GBps_per_gpu': hp.h2d_bytes / dt / 1e9,
            'note': 'bound by the host->device copy of 67 MB of fp32 features per image; all ranks '
                    'of one box share the host memory / PCIe root complexes'}


def run_dropin_numpy(args, ctx, labels, feats):
    """`dropin_numpy`: the reference-named sequence exactly as estimate_road_mask calls it
    (batch_spalign_kmeans.py:444-457) -- NumPy arrays in (int64 label maps as skimage yields,
    fp32 NCHW feature maps), NumPy arrays out, pageable memory -- and the same sequence with the
    explicit state handle / torch CUDA carriers."""
    import types
    import torch
    from superpixel_align_b200 import batch_spalign_kmeans as bsk
    n = min(8, labels.shape[0])
    a = types.SimpleNamespace(gpu=ctx.dev.index or 0, n_clusters=K, without_pos=False, y_rel_pos=PRIOR[0],
                              x_rel_pos=PRIOR[1], y_rel_sigma=PRIOR[2], x_rel_sigma=PRIOR[3])
    lab_np = labels[:n].cpu().numpy().astype(np.int64)
    f_nchw = feats[:n].reshape(n, FH, FW, C).permute(0, 3, 1, 2).contiguous()
    f_np = f_nchw.cpu().numpy()
    res = {'images': n}

    def seq_numpy():
        np.random.seed(1111)
        f, n_per = bsk.batch_superpixel_align(a, None, None, lab_np, f_np)
        w = bsk.batch_create_prior(a, lab_np)
        return bsk.batch_weighted_kmeans(a, lab_np, f, w, n_per)

    def seq_torch():
        np.random.seed(1111)
        st = bsk.prepare_batch(a, labels[:n], feature_shape=(FH, FW))
        f, n_per = bsk.batch_superpixel_align(a, None, None, st, f_nchw)
        w = bsk.batch_create_prior(a, st)
        return bsk.batch_weighted_kmeans(a, st, f, w, n_per)

    for name, fn in (('numpy_in_numpy_out', seq_numpy), ('torch_cuda_with_state_handle', seq_torch)):
        bsk.clear_cache()
        fn()
        torch.cuda.synchronize()
        t0 = time.time()
        reps = 2
        for _ in range(reps):
            bsk.clear_cache()
            fn()
        torch.cuda.synchronize()
        dt = (time.time() - t0) / reps
        res[name] = {'images_per_s': n / dt, 'ms_per_image': 1e3 * dt / n}
    c_np, _ = seq_numpy()
    c_t, _ = seq_torch()
    res['same_cluster_maps'] = bool(np.array_equal(np.asarray(c_np), c_t.cpu().numpy()))
    bsk.clear_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', type=int, default=1, choices=[1, 3])
    ap.add_argument('--images', type=int, default=300, help='images per GPU per step (config 1)')
    ap.add_argument('--e2e-images', type=int, default=48, help='images per e2e step')
    ap.add_argument('--e2e-sub-batch', type=int, default=8, help='images per host->device sub-batch')
    ap.add_argument('--host-pool', type=int, default=16, help='distinct pinned host images')
    ap.add_argument('--sweep-images', type=int, default=300, help='images per S of the sweep')
    ap.add_argument('--verify', type=int, default=-1,
                    help='images of the batch checked against the CPU oracle (-1 = all, 0 = none)')
    ap.add_argument('--quick', action='store_true', help='main line + e2e only')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-global-kmeans', action='store_true')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
