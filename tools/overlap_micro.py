"""Whole hot path on synthetic inputs (no DRN): sequential run_batch vs run_batch_overlapped.
python tools/overlap_micro.py [images]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from superpixel_align_b200 import pipeline, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
H, W, FH, FW, C, GY, GX = 1024, 2048, 128, 256, 512, 25, 40
dev = torch.device('cuda', 0)
labels = synth.voronoi_labels_torch(n, H, W, GY, GX, first_index=0, device=dev)
g = torch.Generator(device=dev).manual_seed(3)
feats = torch.empty((n, FH * FW, C), dtype=torch.float32, device=dev)
for i in range(0, n, 10):
    # smooth low-rank "feature maps" so that the clustering behaves like real descriptors
    m = min(10, n - i)
    base = torch.randn((m, 8, C), generator=g, device=dev)
    coef = torch.rand((m, FH * FW, 8), generator=g, device=dev)
    yy = torch.arange(FH * FW, device=dev) // FW
    coef[:, :, 0] = (yy / FH)[None, :] * 3
    feats[i:i + m] = torch.relu(coef @ base + 0.3 * torch.randn((m, FH * FW, C), generator=g, device=dev))
n_sp = [GY * GX] * n


def timed(fn, reps=5):
    for _ in range(3):
        np.random.seed(1111)
        out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        np.random.seed(1111)
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


ms, ref = timed(lambda: pipeline.run_batch(labels, feats, n_sp, FH, FW))
print('sequential          %.3f ms  (%.0f images/s)  iters mean %.1f max %d' %
      (ms, n / ms * 1e3, ref.iters.float().mean().item(), ref.iters.max().item()))
for po in (10, 12, 15, 20):
    ms, out = timed(lambda: pipeline.run_batch(labels, feats, n_sp, FH, FW, paint_overlap=po))
    same = bool((out.cluster_map == ref.cluster_map).all()) and bool((out.iters == ref.iters).all()) \
        and bool((out.road_mask == ref.road_mask).all())
    print('paint_overlap=%2d  %.3f ms  (%.0f images/s)  identical=%s' % (po, ms, n / ms * 1e3, same))
if os.environ.get('OVERLAP_SUBBATCH'):
  for sb, ns in ((150, 2), (50, 6), (30, 10), (20, 15), (20, 5), (50, 3)):
      if sb >= n:
          continue
      ms, out = timed(lambda: pipeline.run_batch_overlapped(labels, feats, n_sp, FH, FW, sub_batch=sb,
                                                            n_streams=ns))
      same = bool((out.cluster_map == ref.cluster_map).all()) and bool((out.iters == ref.iters).all())
      print('overlapped sb=%3d streams=%d  %.3f ms  (%.0f images/s)  identical=%s' %
            (sb, ns, ms, n / ms * 1e3, same))
