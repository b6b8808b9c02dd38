"""K1 micro-benchmark: overlap_csr over n synthetic label maps (no features, no DRN).
python tools/k1_micro.py [n_img] [reps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from superpixel_align_b200 import ops, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device('cuda', 0)
pool = n   # distinct maps: a small pool would sit in the 126 MB L2 and flatter K1 (round 2 lesson)
base = synth.voronoi_labels_torch(pool, 1024, 2048, 25, 40, device=dev)
labels = base.repeat((n + pool - 1) // pool, 1, 1)[:n].contiguous()
n_sp = [1000] * n
for _ in range(3):
    ov = ops.overlap_csr(labels, 128, 256, n_sp, prior=(0.75, 0.5, 0.1, 0.1))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    ov = ops.overlap_csr(labels, 128, 256, n_sp, prior=(0.75, 0.5, 0.1, 0.1))
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print('n_img=%d  K1 %.3f ms  %.2f us/image  nnz=%d' % (n, ms, 1000 * ms / n, ov.validate()))
