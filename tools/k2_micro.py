"""K2 micro-benchmark: pool over n synthetic images (distinct label maps, random features).
python tools/k2_micro.py [n_img] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from superpixel_align_b200 import ops, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device('cuda', 0)
labels = synth.voronoi_labels_torch(n, 1024, 2048, 25, 40, device=dev)
g = torch.Generator(device=dev).manual_seed(1)
feats = torch.empty((n, 128 * 256, 512), dtype=torch.float32, device=dev)
for i in range(0, n, 20):
    feats[i:i + 20].normal_(generator=g)
ov = ops.overlap_csr(labels, 128, 256, [1000] * n, prior=(0.75, 0.5, 0.1, 0.1))
for _ in range(3):
    out = ops.pool(feats, ov)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    out = ops.pool(feats, ov)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
nnz = ov.validate()
bytes_ = n * (32768 * 512 * 4 + 1001 * 4 + 1000 * 20 + 1000 * 516 * 4) + nnz * 8
print('n_img=%d  K2 %.3f ms  %.2f us/image  %.0f GB/s algorithmic  checksum %.6e' %
      (n, ms, 1000 * ms / n, bytes_ / ms / 1e6, out.double().sum().item()))
