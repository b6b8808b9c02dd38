import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from superpixel_align_b200 import pipeline, synth, ops
n = 300
H, W, FH, FW, C, GY, GX = 1024, 2048, 128, 256, 512, 25, 40
dev = torch.device('cuda', 0)
labels = synth.voronoi_labels_torch(n, H, W, GY, GX, first_index=0, device=dev)
g = torch.Generator(device=dev).manual_seed(3)
feats = torch.empty((n, FH * FW, C), dtype=torch.float32, device=dev)
for i in range(0, n, 10):
    m = min(10, n - i)
    base = torch.randn((m, 8, C), generator=g, device=dev)
    coef = torch.rand((m, FH * FW, 8), generator=g, device=dev)
    yy = torch.arange(FH * FW, device=dev) // FW
    coef[:, :, 0] = (yy / FH)[None, :] * 3
    feats[i:i + m] = torch.relu(coef @ base + 0.3 * torch.randn((m, FH * FW, C), generator=g, device=dev))
n_sp = [GY * GX] * n
for _ in range(2):
    np.random.seed(1111)
    out = pipeline.run_batch(labels, feats, n_sp, FH, FW)
torch.cuda.synchronize()
ops.kmeans_debug_stats(reset=True)
np.random.seed(1111)
out = pipeline.run_batch(labels, feats, n_sp, FH, FW)
torch.cuda.synchronize()
print('iters mean %.1f max %d' % (out.iters.float().mean().item(), out.iters.max().item()))
ops.kmeans_debug_stats(reset=True)
