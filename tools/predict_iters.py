"""Does the number of rows that change in the first iterations predict a group's iteration count?
(scheduling experiment for the k-means finish kernel)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from superpixel_align_b200 import ops, pipeline, synth

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
ov = ops.overlap_csr(labels, FH, FW, [GY * GX] * n, prior=(0.75, 0.5, 0.1, 0.1))
X = ops.pool(feats, ov, append_pos=True)
w = ov.weights()
np.random.seed(1111)
flat, off, _ = pipeline.draw_shuffles(4, [GY * GX] * n)
init, _ = ops.kmeans_init_device(w, ov.sp_off, torch.from_numpy(flat).to(dev), torch.from_numpy(off).to(dev))
km = ops.KMeansLarge(X, w, init, 4, ov.sp_off_host, tail=False)
km.init_centers()
sig = []
for it in range(3):
    km._sweep(1)
    torch.cuda.synchronize()
    sig.append(km.totals[:, -1].cpu().numpy().copy())
    if it == 0:
        cd = km.cdelta.cpu().numpy().max(1)
res = km.run(blocking=True)
iters = res.iters.cpu().numpy()
from scipy.stats import spearmanr
for i, s_ in enumerate(sig):
    print('changed rows in iteration %d vs total iterations: spearman %.3f' % (i + 1, spearmanr(s_, iters)[0]))
print('max centre drift after iteration 1 vs iterations: spearman %.3f' % spearmanr(cd, iters)[0])
top = np.argsort(-iters)[:30]
for name, s_ in (('changed@1', sig[0]), ('changed@2', sig[1]), ('drift@1', cd)):
    rank = np.argsort(-s_)
    pos = {g_: r for r, g_ in enumerate(rank)}
    print(name, 'ranks of the 30 longest groups:', sorted(pos[g_] for g_ in top))
