"""K3 micro-benchmark: per-image k-means over G synthetic descriptor groups (no K1/K2/DRN).
python tools/k3_micro.py [groups] [rows] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from superpixel_align_b200 import ops

G = int(sys.argv[1]) if len(sys.argv) > 1 else 300
S = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device('cuda', 0)
g = torch.Generator(device=dev).manual_seed(1)
N = G * S
cent = torch.randn((G, 6, 514), generator=g, device=dev) * 0.7
lab = torch.randint(0, 6, (N,), generator=g, device=dev)
X = torch.zeros((N, 516), device=dev)
X[:, :514] = cent.reshape(-1, 514)[torch.arange(N, device=dev) // S * 6 + lab] + \
    torch.randn((N, 514), generator=g, device=dev)
X[:, 512] = torch.rand(N, generator=g, device=dev) * 1023
X[:, 513] = torch.rand(N, generator=g, device=dev) * 2047
w = torch.rand(N, generator=g, device=dev, dtype=torch.float64)
init = torch.randint(0, 4, (N,), generator=g, device=dev, dtype=torch.int32)
goff = np.arange(G + 1) * S
Xv = X[:, :514]
KW = dict(tail=os.environ.get('K3_TAIL', '1') == '1',
          tail_after=int(os.environ.get('K3_TAIL_AFTER', '1')),
          tail_slice=int(os.environ.get('K3_SLICE', '0')))
for _ in range(2):
    res = ops.KMeansLarge(Xv, w, init, 4, goff, **KW).run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
l0 = ops.LAUNCHES
e0.record()
for _ in range(reps):
    res = ops.KMeansLarge(Xv, w, init, 4, goff, **KW).run()
e1.record()
torch.cuda.synchronize()
it = res.iters.cpu().numpy()
ms = e0.elapsed_time(e1) / reps
print('G=%d S=%d  K3 %.3f ms  iters mean %.1f max %d  %.2f us per image-iteration  launches/run %d'
      % (G, S, ms, it.mean(), it.max(), 1000 * ms / it.sum(), (ops.LAUNCHES - l0) // reps))
print('screened/exact', ops.kmeans_debug_stats(reset=True))

if os.environ.get('K3_TIMELINE'):
    # per-launch timeline: CUDA events around every sweep of one blocking run
    km = ops.KMeansLarge(Xv, w, init, 4, goff, tail=False)
    evs = []
    orig = km._sweep

    def timed(mode):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        orig(mode)
        b.record()
        evs.append((mode, km.n_chunks, a, b))
    km._sweep = timed
    km.run(blocking=True, poll=1, first_poll=1)
    torch.cuda.synchronize()
    tot = 0.0
    for i, (mode, nch, a, b) in enumerate(evs):
        ms_i = a.elapsed_time(b)
        tot += ms_i
        print('launch %2d mode %d chunks %4d  %.1f us' % (i, mode, nch, 1000 * ms_i))
    print('sum of launches %.3f ms' % tot)
