"""cProfile of the NumPy-in / NumPy-out drop-in sequence (batch_superpixel_align ->
batch_create_prior -> batch_weighted_kmeans) on 8 synthetic 1024x2048 images."""
import cProfile
import os
import pstats
import sys
import time
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from superpixel_align_b200 import batch_spalign_kmeans as bsk, synth

n, H, W, FH, FW, C = 8, 1024, 2048, 128, 256, 512
dev = torch.device('cuda', 0)
lab_np = synth.voronoi_labels_torch(n, H, W, 25, 40, device=dev).cpu().numpy().astype(np.int64)
f_np = np.random.RandomState(0).standard_normal((n, C, FH, FW)).astype(np.float32)
a = types.SimpleNamespace(gpu=0, n_clusters=4, without_pos=False, y_rel_pos=0.75, x_rel_pos=0.5,
                          y_rel_sigma=0.1, x_rel_sigma=0.1)


def seq():
    np.random.seed(1111)
    f, n_per = bsk.batch_superpixel_align(a, None, None, lab_np, f_np)
    w = bsk.batch_create_prior(a, lab_np)
    return bsk.batch_weighted_kmeans(a, lab_np, f, w, n_per)


seq()
torch.cuda.synchronize()
bsk.clear_cache()
t0 = time.time()
pr = cProfile.Profile()
pr.enable()
seq()
torch.cuda.synchronize()
pr.disable()
print('%.1f ms per image' % ((time.time() - t0) * 1e3 / n))
pstats.Stats(pr).sort_stats('cumulative').print_stats(28)
