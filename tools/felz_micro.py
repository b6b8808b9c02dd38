import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from superpixel_align_b200 import ops, synth
dev = torch.device('cuda', 0)
for (n, H, W) in ((30, 224, 224), (148, 224, 224), (4, 1024, 2048)):
    imgs = synth.smooth_images_torch(n, H, W, first_index=0, device=dev).float() / 255.0
    for _ in range(2):
        lab, nl = ops.felzenszwalb(imgs, 300.0, 0.8, 20)
    torch.cuda.synchronize()
    t = time.time()
    lab, nl = ops.felzenszwalb(imgs, 300.0, 0.8, 20)
    torch.cuda.synchronize()
    dt = time.time() - t
    print('n=%d %dx%d: %.1f ms per batch, %.2f ms per image, segments %s' % (n, H, W, dt * 1e3, dt * 1e3 / n, nl[:4].tolist()))
