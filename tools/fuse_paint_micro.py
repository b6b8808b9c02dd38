import sys; sys.path.insert(0, ".")
import numpy as np, torch
from superpixel_align_b200 import pipeline, synth
n=300; dev=torch.device("cuda",0)
labels = synth.voronoi_labels_torch(n, 1024, 2048, 25, 40, device=dev)
g = torch.Generator(device=dev).manual_seed(3)
feats = torch.empty((n, 128*256, 512), device=dev)
for i in range(0, n, 10):
    base = torch.randn((10, 8, 512), generator=g, device=dev); coef = torch.rand((10, 128*256, 8), generator=g, device=dev)
    yy = torch.arange(128*256, device=dev)//256; coef[:, :, 0] = (yy/128)[None,:]*3
    feats[i:i+10] = torch.relu(coef @ base + 0.3*torch.randn((10,128*256,512), generator=g, device=dev))
for fuse in (True, False, True):
    for _ in range(3):
        np.random.seed(1111); tm={}; out = pipeline.run_batch(labels, feats, [1000]*n, 128, 256, fuse_paint=fuse, timers=tm)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True); e0.record()
    for _ in range(5):
        np.random.seed(1111); tm={}; out = pipeline.run_batch(labels, feats, [1000]*n, 128, 256, fuse_paint=fuse, timers=tm)
    e1.record(); torch.cuda.synchronize()
    print("fuse_paint", fuse, "%.3f ms/step" % (e0.elapsed_time(e1)/5), {k: round(v[0].elapsed_time(v[1]),3) for k,v in tm.items()})
