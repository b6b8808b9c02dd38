"""DRN-C-26 forward at 1024x2048 (the input producer of the e2e leg): ms per image for a few
cuDNN settings.  python tools/backbone_micro.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from superpixel_align_b200 import drn

dev = torch.device('cuda', 0)
model = drn.drn_c_26(device=dev)
folded = drn.drn_c_26(device=dev, fold_bn=True)
fused = drn.drn_c_26(device=dev, fold_bn=True, fused=True)
MODELS = {'plain': model, 'folded_bn': folded, 'fused_conv_relu': fused}


def run(b, dtype, bench, reps=4, model=model):
    torch.backends.cudnn.benchmark = bench
    x = torch.randn((b, 3, 1024, 2048), device=dev).contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        for _ in range(2):
            if dtype is None:
                y = model(x)
            else:
                with torch.autocast('cuda', dtype=dtype):
                    y = model(x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            if dtype is None:
                y = model(x)
            else:
                with torch.autocast('cuda', dtype=dtype):
                    y = model(x)
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps / b


x = torch.randn((2, 3, 256, 512), device=dev).contiguous(memory_format=torch.channels_last)
with torch.no_grad():
    ref = model(x)
    for name, mm in MODELS.items():
        print('%s: max abs diff vs plain %.3e (max abs %.3e)' % (name, (mm(x) - ref).abs().max().item(), ref.abs().max().item()))
for name, mm in MODELS.items():
    for b in (4, 8):
        print('%s batch=%d fp32/tf32: %.2f ms/image' % (name, b, run(b, None, False, model=mm)))
for bench in (False,):
    for b in (8,):
        for dtype in (None, torch.bfloat16):
            try:
                ms = run(b, dtype, bench)
                print('cudnn.benchmark=%s batch=%d dtype=%s: %.2f ms/image' % (bench, b, dtype or 'fp32/tf32', ms))
            except Exception as e:
                print('cudnn.benchmark=%s batch=%d dtype=%s: %r' % (bench, b, dtype, e))
print('tf32 conv allowed:', torch.backends.cudnn.allow_tf32, 'matmul:', torch.backends.cuda.matmul.allow_tf32)
