"""Single launches of the fused bottleneck kernel (csrc/conv_rb.cu) and of the memory-bound 1x1 layers around it,
timed with CUDA events (warm) — or one launch each under ncu (RGBD_NCU=1)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.nn as nn
DEV = torch.device("cuda:0")
NCU = os.environ.get("RGBD_NCU") == "1"
from rgbd_b200.engine import Builder, PackedConv

def timeit(b, name, extra=""):
    for _ in range(0 if NCU else 3):
        b.prog.run()
    torch.cuda.synchronize()
    reps = 1 if NCU else 5
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        b.prog.run()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / reps
    fl = sum(getattr(op, "flops", 0) for op in b.prog.ops)
    by = sum(getattr(op, "bytes", 0) for op in b.prog.ops)
    print(f"== {name}: {us:.0f} us, {fl/us/1e6:.0f} TFLOP/s, {by/us/1e3:.0f} GB/s algorithmic {extra}", flush=True)

def fused(N, H, W, cin=192, final_relu=False):
    b = Builder(DEV, torch.bfloat16, tensor_cores=True)
    x = b.alloc(N, H, W, cin); x.buf.normal_()
    c1, c2, c3 = nn.Conv2d(cin, 96, 1), nn.Conv2d(96, 96, 3, 1, 1), nn.Conv2d(96, 192, 1)
    pcs = [PackedConv(m, DEV) for m in (c1, c2, c3)]
    res = x
    if cin != 192:
        res = b.alloc(N, H, W, 192); res.buf.normal_()
    b.fused_block(*pcs, x, res=res, final_relu=final_relu)
    timeit(b, f"fused 1x1-3x3-1x1 {cin}->96->96->192 @{H}x{W} N={N}")

def conv(name, mod, N, H, W, res=False, gate=False):
    b = Builder(DEV, torch.bfloat16, tensor_cores=True)
    x = b.alloc(N, H, W, mod.in_channels); x.buf.normal_()
    r = m = None
    if res:
        r = b.alloc(N, H, W, mod.out_channels); r.buf.normal_()
    if gate:
        m = b.alloc(N, H, W, mod.out_channels); m.buf.normal_()
    b.conv(PackedConv(mod, DEV), x, res=r, mul=m, epi=1 if gate else 0)
    timeit(b, f"{name} N={N}")

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16
fused(N, 256, 320)
fused(N, 128, 160)
fused(N, 64, 80)
conv("1x1 48->192 gate @256x320", nn.Conv2d(48, 192, 1), N, 256, 320, gate=True)
conv("1x1 384->48 @256x320", nn.Conv2d(384, 48, 1), N, 256, 320)
conv("1x1 384->192 @256x320", nn.Conv2d(384, 192, 1), N, 256, 320)
conv("3x3 96->96 @256x320", nn.Conv2d(96, 96, 3, 1, 1), N, 256, 320)
