"""Where does each warp role of conv_halo_kernel stall?  Cycle counters of CTA 0 for single layers."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.nn as nn
DEV = "cuda:0"
dbg = torch.zeros(16, dtype=torch.int64, device=DEV)
NCU = os.environ.get("RGBD_NCU") == "1"   # under ncu: no tracing, one launch per layer
if not NCU:
    os.environ["RGBD_TC_TRACE"] = str(dbg.data_ptr())
from rgbd_b200.engine import Builder, PackedConv, View
def run(name, mod, N, H, W, res=False, gate=False):
    b = Builder(torch.device(DEV), torch.bfloat16, tensor_cores=True)
    x = b.alloc(N, H, W, mod.in_channels); x.buf.normal_()
    r = None
    if res:
        r = b.alloc(N, H // mod.stride[0], W // mod.stride[0], mod.out_channels); r.buf.normal_()
    m = None
    if gate:
        m = b.alloc(N, H, W, mod.out_channels); m.buf.normal_()
    out = b.conv(PackedConv(mod, torch.device(DEV)), x, res=r, mul=m, epi=1 if gate else 0)
    for _ in range(0 if NCU else 3): b.prog.run()
    torch.cuda.synchronize(); dbg.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); b.prog.run(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3
    d = dbg.cpu().tolist()
    fl = sum(getattr(op, "flops", 0) for op in b.prog.ops)
    n = max(1, d[6])
    print(f"== {name}: {us:.0f} us, {fl/us/1e6:.0f} TFLOP/s | CTA0: {d[6]} tiles, mma thread total {d[5]} cyc ({d[5]/n:.0f}/tile): "
          f"wait tmem_empty {d[2]/n:.0f} a_full {d[3]/n:.0f} b_full {d[4]/n:.0f} | producers wait: a_empty {d[0]/n:.0f} b_empty {d[1]/n:.0f} "
          f"| epilogue warp: total {d[8]/n:.0f}/tile, wait tmem_full {d[7]/n:.0f} [ld wait {d[9]/n:.0f} process {d[10]/n:.0f} issue+store {d[11]/n:.0f}]", flush=True)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
run("1x1 48->192 gate @256x320", nn.Conv2d(48, 192, 1), B, 256, 320, gate=True)
run("1x1 96->192 + res @256x320", nn.Conv2d(96, 192, 1), B, 256, 320, res=True)
run("1x1 192->96 @256x320", nn.Conv2d(192, 96, 1), B, 256, 320)
run("3x3 96->96 @256x320", nn.Conv2d(96, 96, 3, 1, 1), B, 256, 320)
run("3x3 192->192 @256x320", nn.Conv2d(192, 192, 3, 1, 1), B, 256, 320)
run("5x5s2 384->192 @256x320", nn.Conv2d(384, 192, 5, 2, 2), B, 256, 320)
run("5x5 224->128 @32x40", nn.Conv2d(224, 128, 5, 1, 2), B, 32, 40)
run("5x5 512->384 @32x40", nn.Conv2d(512, 384, 5, 1, 2), B, 32, 40)
run("1x1 2816->469 @32x40", nn.Conv2d(2816, 469, 1), B, 32, 40)
run("deconv5x5s2 192->192 @128x160", nn.ConvTranspose2d(192, 192, 5, 2, 2, 1), B, 128, 160)
