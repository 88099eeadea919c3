"""Single launches of the fused bottleneck kernel (csrc/conv_rb.cu) and of the memory-bound 1x1 layers around it,
timed with CUDA events (warm) — or one launch each under ncu (RGBD_NCU=1)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, torch.nn as nn
DEV = torch.device("cuda:0")
NCU = os.environ.get("RGBD_NCU") == "1"
dbg = torch.zeros(32, dtype=torch.int64, device=DEV)
if not NCU and os.environ.get("RGBD_BUILD_DEFINES", "").find("RGBD_TIMING_PROBES") >= 0:
    os.environ["RGBD_TC_TRACE"] = str(dbg.data_ptr())      # cycle counters of CTA 0 (development build of the library)
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
    dbg.zero_()
    timeit(b, f"fused 1x1-3x3-1x1 {cin}->96->96->192 @{H}x{W} N={N}")
    d = dbg.cpu().tolist()
    if d[10]:
        n = d[10]
        print(f"   CTA0 {n} tiles | MMA warp 2: {d[9]/n:.0f} cyc/tile; waits: w1 {d[2]/n:.0f} x {d[3]/n:.0f} d3_empty+t1_ready {d[4]/n:.0f} w2 {d[5]/n:.0f} "
              f"t2_ready {d[6]/n:.0f} d3_empty(P3) {d[7]/n:.0f} w3 {d[8]/n:.0f} | producers: x_empty {d[0]/n:.0f} w_empty {d[1]/n:.0f} | "
              f"epilogue warp 4: {d[17]/n:.0f} cyc/tile; wait d1 {d[11]/n:.0f} Ep1 {d[12]/n:.0f} wait d2 {d[13]/n:.0f} Ep2 {d[14]/n:.0f} wait d3 {d[15]/n:.0f} Ep3 {d[16]/n:.0f} [in Ep3: staging-free wait {d[19]/n:.0f} fence {d[20]/n:.0f}] residual regs+issue {d[18]/n:.0f} tile setup {d[21]/n:.0f}", flush=True)

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
