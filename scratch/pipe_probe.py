"""Are conv kernels slower inside the pipelined run (SMs fenced off by rANS blocks) than alone?"""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, rgbd_b200
from gpu_utils import make_model
from rgbd_b200.synthetic import synthetic_pairs, pad_to_multiple
from rgbd_b200.pipeline import RoundTripPipeline
S, B = int(sys.argv[1]), int(sys.argv[2])
net, sd = make_model(rgbd_b200.ELIC_united, "realistic", 0, precision="bf16")
rgb, depth = synthetic_pairs(B, 480, 640, seed=1)
rgb, depth = pad_to_multiple(rgb).cuda(), pad_to_multiple(depth).cuda()
pipe = RoundTripPipeline(net, S)
jobs = [(rgb, depth)] * (2 * S)
pipe.run(jobs)
torch.cuda.synchronize()
WATCH = ["tc conv5x5s2 384->192 @128x160", "tc conv3x3 192->192 @256x320", "tc conv1x1 96->192 @256x320", "tc conv5x5 512->384 @32x40"]
records = {w: [] for w in WATCH}
def wrap(prog):
    for i, op in enumerate(prog.ops):
        lab = getattr(op, "label", "")
        if lab in WATCH:
            def timed(sp, op=op, lab=lab):
                st = torch.cuda.current_stream()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(st); op(sp); b.record(st)
                records[lab].append((a, b))
            timed.label = lab
            prog.ops[i] = timed
for key, prog in net._programs.items():
    wrap(prog)
# alone: one program at a time
c = net.compress(rgb, depth); net.decompress(c["r_strings"], c["d_strings"], c["shape"])
torch.cuda.synchronize()
alone = {w: sum(a.elapsed_time(b) for a, b in records[w]) / max(1, len(records[w])) for w in WATCH}
for w in WATCH: records[w].clear()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); pipe.run(jobs * 2); e1.record(); torch.cuda.synchronize()
print(f"pipelined {S}x{B}: {4*S*B/(e0.elapsed_time(e1)/1e3):.1f} pairs/s")
for w in WATCH:
    d = sorted(a.elapsed_time(b) for a, b in records[w])
    print(f"{w:40s} alone {alone[w]*1e3:7.0f} us | pipelined n={len(d)} median {d[len(d)//2]*1e3:7.0f} us  p10 {d[len(d)//10]*1e3:7.0f}  p90 {d[9*len(d)//10]*1e3:7.0f}")
