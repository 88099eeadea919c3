"""Per-op timing of one encoder + decoder program run (CUDA events per op)."""
import sys, os, ctypes as C, collections, re
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, rgbd_b200
from gpu_utils import make_model
from rgbd_b200.synthetic import synthetic_pairs, pad_to_multiple
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
TOP = int(sys.argv[2]) if len(sys.argv) > 2 else 45
net, sd = make_model(rgbd_b200.ELIC_united, "realistic", 0, precision="bf16")
rgb, depth = synthetic_pairs(B, 480, 640, seed=1)
rgb, depth = pad_to_multiple(rgb).cuda(), pad_to_multiple(depth).cuda()
for _ in range(2):
    c = net.compress(rgb, depth); r = net.decompress(c["r_strings"], c["d_strings"], c["shape"])
torch.cuda.synchronize()
rows = []
for kind, prog in (("enc", net._program("encoder", B, 512, 640)), ("dec", net._program("decoder", B, 8, 10))):
    sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    evs = []
    for op in prog.ops:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); op(sp); b.record(); evs.append((a, b))
    torch.cuda.synchronize()
    for op, (a, b) in zip(prog.ops, evs):
        rows.append((kind, getattr(op, "label", "other"), getattr(op, "flops", 0), a.elapsed_time(b)))
tot = sum(r[3] for r in rows)
nor = sum(r[3] for r in rows if "rans" not in r[1])
print(f"B={B} total op time {tot:.1f} ms ({tot/B:.2f} ms/pair); without rANS {nor:.1f} ms ({nor/B:.2f} ms/pair)")
# category summary: kernel size / kind and resolution
cat = collections.defaultdict(lambda: [0, 0.0, 0.0])
for kind, label, fl, ms in rows:
    m = re.match(r"(tc|simt) (\S+) (\d+)->(\d+) @(\d+)x(\d+)", label)
    if m:
        key = f"{m.group(1)} {m.group(2)} @{m.group(5)}x{m.group(6)}"
    else:
        key = label
    a = cat[key]; a[0] += 1; a[1] += fl; a[2] += ms
print("-- by kernel shape and resolution")
for k, (n, fl, ms) in sorted(cat.items(), key=lambda kv: -kv[1][2]):
    if "rans" in k: continue
    print(f"{k:44s} {n:4d} {ms:9.2f} ms {100*ms/nor:5.1f}% {fl/1e9:9.1f} GF {(fl/ms/1e9 if ms else 0):8.1f} TF/s")
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for kind, label, fl, ms in rows:
    a = agg[(kind, label)]; a[0] += 1; a[1] += fl; a[2] += ms
print(f"{'op':60s} {'n':>4s} {'ms':>9s} {'share':>6s} {'GFLOP':>9s} {'TFLOP/s':>8s}")
for k, (n, fl, ms) in sorted(agg.items(), key=lambda kv: -kv[1][2])[:TOP]:
    print(f"{k[0]+' '+k[1]:60s} {n:4d} {ms:9.2f} {100*ms/tot:5.1f}% {fl/1e9:9.1f} {(fl/ms/1e9 if ms else 0):8.1f}")
