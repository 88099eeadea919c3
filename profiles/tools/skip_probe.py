"""Timing experiment: pipelined throughput with the rANS kernels skipped (after one real warm-up step so every buffer
holds plausible data) = what the conv pipeline alone sustains; the difference to the real run is what the coder's
serial chains / SM fencing cost.
Needs a library built with the probes: RGBD_BUILD_DEFINES=-DRGBD_TIMING_PROBES python <pkg>/build.py --force"""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, rgbd_b200
from gpu_utils import make_model
from rgbd_b200.synthetic import synthetic_pairs, pad_to_multiple
from rgbd_b200.pipeline import RoundTripPipeline
S, B = int(sys.argv[1]), int(sys.argv[2])
net, sd = make_model(rgbd_b200.ELIC_united, "realistic", 0, precision="bf16")
net.use_cuda_graph = False
rgb, depth = synthetic_pairs(B, 480, 640, seed=1)
rgb, depth = pad_to_multiple(rgb).cuda(), pad_to_multiple(depth).cuda()
pipe = RoundTripPipeline(net, S)
jobs = [(rgb, depth)] * (2 * S)
pipe.run(jobs); torch.cuda.synchronize()
for skip in (0, 2, 1):
    if skip: os.environ["RGBD_RANS_SKIP"] = str(skip)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); pipe.run(jobs * 2); e1.record(); torch.cuda.synchronize()
    print(f"{['real        ', 'rANS skipped', 'rANS -> sleep of the same length'][skip]} {S}x{B} eager: {4*S*B/(e0.elapsed_time(e1)/1e3):.1f} pairs/s")
