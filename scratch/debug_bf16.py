import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, rgbd_b200
from rgbd_b200.synthetic import synthetic_pairs
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from gpu_utils import make_model
DEV = "cuda:0"
for tc in (True, False):
    net, sd = make_model(rgbd_b200.ELIC_united, "mid", 0, precision="bf16", tensor_cores=tc)
    rgb, depth = synthetic_pairs(3, 128, 192, seed=77)
    res = {}
    for tag, sl in (("b3", slice(0, 3)), ("b1", slice(0, 1)), ("b3again", slice(0, 3))):
        out = net.compress(rgb[sl].to(DEV), depth[sl].to(DEV))
        B = sl.stop - sl.start
        p = net._program("encoder", B, 128, 192)
        torch.cuda.synchronize()
        res[tag] = dict(y_r=p.io["y"]["r"].torch()[0].clone(), y_d=p.io["y"]["d"].torch()[0].clone(),
                        z_r=p.io["z"]["r"].torch()[0].clone(), yhat_r=p.io["yhat"]["r"].torch()[0].clone(),
                        yhat_d=p.io["yhat"]["d"].torch()[0].clone(),
                        sym_r=p.io["st"]["r"]["ysym"][0].clone(), zsym=p.io["st"]["r"]["zsym"][0].clone(),
                        bytes=(out["r_strings"][0][0], out["r_strings"][1][0]))
    for a, b in (("b3", "b1"), ("b3", "b3again")):
        print("tc", tc, a, "vs", b, {k: (bool(torch.equal(res[a][k], res[b][k])) if k != "bytes" else res[a][k] == res[b][k]) for k in res[a]})
        for k in ("y_r", "y_d", "z_r"):
            d = (res[a][k].float() - res[b][k].float()).abs()
            print("    ", k, "maxdiff", float(d.max()), "n diff", int((d > 0).sum()), "of", d.numel())
