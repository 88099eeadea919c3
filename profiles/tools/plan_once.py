"""One compress + decompress of B pairs through freshly built launch plans, eager launches, inside a
cudaProfilerStart/Stop window — the target of the ncu captures in profiles/ (launch list, DRAM traffic, --set full):

    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
        --clock-control none --csv --log-file gpurun_out/plan.csv python profiles/tools/plan_once.py ELIC_united 480 640 8 bf16
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import rgbd_b200  # noqa: E402
from rgbd_b200.synthetic import pad_to_multiple, synthetic_pairs, synthetic_state_dict  # noqa: E402


def main():
    model, H, W, B, precision = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
    preset = sys.argv[6] if len(sys.argv) > 6 else "realistic"
    net = getattr(rgbd_b200, model)(config=rgbd_b200.model_config(), channel=4, precision=precision).eval()
    net.load_state_dict(synthetic_state_dict(net, 0, preset))
    net.update(force=True)
    net = net.to("cuda:0")
    rgb, depth = synthetic_pairs(B, H, W, seed=1234)
    rgb, depth = pad_to_multiple(rgb).cuda(), pad_to_multiple(depth).cuda()
    for it in range(2):
        if it == 1:
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
        c = net.compress(rgb, depth)
        r = net.decompress(c["r_strings"], c["d_strings"], c["shape"])
        torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    enc = net._program("encoder", B, rgb.shape[2], rgb.shape[3])
    dec = net._program("decoder", B, rgb.shape[2] // 64, rgb.shape[3] // 64)
    conv_bytes = sum(getattr(op, "bytes", 0) for p in (enc, dec) for op in p.ops if getattr(op, "is_conv", False))
    print("PLAN", model, H, W, B, precision, "conv_launches", sum(1 for p in (enc, dec) for op in p.ops if getattr(op, "is_conv", False)),
          "algorithmic_conv_bytes", conv_bytes, "gflop", (enc.flops + dec.flops) / 1e9, "bytes_out",
          sum(len(s) for k in ("r_strings", "d_strings") for g in c[k] for s in g), float(r["x_hat"]["r"].mean()))


if __name__ == "__main__":
    main()
