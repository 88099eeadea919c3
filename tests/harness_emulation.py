"""What the reference harness does to a model class, step by step, without the reference tree (which does not exist
on the GPU box).  Run as a script by tests/test_gpu_dropin_harness.py in a process of its own, because step 1 changes
torch's global default tensor type.

  playground/test.py:20             torch.set_default_tensor_type('torch.cuda.FloatTensor')
  testing/tester.py:55-59           net = modelZoo[...](config=model_config, channel=4).eval()
  testing/tester.py:100-108         ckpt = torch.load(path); net.load_state_dict(ckpt["state_dict"]);
                                    net.update(force=True); net = net.to("cuda")
  testing/tester_united.py:52-60    rgb.to("cuda"), pad(rgb, "replicate0") (dataset/utils.py:58-67,103-110)
  testing/tester_united.py:141-167  compress, write_uints + write_body per modality, bpp from the file size
  testing/tester_united.py:169-195  read_uints + read_body, decompress(rgb_strings, depth_strings, shape), crop0
  utils/metrics.py:8-14             PSNR of the cropped reconstruction
"""
import json
import math
import os
import sys
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(workdir, cls_name, precision, size="150x200"):
    import torch
    import torch.nn.functional as F
    import rgbd_b200
    from rgbd_b200 import bitstream_io as bio
    from rgbd_b200.synthetic import synthetic_pairs, synthetic_state_dict

    # a checkpoint file as a training run would have left it (saved after update(), tensors on the CPU)
    cls = getattr(rgbd_b200, cls_name)
    seed_net = cls(config=rgbd_b200.model_config(), channel=4).eval()
    seed_net.load_state_dict(synthetic_state_dict(seed_net, 0, "mid"))
    seed_net.update(force=True)
    ckpt_path = os.path.join(workdir, "checkpoint_best_loss.pth.tar")
    torch.save({"state_dict": seed_net.state_dict(), "epoch": 7}, ckpt_path)
    h0, w0 = (int(v) for v in size.split("x"))
    rgb_cpu, depth_cpu = synthetic_pairs(1, h0, w0, seed=11)        # not a multiple of 64: exercises pad / crop0
    del seed_net

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        torch.set_default_tensor_type('torch.cuda.FloatTensor')      # playground/test.py:20
    assert torch.zeros(1).is_cuda                                    # every device-less factory call now lands on the GPU

    net = cls(config=rgbd_b200.model_config(), channel=4).eval()     # tester.py:55-59 (parameters are born on the GPU)
    net.precision = precision
    checkpoint = torch.load(ckpt_path)                               # tester.py:103
    net.load_state_dict(checkpoint["state_dict"])
    net.update(force=True)
    net = net.to("cuda")
    assert checkpoint["epoch"] == 7

    B, C, H, W = rgb_cpu.shape
    rgb, depth = rgb_cpu.to("cuda"), depth_cpu.to("cuda")

    def pad0(x, p=64, mode="replicate"):
        h, w = x.size(2), x.size(3)
        ph = p * (h // p + 1) - h if h % p else 0
        pw = p * (w // p + 1) - w if w % p else 0
        return F.pad(x, (0, pw, 0, ph), mode=mode, value=0)

    rgb_pad, depth_pad = pad0(rgb), pad0(depth)
    torch.cuda.synchronize()
    with torch.no_grad():
        out = net.compress(rgb_pad, depth_pad)
    torch.cuda.synchronize()
    rp, dp = os.path.join(workdir, "depth_bin", "img0"), os.path.join(workdir, "rgb_bin", "img0")   # (sic, :62-63)
    rgb_bpp, depth_bpp = bio.save_compressed(out, (H, W), rp, dp)

    rgb_strings, depth_strings, shape, original_size = bio.load_compressed(rp, dp)
    torch.cuda.synchronize()
    with torch.no_grad():
        rec = net.decompress(rgb_strings, depth_strings, shape)
    torch.cuda.synchronize()
    rgb_x_hat = rec["x_hat"]["r"][:, :, :original_size[0], :original_size[1]]       # crop0
    depth_x_hat = rec["x_hat"]["d"][:, :, :original_size[0], :original_size[1]]

    def psnr(a, b):
        mse = torch.mean((a.clamp(0, 1).cuda() - b.clamp(0, 1).cuda()) ** 2).item()
        return -10 * math.log10(mse)

    # forward() under the same global state
    fwd = net(rgb_pad, depth_pad)
    depth16 = (depth_x_hat * 10000).cpu().squeeze().numpy().astype("uint16")         # tester_united.py:101-108
    extra = {}
    if min(H, W) > 160:      # utils/metrics.py:8-14 + the exports on the GPU, under the same global state (MS-SSIM needs > 160 px)
        import numpy as np
        p_gpu, m_gpu = rgbd_b200.compute_metrics(rgb_x_hat, rgb)
        d16 = rgbd_b200.metrics.export_depth_u16(depth_x_hat.contiguous(), 10000.0).cpu().numpy()
        u8 = rgbd_b200.metrics.export_u8(rgb_x_hat.contiguous())
        extra = {"metrics_psnr_matches": abs(p_gpu - psnr(rgb_x_hat, rgb)) < 1e-3, "ms_ssim": m_gpu,
                 "depth16_matches": bool(np.array_equal(d16[0], depth16)), "u8_shape": list(u8.shape)}
    print(json.dumps({**extra,
        "ok": True, "rgb_bpp": rgb_bpp, "depth_bpp": depth_bpp, "shape": list(shape),
        "rgb_psnr": psnr(rgb_x_hat, rgb), "depth_psnr": psnr(depth_x_hat, depth),
        "x_hat_shapes": [list(rgb_x_hat.shape), list(depth_x_hat.shape)],
        "x_hat_on_cuda": bool(rgb_x_hat.is_cuda), "cost_time": rec["cost_time"],
        "fwd_equals_decompress": bool(torch.equal(fwd["x_hat"]["r"].clamp(0, 1), rec["x_hat"]["r"])),
        "depth16_max": int(depth16.max()),
        "y_bytes": [len(out["r_strings"][0][0]), len(out["d_strings"][0][0])],
    }))


if __name__ == "__main__":
    main(*sys.argv[1:5])
