"""The drop-in boundary under the reference harness's global state (SURVEY §8b "Threading/ownership"):
playground/test.py:20 makes CUDA the default tensor type before anything is constructed.  tests/harness_emulation.py
replays the harness's steps (construct, torch.load, load_state_dict, update(force=True), .to("cuda"), pad, compress,
container files, decompress, crop0, PSNR, forward) in a fresh process with that default set."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("cls_name,precision,size", [("ELIC_united", "fp32", "150x200"), ("ELIC_united", "bf16", "150x200"),
                                                     ("ELIC_united_R2D", "fp32", "150x200"), ("STF_united", "fp32", "260x300")])
def test_harness_sequence_under_cuda_default_tensor_type(tmp_path, cls_name, precision, size):
    r = subprocess.run([sys.executable, os.path.join(HERE, "harness_emulation.py"), str(tmp_path), cls_name, precision, size],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    res = json.loads(r.stdout.strip().splitlines()[-1])
    h, w = (int(v) for v in size.split("x"))
    assert res["ok"] and res["x_hat_on_cuda"]
    assert res["shape"] == [-(-h // 64), -(-w // 64)]                # e.g. 150x200 -> padded 192x256 -> z 3x4
    assert res["x_hat_shapes"] == [[1, 3, h, w], [1, 1, h, w]]
    assert res["rgb_bpp"] > 0.1 and res["depth_bpp"] > 0.1 and all(n > 100 for n in res["y_bytes"])
    assert res["rgb_psnr"] > 0 and res["depth_psnr"] > 0 and res["cost_time"] >= 0
    if min(h, w) > 160:      # the tester's metrics and exports on the GPU under the same global state
        assert res["metrics_psnr_matches"] and 0 < res["ms_ssim"] <= 1 and res["depth16_matches"] and res["u8_shape"] == [1, h, w, 3]
    print("forward().clamp == decompress(compress()):", res["fwd_equals_decompress"])
