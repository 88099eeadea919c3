"""csrc/metrics.cu + rgbd_b200.metrics against oracle/msssim_oracle.py: PSNR and MS-SSIM of utils/metrics.py:8-14 within
fp32 rounding (1e-5 relative on MS-SSIM, 1e-4 dB on PSNR), the 8-bit / 16-bit exports bit for bit."""
import numpy as np
import pytest
import torch

import rgbd_b200
from oracle import msssim_oracle as M
from rgbd_b200 import metrics

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("shape", [(1, 3, 480, 640), (2, 1, 177, 203), (1, 3, 161, 161)])
def test_compute_metrics_matches_oracle(shape):
    g = torch.Generator().manual_seed(sum(shape))
    a = torch.rand(shape, generator=g)
    b = (a + 0.08 * torch.randn(shape, generator=g)) * 1.1 - 0.05        # leaves [0, 1]: both sides clamp first
    p, m = metrics.compute_metrics(b.to(DEV), a.to(DEV))
    wp, wm = M.compute_metrics(b, a)
    assert abs(p - wp) < 1e-4 and abs(m - wm) < 1e-5 * wm, (p, wp, m, wm)
    per = metrics.ms_ssim_per_channel(b.to(DEV), a.to(DEV)).float()
    want = M.ms_ssim(b.clamp(0, 1), a.clamp(0, 1), size_average=False)
    assert per.shape == want.shape and torch.allclose(per, want, rtol=2e-5, atol=0)
    again = metrics.ms_ssim_per_channel(b.to(DEV), a.to(DEV)).float()
    assert torch.equal(per, again), "fixed-order reductions: same bits on every run"
    assert rgbd_b200.compute_metrics is metrics.compute_metrics


def test_small_images_are_refused_like_the_package():
    x = torch.rand(1, 1, 160, 200, device=DEV)
    with pytest.raises(ValueError):
        metrics.ms_ssim_per_channel(x, x)
    with pytest.raises(rgbd_b200.lib.RgbdError):
        metrics.compute_metrics(x.cpu(), x.cpu())


def test_exports_match_reference_arithmetic_bit_for_bit():
    g = torch.Generator().manual_seed(5)
    x = torch.rand(2, 3, 64, 96, generator=g) * 1.4 - 0.2
    x[0, 0, 0, :4] = torch.tensor([0.0, 1.0, 0.5, 254.999 / 255])
    got = metrics.export_u8(x.to(DEV)).cpu()
    assert torch.equal(got, M.export_u8(x))
    crop = metrics.export_u8(x.to(DEV), crop=(50, 70)).cpu()
    assert torch.equal(crop, M.export_u8(x)[:, :50, :70])
    d = torch.rand(2, 1, 64, 96, generator=g) * 0.7
    for scale in (10000.0, 100000.0):           # NYUv2 / SUN RGB-D (values above 0.65535 wrap, as astype(uint16) does)
        got = metrics.export_depth_u16(d.to(DEV), scale, crop=(60, 90)).cpu().numpy()
        want = M.export_depth_u16(d, scale)[:, :60, :90]
        assert got.dtype == np.uint16 and np.array_equal(got, want), scale
