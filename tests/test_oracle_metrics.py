"""oracle/msssim_oracle.py (the restatement the GPU metric kernels are checked against) held to properties the algorithm
must have — there is no golden vector for pytorch_msssim in the reference (the module header says "parity unpinned")."""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import msssim_oracle as M


def test_window_and_filter_against_a_direct_2d_convolution():
    win = M._window()
    assert abs(float(win.sum()) - 1) < 1e-6 and torch.equal(win, win.flip(0)) and win.numel() == 11
    x = torch.rand(2, 3, 40, 37)
    k2 = torch.outer(win, win)
    want = F.conv2d(x, k2.view(1, 1, 11, 11).repeat(3, 1, 1, 1), groups=3)
    assert torch.allclose(M._gaussian_filter(x, win), want, atol=1e-6)


def test_identity_monotonicity_and_range():
    g = torch.Generator().manual_seed(0)
    x = torch.rand(1, 3, 176, 192, generator=g)
    assert abs(float(M.ms_ssim(x, x)) - 1) < 1e-6
    vals = [float(M.ms_ssim(x, (x + s * torch.randn(x.shape, generator=g)).clamp(0, 1))) for s in (0.01, 0.05, 0.2)]
    assert 1 > vals[0] > vals[1] > vals[2] > 0
    per = M.ms_ssim(x, (x * 0.9).clamp(0, 1), size_average=False)
    assert per.shape == (1, 3)


def test_metrics_and_exports_match_the_reference_arithmetic():
    g = torch.Generator().manual_seed(1)
    a, b = torch.rand(1, 3, 176, 176, generator=g) * 1.2 - 0.1, torch.rand(1, 3, 176, 176, generator=g)
    p, m = M.compute_metrics(a, b)
    mse = float(((a.clamp(0, 1) - b) ** 2).mean())
    assert abs(p + 10 * np.log10(mse)) < 1e-9 and 0 < m < 1
    u8 = M.export_u8(a)
    assert u8.dtype == torch.uint8 and u8.shape == (1, 176, 176, 3) and int(u8.max()) == 255 and int(u8.min()) == 0
    d = torch.rand(1, 1, 8, 8, generator=g)
    assert M.export_depth_u16(d, 10000).dtype == np.uint16 and int(M.export_depth_u16(d * 0 + 0.7, 100000)[0, 0, 0]) == 70000 % 65536
