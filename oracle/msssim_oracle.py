"""TEST INFRASTRUCTURE — torch-CPU restatement of `pytorch_msssim.ms_ssim` and of the reference's metric / export
arithmetic (utils/metrics.py:8-14, utils/IOutils.py:100-102, testing/tester_united.py:101-108).

PARITY UNPINNED for MS-SSIM: the algorithm lives in the third-party package `pytorch_msssim` (imported by
utils/metrics.py:5; the reference's requirements pin no version, 1.0.0 is the current release), which is neither vendored
under /root/reference nor installed in this image, so there is no golden vector to pin this restatement against.  It
follows the package's published algorithm: `_fspecial_gauss_1d(11, 1.5)`; `gaussian_filter` = the 1-D window applied by
`conv2d` along H, then along W, no padding; `_ssim` with K = (0.01, 0.03); five levels with
`avg_pool2d(kernel_size=2, padding=side % 2)` in between; relu on cs / ssim; product of the level terms raised to the
weights (0.0448, 0.2856, 0.3001, 0.2363, 0.1333); mean over batch and channels.  PSNR and the integer exports are plain
torch and are what the CUDA kernels are compared with bit for bit.
"""
import numpy as np
import torch
import torch.nn.functional as F

WEIGHTS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)


def _window(size=11, sigma=1.5):
    coords = torch.arange(size, dtype=torch.float32) - size // 2
    g = torch.exp(-(coords ** 2) / (2 * sigma ** 2))
    return g / g.sum()


def _gaussian_filter(x, win):
    C = x.shape[1]
    out = F.conv2d(x, win.view(1, 1, -1, 1).repeat(C, 1, 1, 1), groups=C)      # along H
    return F.conv2d(out, win.view(1, 1, 1, -1).repeat(C, 1, 1, 1), groups=C)   # along W


def _ssim(X, Y, data_range, win, K=(0.01, 0.03)):
    C1, C2 = (K[0] * data_range) ** 2, (K[1] * data_range) ** 2
    mu1, mu2 = _gaussian_filter(X, win), _gaussian_filter(Y, win)
    mu1_sq, mu2_sq, mu1_mu2 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    s1 = _gaussian_filter(X * X, win) - mu1_sq
    s2 = _gaussian_filter(Y * Y, win) - mu2_sq
    s12 = _gaussian_filter(X * Y, win) - mu1_mu2
    cs_map = (2 * s12 + C2) / (s1 + s2 + C2)
    ssim_map = ((2 * mu1_mu2 + C1) / (mu1_sq + mu2_sq + C1)) * cs_map
    return ssim_map.flatten(2).mean(-1), cs_map.flatten(2).mean(-1)


def ms_ssim(X, Y, data_range=1.0, size_average=True):
    X, Y = X.float(), Y.float()
    assert min(X.shape[-2:]) > (11 - 1) * 2 ** 4, "image side should be larger than 160"
    win = _window()
    mcs = []
    for i in range(len(WEIGHTS)):
        ssim_pc, cs = _ssim(X, Y, data_range, win)
        if i < len(WEIGHTS) - 1:
            mcs.append(torch.relu(cs))
            pad = [s % 2 for s in X.shape[2:]]
            X, Y = F.avg_pool2d(X, kernel_size=2, padding=pad), F.avg_pool2d(Y, kernel_size=2, padding=pad)
    vals = torch.stack(mcs + [torch.relu(ssim_pc)], dim=0)
    w = torch.tensor(WEIGHTS, dtype=vals.dtype).view(-1, 1, 1)
    out = torch.prod(vals ** w, dim=0)
    return out.mean() if size_average else out


def compute_metrics(a, b, max_val=1.0):
    """utils/metrics.py:8-14 on the CPU."""
    a, b = a.clamp(0, 1), b.clamp(0, 1)
    mse = torch.mean((a - b) ** 2).item()
    return 20 * np.log10(max_val) - 10 * np.log10(mse), ms_ssim(a, b, data_range=max_val).item()


def export_u8(x):
    """ToPILImage of saveImg: clamp, mul(255), byte() (truncation); [N, C, H, W] -> [N, H, W, C] uint8."""
    return x.clamp(0, 1).mul(255).byte().permute(0, 2, 3, 1).contiguous()


def export_depth_u16(x, scale):
    """(depth_x_hat * scale).cpu().squeeze().numpy().astype("uint16") of testing/tester_united.py:101-105."""
    return (x * scale).squeeze(1).numpy().astype("uint16")
