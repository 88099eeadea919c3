"""Quality metrics and image export next to the codec path, on the GPU.

Mirrors what the reference's tester does to every reconstruction (testing/tester_united.py:92-123):

    compute_metrics(a, b, max_val=1) -> (psnr, ms_ssim)        utils/metrics.py:8-14
    AverageMeter                                                utils/metrics.py:17-31
    export_u8(x, crop)        saveImg's clamp / * 255 / truncate            utils/IOutils.py:100-102
    export_depth_u16(x, scale, crop)   (x * scale).astype(uint16)           testing/tester_united.py:101-108

MS-SSIM is `pytorch_msssim.ms_ssim` (the package utils/metrics.py:5 imports; v1.0.0: 11-tap Gaussian window, sigma 1.5,
five levels, weights 0.0448 / 0.2856 / 0.3001 / 0.2363 / 0.1333, 2x2 average pooling between levels), computed by
csrc/metrics.cu — every reduction two-stage in a fixed order, so the same inputs give the same bits on every run.  No torch
arithmetic and no CPU fallback: the inputs must live on a CUDA device.
"""
import ctypes as C
import math

import torch

from . import lib as L

MS_SSIM_WEIGHTS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)
_WIN = 11


def _sp(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _prep(a, b):
    if a.shape != b.shape or a.dim() != 4:
        raise ValueError("expected two [N, C, H, W] tensors of the same shape")
    if not a.is_cuda:
        raise L.RgbdError("metrics run on a CUDA device only (no CPU fallback)")
    a = a.detach().to(torch.float32).contiguous()
    b = b.detach().to(device=a.device, dtype=torch.float32).contiguous()
    return a, b


@torch.no_grad()
def mse_per_image(a, b, clamp01=True):
    """[N] float64 tensor: mean squared error of every image (inputs clamped to [0, 1] first, utils/metrics.py:9-11)."""
    a, b = _prep(a, b)
    N = a.shape[0]
    n_per = a[0].numel()
    n_part = max(1, min(1024, n_per // 4096))
    work = torch.empty(N * n_part, dtype=torch.float64, device=a.device)
    out = torch.empty(N, dtype=torch.float64, device=a.device)
    with torch.cuda.device(a.device):
        L.call("rgbd_sq_error_sums", a.data_ptr(), b.data_ptr(), N, n_per, int(clamp01), work.data_ptr(), n_part, out.data_ptr(),
               _sp(a.device))
    return out / n_per


@torch.no_grad()
def ms_ssim_per_channel(a, b, data_range=1.0, clamp01=True):
    """[N, C] float64 tensor of pytorch_msssim.ms_ssim(..., size_average=False) values."""
    a, b = _prep(a, b)
    N, Cc, H, W = a.shape
    if min(H, W) <= (_WIN - 1) * 2 ** 4:
        raise ValueError("image side should be larger than 160 pixels for a five-level MS-SSIM (pytorch_msssim's own check)")
    planes = N * Cc
    dev = a.device
    x, y = a.view(planes, H, W), b.view(planes, H, W)
    sums = torch.empty((len(MS_SSIM_WEIGHTS), planes, 2), dtype=torch.float64, device=dev)
    counts = []
    lib = L.load()
    with torch.cuda.device(dev):
        sp = _sp(dev)
        clamp = int(clamp01)
        for lvl in range(len(MS_SSIM_WEIGHTS)):
            h, w = x.shape[1:]
            work = torch.empty(int(lib.rgbd_ssim_work_elems(planes, h, w)), dtype=torch.float64, device=dev)
            L.call("rgbd_ssim_level", x.data_ptr(), y.data_ptr(), planes, h, w, float(data_range), clamp, work.data_ptr(),
                   sums[lvl].data_ptr(), sp)
            counts.append((h - _WIN + 1) * (w - _WIN + 1))
            if lvl + 1 < len(MS_SSIM_WEIGHTS):
                ho, wo = (h + 2 * (h & 1) - 2) // 2 + 1, (w + 2 * (w & 1) - 2) // 2 + 1
                nx = torch.empty((planes, ho, wo), dtype=torch.float32, device=dev)
                ny = torch.empty((planes, ho, wo), dtype=torch.float32, device=dev)
                L.call("rgbd_avgpool2", x.data_ptr(), nx.data_ptr(), planes, h, w, clamp, sp)
                L.call("rgbd_avgpool2", y.data_ptr(), ny.data_ptr(), planes, h, w, clamp, sp)
                x, y, clamp = nx, ny, 0          # (the clamp is applied once, when the full-resolution images are read)
    # the last, tiny step on the host in float64: prod_l relu(cs_l)^w_l (l < 4) * relu(ssim_4)^w_4
    s = sums.cpu()
    val = torch.ones(planes, dtype=torch.float64, device="cpu")     # (explicit: the reference harness makes CUDA the default device)
    for lvl, wgt in enumerate(MS_SSIM_WEIGHTS):
        which = 0 if lvl + 1 == len(MS_SSIM_WEIGHTS) else 1
        mean = (s[lvl, :, which] / counts[lvl]).to(torch.float32).to(torch.float64)      # the package works in fp32
        val = val * torch.relu(mean) ** wgt
    return val.view(N, Cc)


@torch.no_grad()
def compute_metrics(a, b, max_val: float = 1):
    """(psnr, ms_ssim) exactly as utils/metrics.py:8-14: both images clamped to [0, 1], one MSE over the whole batch,
    ms_ssim averaged over batch and channels."""
    mse = float(mse_per_image(a, b).mean())
    p = 20 * math.log10(max_val) - 10 * math.log10(mse)
    m = float(ms_ssim_per_channel(a, b, data_range=max_val).mean())
    return p, m


class AverageMeter:
    """Compute running average (utils/metrics.py:17-31)."""

    def __init__(self):
        self.val = 0
        self.avg = 0
        self.sum = 0
        self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


@torch.no_grad()
def export_u8(x, crop=None):
    """saveImg's pixel arithmetic (utils/IOutils.py:100-102 + torchvision ToPILImage: clamp to [0, 1], * 255, truncate):
    [N, C, H, W] fp32 on the GPU -> [N, h, w, C] uint8 on the GPU, optionally cropped to the top-left (h, w) (crop0,
    dataset/utils.py:84-85).  PNG encoding itself is file I/O and stays with the caller."""
    if not x.is_cuda:
        raise L.RgbdError("export runs on a CUDA device only (no CPU fallback)")
    x = x.detach().to(torch.float32).contiguous()
    N, Cc, H, W = x.shape
    h, w = crop or (H, W)
    out = torch.empty((N, h, w, Cc), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        L.call("rgbd_quantize_u8", x.data_ptr(), out.data_ptr(), N, Cc, H, W, h, w, _sp(x.device))
    return out


@torch.no_grad()
def export_depth_u16(x, scale=10000.0, crop=None):
    """The 16-bit depth PNG payload of testing/tester_united.py:101-108: (x * scale).astype(uint16) — 10000 for NYUv2, 100000
    for SUN RGB-D — of a [N, 1, H, W] image -> [N, h, w] uint16-valued int16 storage (torch has no uint16 arithmetic; view
    the result with .view(torch.uint16) or numpy)."""
    if not x.is_cuda:
        raise L.RgbdError("export runs on a CUDA device only (no CPU fallback)")
    x = x.detach().to(torch.float32).contiguous()
    N, Cc, H, W = x.shape
    if Cc != 1:
        raise ValueError("depth images have one channel")
    h, w = crop or (H, W)
    out = torch.empty((N, h, w), dtype=torch.int16, device=x.device)
    with torch.cuda.device(x.device):
        L.call("rgbd_quantize_u16", x.data_ptr(), out.data_ptr(), N, H, W, h, w, float(scale), _sp(x.device))
    return out.view(torch.uint16)
