"""Host-side mirror of the two CompressAI entropy models the ELIC_united path uses.

These modules hold exactly the reference's parameters / buffers (same state_dict keys) and
rebuild the CDF tables in `update()`; all per-symbol work (quantise, index, likelihood,
rANS) is done by the CUDA kernels behind the C-ABI.

Reference: CompressAI/compressai/entropy_models/entropy_models.py
  EntropyModel buffers :88-91, EntropyBottleneck :269-446, GaussianConditional :450-568.
Table construction always runs on the CPU in fp32, whatever the default tensor type is: the
reference harness sets torch.set_default_tensor_type('torch.cuda.FloatTensor') (playground/test.py:20),
so every factory call here names its device.  (The reference computes the tables on whatever
device the buffers happen to live on; the CPU is the one choice that gives the same tables on
every machine, and it is what the golden vectors were generated with.)  The pmf -> cdf step goes
through rgbd_pmf_to_quantized_cdf (C-ABI, host) instead of compressai._CXX.
"""
import ctypes as C
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import lib as _lib

_TAIL_MASS = 1e-9


def _norm_ppf(q):
    """Inverse normal CDF in double precision (reference uses scipy.stats.norm.ppf,
    entropy_models.py:497-498)."""
    try:
        from scipy.stats import norm
        return float(norm.ppf(q))
    except Exception:  # pragma: no cover - scipy is present in this image
        from statistics import NormalDist
        return NormalDist().inv_cdf(q)


class _Bound(nn.Module):
    """Parameter-less stand-in for compressai.ops.LowerBound: only its `bound` buffer is part
    of the state_dict (bound_ops.py:37-42)."""

    def __init__(self, bound):
        super().__init__()
        self.register_buffer("bound", torch.Tensor([float(bound)]))


def pmf_to_quantized_cdf(pmf, precision=16):
    """float32 pmf (1-D tensor / array) -> int32 CDF of len(pmf)+1 (ops.cpp:24-81)."""
    p = np.ascontiguousarray(np.asarray(pmf, dtype=np.float32))
    out = np.zeros(p.size + 1, dtype=np.uint32)
    _lib.call("rgbd_pmf_to_quantized_cdf", p.ctypes.data_as(C.c_void_p), p.size, precision,
              out.ctypes.data_as(C.c_void_p))
    return torch.from_numpy(out.astype(np.int32))


def encoder_records(flat_cdf, base, length):
    """Per CDF bin the encoder's exact-reciprocal record {u64 rcp, u32 bias, u32 range | shift << 16}
    (ryg_rans Rans64EncSymbolInit, rans64.h:167-247), computed with Python integers.
    Returns int64 [total, 2] (16 bytes per bin, little endian)."""
    total = int(flat_cdf.size)
    out = np.zeros((total, 2), dtype=np.uint64)
    for t in range(len(length)):
        b, n = int(base[t]), int(length[t])
        for v in range(n - 1):
            start = int(flat_cdf[b + v])
            rng = (int(flat_cdf[b + v + 1]) - start) & 0xFFFF     # uint16 cast of the reference
            if rng >= 2:
                s = (rng - 1).bit_length()                          # ceil(log2(range))
                rcp = ((1 << (s + 63)) + rng - 1) // rng
                shift, bias = s - 1, start
            else:                                                   # range 1: q = mulhi(x, 2^64-1) = x - 1
                rcp, shift, bias = (1 << 64) - 1, 0, start + 65535
            out[b + v, 0] = rcp
            out[b + v, 1] = bias | ((rng | (shift << 16)) << 32)
    return out.view(np.int64)


class DeviceTables:
    """Compacted uint16 CDF tables + (base, length, offset) resident on the device."""

    def __init__(self, quantized_cdf, cdf_length, offset, device):
        cdf = quantized_cdf.detach().cpu().numpy().astype(np.int64)
        length = cdf_length.detach().cpu().numpy().astype(np.int32).reshape(-1)
        off = offset.detach().cpu().numpy().astype(np.int32).reshape(-1)
        base = np.zeros_like(length)
        base[1:] = np.cumsum(length[:-1])
        flat = np.concatenate([cdf[i, : length[i]] for i in range(len(length))]) if len(length) else np.zeros(0)
        flat16 = (flat & 0xFFFF).astype(np.uint16)  # 65536 -> 0, like the reference's uint16 casts
        self.cdf = torch.from_numpy(flat16.view(np.int16)).to(device)
        self.base = torch.from_numpy(base).to(device)
        self.length = torch.from_numpy(length).to(device)
        self.offset = torch.from_numpy(off).to(device)
        self.enc_rec = torch.from_numpy(encoder_records(flat, base, length)).to(device)
        self.struct = _lib.RansTables(self.cdf.data_ptr(), self.base.data_ptr(), self.length.data_ptr(),
                                      self.offset.data_ptr(), int(len(length)), int(flat16.size),
                                      self.enc_rec.data_ptr())


class EntropyModelBase(nn.Module):
    def __init__(self, likelihood_bound=1e-9):
        super().__init__()
        self.likelihood_bound = float(likelihood_bound)
        self.likelihood_lower_bound = _Bound(likelihood_bound)
        self.register_buffer("_offset", torch.IntTensor())
        self.register_buffer("_quantized_cdf", torch.IntTensor())
        self.register_buffer("_cdf_length", torch.IntTensor())
        self._dev_tables = None

    @property
    def offset(self):
        return self._offset

    @property
    def quantized_cdf(self):
        return self._quantized_cdf

    @property
    def cdf_length(self):
        return self._cdf_length

    def _pmf_to_cdf(self, pmf, tail_mass, pmf_length, max_length):
        cdf = torch.zeros((len(pmf_length), max_length + 2), dtype=torch.int32, device="cpu")
        for i in range(len(pmf_length)):
            prob = torch.cat((pmf[i, : int(pmf_length[i])], tail_mass[i]), dim=0)
            q = pmf_to_quantized_cdf(prob, 16)
            cdf[i, : q.numel()] = q
        return cdf

    def device_tables(self, device):
        if self._quantized_cdf.numel() == 0:
            raise ValueError("Uninitialized CDFs. Run update() first")
        t = self._dev_tables
        if t is None or t.cdf.device != torch.device(device):
            self._dev_tables = DeviceTables(self._quantized_cdf, self._cdf_length, self._offset, device)
        return self._dev_tables

    def invalidate(self):
        self._dev_tables = None


class EntropyBottleneck(EntropyModelBase):
    """Factorised prior on z (entropy_models.py:269-446): parameter holder + update()."""

    def __init__(self, channels, tail_mass=1e-9, init_scale=10, filters=(3, 3, 3, 3)):
        super().__init__()
        self.channels = int(channels)
        self.filters = tuple(int(f) for f in filters)
        self.init_scale = float(init_scale)
        self.tail_mass = float(tail_mass)
        widths = (1,) + self.filters + (1,)
        scale = self.init_scale ** (1 / (len(self.filters) + 1))
        for i in range(len(self.filters) + 1):
            init = np.log(np.expm1(1 / scale / widths[i + 1]))
            self.register_parameter(f"_matrix{i:d}", nn.Parameter(torch.full((channels, widths[i + 1], widths[i]), float(init))))
            self.register_parameter(f"_bias{i:d}", nn.Parameter(torch.empty(channels, widths[i + 1], 1).uniform_(-0.5, 0.5)))
            if i < len(self.filters):
                self.register_parameter(f"_factor{i:d}", nn.Parameter(torch.zeros(channels, widths[i + 1], 1)))
        q = torch.Tensor([-self.init_scale, 0, self.init_scale])
        self.quantiles = nn.Parameter(q.repeat(channels, 1, 1))
        target = np.log(2 / self.tail_mass - 1)
        self.register_buffer("target", torch.Tensor([-target, 0, target]))

    def medians(self):
        return self.quantiles[:, 0, 1].detach()

    def _logits_cumulative(self, v):
        # v: [C, 1, L] on the CPU (entropy_models.py:369-389)
        logits = v
        for i in range(len(self.filters) + 1):
            m = getattr(self, f"_matrix{i:d}").detach().float().cpu()
            logits = torch.matmul(F.softplus(m), logits)
            logits = logits + getattr(self, f"_bias{i:d}").detach().float().cpu()
            if i < len(self.filters):
                f = getattr(self, f"_factor{i:d}").detach().float().cpu()
                logits = logits + torch.tanh(f) * torch.tanh(logits)
        return logits

    @torch.no_grad()
    def update(self, force=False):
        # entropy_models.py:320-360 (the early-return is commented out in the reference)
        quant = self.quantiles.detach().float().cpu()
        medians = quant[:, 0, 1]
        minima = torch.clamp(torch.ceil(medians - quant[:, 0, 0]).int(), min=0)
        maxima = torch.clamp(torch.ceil(quant[:, 0, 2] - medians).int(), min=0)
        pmf_start = medians - minima
        pmf_length = maxima + minima + 1
        max_length = int(pmf_length.max())
        samples = torch.arange(max_length, device="cpu")[None, :] + pmf_start[:, None, None]
        lower = self._logits_cumulative(samples - 0.5)
        upper = self._logits_cumulative(samples + 0.5)
        sign = -torch.sign(lower + upper)
        pmf = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))[:, 0, :]
        tail = torch.sigmoid(lower[:, 0, :1]) + torch.sigmoid(-upper[:, 0, -1:])
        dev = self.quantiles.device
        self._quantized_cdf = self._pmf_to_cdf(pmf, tail, pmf_length, max_length).to(dev)
        self._offset = (-minima).to(dev)
        self._cdf_length = (pmf_length + 2).to(dev)
        self.invalidate()
        return True

    @torch.no_grad()
    def packed_params(self, device):
        """[C][59] floats for rgbd_eb_likelihood: softplus(matrix), bias, tanh(factor) per layer,
        then the median."""
        cols = []
        n = len(self.filters)
        for i in range(n + 1):
            m = F.softplus(getattr(self, f"_matrix{i:d}").detach().float().cpu())
            cols.append(m.reshape(self.channels, -1))
            cols.append(getattr(self, f"_bias{i:d}").detach().float().cpu().reshape(self.channels, -1))
            if i < n:
                cols.append(torch.tanh(getattr(self, f"_factor{i:d}").detach().float().cpu()).reshape(self.channels, -1))
        cols.append(self.quantiles.detach().float().cpu()[:, 0, 1:2])
        p = torch.cat(cols, dim=1).contiguous()
        assert p.shape[1] == 59, p.shape
        return p.to(device)


def get_scale_table(lo=0.11, hi=256, levels=64):
    """utils/moduleFunc.py:11-12, always evaluated on the CPU: the table (and with it every CDF) must not depend on
    which device happens to be the default (the reference harness makes CUDA the default; exp differs in the
    last ulp between devices)."""
    return torch.exp(torch.linspace(math.log(lo), math.log(hi), levels, device="cpu"))


class GaussianConditional(EntropyModelBase):
    """Gaussian conditional (entropy_models.py:450-568): buffer holder + table update()."""

    def __init__(self, scale_table=None, scale_bound=0.11, tail_mass=1e-9):
        super().__init__()
        self.tail_mass = float(tail_mass)
        self.lower_bound_scale = _Bound(scale_bound)
        self.register_buffer("scale_table", torch.Tensor())
        self.register_buffer("scale_bound", torch.Tensor([float(scale_bound)]))

    def update_scale_table(self, scale_table, force=False):
        if self._offset.numel() > 0 and not force:
            return False
        dev = self.scale_table.device
        if isinstance(scale_table, torch.Tensor):
            vals = scale_table.detach().to(device="cpu", dtype=torch.float32).reshape(-1)
        else:
            vals = torch.tensor([float(v) for v in scale_table], dtype=torch.float32, device="cpu")
        self.scale_table = vals.clone().to(dev)
        self.update()
        return True

    @staticmethod
    def _std_cumulative(v):
        return 0.5 * torch.erfc(float(-(2 ** -0.5)) * v)

    @torch.no_grad()
    def update(self):
        # entropy_models.py:511-532
        table = self.scale_table.detach().float().cpu()
        multiplier = -_norm_ppf(self.tail_mass / 2)
        pmf_center = torch.ceil(table * multiplier).int()
        pmf_length = 2 * pmf_center + 1
        max_length = int(torch.max(pmf_length))
        samples = torch.abs(torch.arange(max_length, device="cpu").int() - pmf_center[:, None]).float()
        scale = table.unsqueeze(1)
        upper = self._std_cumulative((0.5 - samples) / scale)
        lower = self._std_cumulative((-0.5 - samples) / scale)
        pmf = upper - lower
        tail = 2 * lower[:, :1]
        dev = self.scale_table.device
        self._quantized_cdf = self._pmf_to_cdf(pmf, tail, pmf_length, max_length).to(dev)
        self._offset = (-pmf_center).to(dev)
        self._cdf_length = (pmf_length + 2).to(dev)
        self.invalidate()
