"""ORACLE (test infrastructure): restatement of the reference's table construction, `update(force=True)`.

Turns a state_dict whose CDF buffers are empty (or stale) into one whose four entropy models carry
freshly built tables, without touching the product library: the pmf -> cdf step goes through the
reference's own compiled `_CXX.pmf_to_quantized_cdf` (oracle/_ref) when it was built, else through the
plain-C restatement (oracle/rans_oracle.c).  Used by bench.py's CPU arm and by tests that compare the
product's `update()` with it.

What it follows (paths relative to the reference root):
  GaussianConditional.update          CompressAI/compressai/entropy_models/entropy_models.py:511-532
  EntropyBottleneck.update            same file :320-360   (logits chain :369-389)
  _pmf_to_cdf                         same file :166-172
  get_scale_table                     utils/moduleFunc.py:11-12
  models/elic_united.py:580-586       which models are updated
Everything is computed on the CPU in fp32 (as tests/golden/* were).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import coder as _port
from .ref_loader import ref_ext_available


def _pmf_to_quantized_cdf():
    if ref_ext_available():
        from . import ref_coder
        return ref_coder.pmf_to_quantized_cdf
    return _port.pmf_to_quantized_cdf


def scale_table(lo=0.11, hi=256, levels=64):
    return torch.exp(torch.linspace(math.log(lo), math.log(hi), levels, device="cpu"))


def _pmf_to_cdf(pmf, tail_mass, pmf_length, max_length):
    to_cdf = _pmf_to_quantized_cdf()
    cdf = torch.zeros((len(pmf_length), max_length + 2), dtype=torch.int32, device="cpu")
    for i in range(len(pmf_length)):
        prob = torch.cat((pmf[i, : int(pmf_length[i])], tail_mass[i]), dim=0)
        q = torch.from_numpy(np.asarray(to_cdf(prob.numpy(), 16), dtype=np.int64).astype(np.int32))
        cdf[i, : q.numel()] = q
    return cdf


def gaussian_tables(table, tail_mass=1e-9):
    """-> (_quantized_cdf int32 [n, L], _offset int32 [n], _cdf_length int32 [n])"""
    from scipy.stats import norm
    table = table.detach().float().cpu()
    multiplier = -float(norm.ppf(tail_mass / 2))
    pmf_center = torch.ceil(table * multiplier).int()
    pmf_length = 2 * pmf_center + 1
    max_length = int(torch.max(pmf_length))
    samples = torch.abs(torch.arange(max_length, device="cpu").int() - pmf_center[:, None]).float()
    scale = table.unsqueeze(1)

    def std_cumulative(v):
        return 0.5 * torch.erfc(float(-(2 ** -0.5)) * v)

    upper = std_cumulative((0.5 - samples) / scale)
    lower = std_cumulative((-0.5 - samples) / scale)
    pmf = upper - lower
    tail = 2 * lower[:, :1]
    return _pmf_to_cdf(pmf, tail, pmf_length, max_length), -pmf_center, pmf_length + 2


def bottleneck_tables(sd, prefix, filters=(3, 3, 3, 3)):
    quant = sd[prefix + ".quantiles"].detach().float().cpu()

    def logits_cumulative(v):
        logits = v
        for i in range(len(filters) + 1):
            logits = torch.matmul(F.softplus(sd[f"{prefix}._matrix{i}"].detach().float().cpu()), logits)
            logits = logits + sd[f"{prefix}._bias{i}"].detach().float().cpu()
            if i < len(filters):
                f = sd[f"{prefix}._factor{i}"].detach().float().cpu()
                logits = logits + torch.tanh(f) * torch.tanh(logits)
        return logits

    medians = quant[:, 0, 1]
    minima = torch.clamp(torch.ceil(medians - quant[:, 0, 0]).int(), min=0)
    maxima = torch.clamp(torch.ceil(quant[:, 0, 2] - medians).int(), min=0)
    pmf_start = medians - minima
    pmf_length = maxima + minima + 1
    max_length = int(pmf_length.max())
    samples = torch.arange(max_length, device="cpu")[None, :] + pmf_start[:, None, None]
    lower = logits_cumulative(samples - 0.5)
    upper = logits_cumulative(samples + 0.5)
    sign = -torch.sign(lower + upper)
    pmf = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))[:, 0, :]
    tail = torch.sigmoid(lower[:, 0, :1]) + torch.sigmoid(-upper[:, 0, -1:])
    return _pmf_to_cdf(pmf, tail, pmf_length, max_length), -minima, pmf_length + 2


def updated_state_dict(sd):
    """state_dict after `net.update(force=True)` (models/elic_united.py:580-586)."""
    out = {k: v.detach().cpu().clone() for k, v in sd.items()}
    table = scale_table()
    g = gaussian_tables(table)
    for name in ("rgb_gaussian_conditional", "depth_gaussian_conditional"):
        out[name + ".scale_table"] = table.clone()
        for key, v in zip(("_quantized_cdf", "_offset", "_cdf_length"), g):
            out[f"{name}.{key}"] = v.clone()
    for name in ("rgb_entropy_bottleneck", "depth_entropy_bottleneck"):
        for key, v in zip(("_quantized_cdf", "_offset", "_cdf_length"), bottleneck_tables(out, name)):
            out[f"{name}.{key}"] = v.clone()
    return out
