"""Deterministic synthetic weights and NYUv2-shaped inputs (there is no network for real
checkpoints or datasets).

`synthetic_state_dict` fills every tensor of a model's state_dict from per-key seeded
generators, so the same "checkpoint" can be rebuilt anywhere (build container, GPU box) and
loaded into the reference model, the oracle and the CUDA model alike.  Random-init weights
give a degenerate codec (all scales clamp to 0.11, the y stream is 8 bytes: SURVEY F4), so the
generator is *calibrated*: variance-preserving conv weights keep |y| at a few units and the
scale half of every EntropyParametersEX output bias is drawn log-uniformly, which yields
realistic rates (a few bits per symbol down to ~0.2) and exercises many CDF tables.
"""
import hashlib
import math

import torch

# With the variance-preserving init below the un-calibrated latent has std ~13 (measured once with
# the reference g_a on synthetic_pairs); presets are quoted as the latent std they aim for.
_Y_STD_UNCALIBRATED = 13.0
PRESETS = {
    # name: (target latent std, (lo, hi) of the log-uniform scale bias)
    "realistic": (0.45, (0.11, 0.7)),
    "mid": (1.2, (0.11, 2.0)),
    "stress": (12.0, (0.15, 8.0)),
}


def _gen(key, seed):
    h = hashlib.sha256(f"{seed}:{key}".encode()).digest()
    g = torch.Generator(device="cpu")
    g.manual_seed(int.from_bytes(h[:7], "little"))
    return g


def synthetic_state_dict(model, seed=0, preset="mid"):
    """Returns a new state_dict for `model` (any object with the reference's state_dict keys)."""
    target, (lo, hi) = PRESETS[preset]
    gain = target / _Y_STD_UNCALIBRATED
    ref = model.state_dict()
    out = {}
    for key, t in ref.items():
        g = _gen(key, seed)
        shape = tuple(t.shape)
        leaf = key.rsplit(".", 1)[-1]
        if "entropy_bottleneck" in key or "gaussian_conditional" in key:
            if leaf.startswith("_matrix"):
                widths = (1, 3, 3, 3, 3, 1)
                i = int(leaf[-1])
                scale = 10 ** (1 / 5)
                v = torch.full(shape, math.log(math.expm1(1 / scale / widths[i + 1])), device="cpu")
            elif leaf.startswith("_bias"):
                v = torch.rand(shape, generator=g, device="cpu") - 0.5
            elif leaf.startswith("_factor"):
                v = torch.zeros(shape, device="cpu")
            elif leaf == "quantiles":
                v = torch.tensor([-10.0, 0.0, 10.0], device="cpu").repeat(shape[0], 1, 1)
                v[:, 0, 1] = (torch.rand(shape[0], generator=g, device="cpu") - 0.5) * 2.0   # non-trivial medians
                v[:, 0, 0] += v[:, 0, 1]
                v[:, 0, 2] += v[:, 0, 1]
            else:
                v = t.clone()       # constants (target, bounds) and the empty table buffers
            out[key] = v.to(t.dtype)
            continue
        if leaf == "weight" and len(shape) == 4:
            fan_in = shape[1] * shape[2] * shape[3]
            if ".deconv." in key or _is_transposed(model, key):
                # ConvTranspose2d weight is [Cin, Cout, k, k]; a stride-2 5x5 kernel touches
                # each output through ~k*k/4 taps
                k = shape[2]
                fan_in = shape[0] * (k * k / 4.0 if k == 5 else k * k)
            std = math.sqrt(2.0 / fan_in)
            v = torch.randn(shape, generator=g, device="cpu") * std
            if ".branch.4." in key or ".conv.4." in key:
                v = v * 0.5     # residual branches: keep the sum's variance from growing too fast
        elif leaf == "weight" and len(shape) == 2:
            v = torch.randn(shape, generator=g, device="cpu") * math.sqrt(1.0 / shape[1])
        elif leaf == "weight" and len(shape) == 1:
            # LayerNorm gains (the Swin blocks of STF_united): near one, deterministic
            v = 1.0 + (torch.rand(shape, generator=g, device="cpu") - 0.5) * 0.2
        elif leaf == "relative_position_bias_table":
            # the reference initialises it from the global RNG (trunc_normal_, std 0.02): make it a function of the key
            v = torch.randn(shape, generator=g, device="cpu") * 0.2
        elif leaf == "bias":
            v = (torch.rand(shape, generator=g, device="cpu") - 0.5) * 0.1
        else:
            v = t.clone()
        out[key] = v.to(t.dtype)
    # calibration: latent gain on the last analysis conv, log-uniform scale bias on every EP head
    # (united: index 16 of each branch; single-modality ELIC: index 13 of its one Sequential)
    for k in [f"g_a.{br}_analysis_transform.16.{leaf}" for br in ("rgb", "depth") for leaf in ("weight", "bias")] + \
            [f"g_a.analysis_transform.13.{leaf}" for leaf in ("weight", "bias")]:
        if k in out:
            out[k] = out[k] * gain
    for key in list(out):
        if "entropy_parameters" in key and key.endswith("fusion.4.bias"):
            g = _gen(key + ":scale", seed)
            n = out[key].shape[0] // 2
            out[key][:n] = torch.exp(torch.rand(n, generator=g, device="cpu") * (math.log(hi) - math.log(lo)) + math.log(lo))
            out[key][n:] *= 0.0
        if "entropy_parameters" in key and key.endswith("fusion.4.weight"):
            out[key] = out[key] * 0.25   # predicted scales/means stay near their bias
    return out


def _is_transposed(model, key):
    mod = model
    try:
        for part in key.split(".")[:-1]:
            mod = getattr(mod, part) if not part.isdigit() else mod[int(part)]
    except Exception:
        return False
    return isinstance(mod, torch.nn.ConvTranspose2d)


def synthetic_pairs(n, height=480, width=640, seed=1234, depth_div=10000.0):
    """NYUv2-shaped pairs: rgb uint8/255 and 16-bit depth / depth_div (dataset/testDataset.py:47-55),
    low-pass filtered noise so the statistics are image-like. Returns fp32 NCHW on the CPU."""
    rgbs, depths = [], []
    for i in range(n):
        g = torch.Generator(device="cpu")
        g.manual_seed(seed + i)
        noise = torch.rand(1, 3, height + 8, width + 8, generator=g, device="cpu")
        smooth = torch.nn.functional.avg_pool2d(noise, 9, stride=1)
        smooth = (smooth - smooth.amin()) / (smooth.amax() - smooth.amin())
        detail = torch.rand(1, 3, height, width, generator=g, device="cpu") * 0.08
        rgb = torch.round(((smooth * 0.92 + detail).clamp(0, 1)) * 255.0) / 255.0
        yy = torch.linspace(0, 1, height, device="cpu").view(1, 1, -1, 1)
        xx = torch.linspace(0, 1, width, device="cpu").view(1, 1, 1, -1)
        dn = torch.nn.functional.avg_pool2d(torch.rand(1, 1, height + 16, width + 16, generator=g, device="cpu"), 17, stride=1)
        d16 = torch.round(700.0 + (9999.0 - 700.0) * (0.5 * yy + 0.2 * xx + 0.3 * (dn - dn.amin()) / (dn.amax() - dn.amin())))
        rgbs.append(rgb)
        depths.append(d16 / depth_div)
    return torch.cat(rgbs), torch.cat(depths)


def pad_to_multiple(x, p=64):
    """pad0: replicate-pad bottom/right to a multiple of p (dataset/utils.py:58-67)."""
    h, w = x.shape[-2:]
    H, W = (h + p - 1) // p * p, (w + p - 1) // p * p
    return torch.nn.functional.pad(x, (0, W - w, 0, H - h), mode="replicate")
