"""ORACLE (test infrastructure): torch-CPU fp32 restatement of the reference's ELIC_united /
ELIC_united_R2D forward / compress / decompress, written as plain functions over a state_dict.

It is the checker for the CUDA path (tests/, smoke()) and the CPU arm that bench.py times
(`cpu_baseline`, `--impl reference`); the product never imports it.

What it follows (paths relative to the reference root):
  g_a / g_s walk, bi_spf concat            modules/transform/analysis.py:168-181, synthesis.py:171-184
  ResidualBottleneck                       modules/layers/res_blk.py:19-27
  AttentionBlock                           CompressAI/compressai/layers/layers.py:206-213
  ESA / SE_Block / bi_spf                  modules/transform/attention.py:22-48,63-97
  h_a, h_s                                 analysis.py:246-249, synthesis.py:316-343,356-380
  EntropyParametersEX / ChannelContextEX   modules/transform/entropy.py:75-78, context.py:21-30
  per-slice 4-step order                   models/elic_united.py:117-190,265-348,454-541
                                           models/elic_united_R2D.py:73-326
  ckbd squeeze order                       utils/ckbd.py:51-105
  quantise / indexes / likelihood          CompressAI/compressai/entropy_models/entropy_models.py:118-146,
                                           369-446,534-568
The arithmetic itself (conv, erfc, round, interpolate, max_pool) is torch CPU, as in the reference.
Parity pin: tests/test_oracle_model.py compares this file with tensors/bytes the unmodified
reference produced (tests/golden/model_*.npz, written by oracle/make_golden.py).
For batch B > 1 it emits one y stream per image (the product's layout).
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import coder


class OracleCodec:
    def __init__(self, state_dict, cross=True, N=192, M=320, slice_ch=(16, 16, 32, 64, 192), use_ref_coder=False):
        self.sd = {k: v.detach().cpu() for k, v in state_dict.items()}
        self.cross, self.N, self.M, self.slice_ch = cross, N, M, list(slice_ch)
        self.coder = coder
        if use_ref_coder:  # the reference's own compiled coder (oracle/_ref), same call shapes
            from . import ref_coder
            self.coder = ref_coder
        self.trace = None

    # ------------------------------------------------------------ primitives
    def _c(self, name, x, stride=1, pad=0):
        return F.conv2d(x, self.sd[name + ".weight"], self.sd[name + ".bias"], stride=stride, padding=pad)

    def _d(self, name, x, k=5, s=2):
        return F.conv_transpose2d(x, self.sd[name + ".weight"], self.sd[name + ".bias"], stride=s,
                                  padding=k // 2, output_padding=s - 1)

    def _rb(self, p, x):
        t = F.relu(self._c(p + ".branch.0", x))
        t = F.relu(self._c(p + ".branch.2", t, pad=1))
        t = self._c(p + ".branch.4", t)
        idn = self._c(p + ".skip", x) if (p + ".skip.weight") in self.sd else x
        return t + idn

    def _ru(self, p, x):
        t = F.relu(self._c(p + ".conv.0", x))
        t = F.relu(self._c(p + ".conv.2", t, pad=1))
        return F.relu(self._c(p + ".conv.4", t) + x)

    def _attn(self, p, x):
        a = x
        for i in range(3):
            a = self._ru(f"{p}.conv_a.{i}", a)
        b = x
        for i in range(3):
            b = self._ru(f"{p}.conv_b.{i}", b)
        b = self._c(p + ".conv_b.3", b)
        return a * torch.sigmoid(b) + x

    def _esa(self, p, x):
        c1_ = self._c(p + ".conv1", x)
        c1 = self._c(p + ".conv2", c1_, stride=2)
        v = F.max_pool2d(c1, kernel_size=7, stride=3)
        v = F.relu(self._c(p + ".conv_max", v, pad=1))
        c3 = F.relu(self._c(p + ".conv3", v, pad=1))
        c3 = self._c(p + ".conv3_", c3, pad=1)
        c3 = F.interpolate(c3, (x.size(2), x.size(3)), mode="bilinear", align_corners=False)
        m = torch.sigmoid(self._c(p + ".conv4", c3 + self._c(p + ".conv_f", c1_)))
        return x * m

    def _se(self, p, x):
        y = x.mean(dim=(2, 3))
        y = torch.sigmoid(F.linear(F.relu(F.linear(y, self.sd[p + ".fc.0.weight"])), self.sd[p + ".fc.2.weight"]))
        return x * y[:, :, None, None]

    def _spf(self, p, r, d):
        rr = F.relu(self._c(p + ".r_ext", r, pad=1))
        dd = F.relu(self._c(p + ".d_ext", d, pad=1))
        d_out = self._esa(p + ".d_esa", torch.cat((dd, rr), 1))
        r_out = self._esa(p + ".r_esa", torch.cat((rr, dd), 1)) if self.cross else None
        return r_out, d_out

    # ------------------------------------------------------------ transforms
    _GA = ["c", "rb", "rb", "rb", "spf", "c", "rb", "rb", "rb", "at", "spf", "c", "rb", "rb", "rb", "spf", "c", "at"]
    _GS = ["at", "dc", "spf", "rb", "rb", "rb", "dc", "at", "spf", "rb", "rb", "rb", "dc", "spf", "rb", "rb", "rb", "dc"]

    def _walk(self, root, kinds, r, d):
        rp, dp = (f"{root}.rgb_analysis_transform", f"{root}.depth_analysis_transform") if root == "g_a" else (
            f"{root}.rgb_synthesis_transform", f"{root}.depth_synthesis_transform")
        for i, kind in enumerate(kinds):
            if kind == "spf":
                rf, df = self._spf(f"{rp}.{i}", r, d)
                if rf is not None:
                    r = torch.cat((r, rf), 1)
                d = torch.cat((d, df), 1)
                continue
            f = {"c": lambda p, x: self._c(p, x, stride=2, pad=2), "dc": self._d, "rb": self._rb, "at": self._attn}[kind]
            r, d = f(f"{rp}.{i}", r), f(f"{dp}.{i}", d)
        return r, d

    def g_a(self, rgb, depth):
        return self._walk("g_a", self._GA, rgb, depth)

    def g_s(self, yr, yd):
        return self._walk("g_s", self._GS, yr, yd)

    def h_a(self, yr, yd):
        out = []
        for p, y in (("h_a.rgb_reduction", yr), ("h_a.depth_reduction", yd)):
            t = F.relu(self._c(p + ".0", y, pad=1))
            t = F.relu(self._c(p + ".2", t, stride=2, pad=2))
            out.append(self._c(p + ".4", t, stride=2, pad=2))
        return out

    def _hblock(self, p, x, last):
        f = self._se(p + ".se", x)
        if last:
            return self._d(p + ".deconv", f, k=3, s=1)
        return F.leaky_relu(self._d(p + ".deconv", f))

    def h_s(self, zr, zd):
        cat = torch.cat
        if self.cross:
            r1 = self._hblock("h_s.r_h_s1", cat((zr, zd), 1), False)
            d1 = self._hblock("h_s.d_h_s1", cat((zd, zr), 1), False)
            r2 = self._hblock("h_s.r_h_s2", cat((r1, d1), 1), False)
            d2 = self._hblock("h_s.d_h_s2", cat((d1, r1), 1), False)
            return self._hblock("h_s.r_h_s3", cat((r2, d2), 1), True), self._hblock("h_s.d_h_s3", cat((d2, r2), 1), True)
        r1 = self._hblock("h_s.r_h_s1", zr, False)
        d1 = self._hblock("h_s.d_h_s1", cat((zd, zr), 1), False)
        r2 = self._hblock("h_s.r_h_s2", r1, False)
        d2 = self._hblock("h_s.d_h_s2", cat((d1, r1), 1), False)
        return self._hblock("h_s.r_h_s3", r2, True), self._hblock("h_s.d_h_s3", cat((d2, r2), 1), True)

    def _ep(self, p, x):
        x = x + self._se(p + ".se", x)
        t = F.relu(self._c(p + ".fusion.0", x))
        t = F.relu(self._c(p + ".fusion.2", t, pad=1))
        return self._c(p + ".fusion.4", t, pad=2).chunk(2, 1)  # (scales, means)

    def _chctx(self, p, x):
        t = F.relu(self._c(p + ".fushion.0", x, pad=2))
        t = F.relu(self._c(p + ".fushion.2", t, pad=2))
        return self._c(p + ".fushion.4", t, pad=2)

    # ------------------------------------------------------------ entropy helpers
    @staticmethod
    def _mask(h, w, parity):
        hh = torch.arange(h)[:, None]
        ww = torch.arange(w)[None, :]
        return ((hh + ww) % 2 == (1 - parity))  # parity 0 = anchor = (h + w) odd

    @staticmethod
    def _squeeze(x, parity):
        """[B,C,H,W] -> [B,C,H,W/2] keeping the parity sites, row by row (ckbd.py:51-64)"""
        B, C, H, W = x.shape
        out = torch.empty(B, C, H, W // 2, dtype=x.dtype)
        a, b = (1, 0) if parity == 0 else (0, 1)
        out[:, :, 0::2, :] = x[:, :, 0::2, a::2]
        out[:, :, 1::2, :] = x[:, :, 1::2, b::2]
        return out

    @staticmethod
    def _unsqueeze(x, parity):
        B, C, H, Wq = x.shape
        out = torch.zeros(B, C, H, Wq * 2, dtype=x.dtype)
        a, b = (1, 0) if parity == 0 else (0, 1)
        out[:, :, 0::2, a::2] = x[:, :, 0::2, :]
        out[:, :, 1::2, b::2] = x[:, :, 1::2, :]
        return out

    def _indexes(self, which, scales):
        table = self.sd[f"{which}_gaussian_conditional.scale_table"]
        bound = self.sd[f"{which}_gaussian_conditional.lower_bound_scale.bound"]
        s = torch.max(scales, bound)
        idx = torch.full(s.shape, len(table) - 1, dtype=torch.int32)
        for t in table[:-1]:
            idx -= (s <= t).int()
        return idx

    def _gauss_likelihood(self, which, y, scales, means):
        bound = self.sd[f"{which}_gaussian_conditional.lower_bound_scale.bound"]
        out = torch.round(y - means) + means
        v = torch.abs(out - means)
        s = torch.max(scales, bound)
        c = float(-(2 ** -0.5))
        upper = 0.5 * torch.erfc(c * ((0.5 - v) / s))
        lower = 0.5 * torch.erfc(c * ((-0.5 - v) / s))
        return torch.clamp_min(upper - lower, 1e-9)

    def gc_tables(self, which):
        p = f"{which}_gaussian_conditional."
        return coder.Tables(self.sd[p + "_quantized_cdf"].numpy(), self.sd[p + "_cdf_length"].numpy(),
                            self.sd[p + "_offset"].numpy())

    def eb_tables(self, which):
        p = f"{which}_entropy_bottleneck."
        return coder.Tables(self.sd[p + "_quantized_cdf"].numpy(), self.sd[p + "_cdf_length"].numpy(),
                            self.sd[p + "_offset"].numpy())

    def _eb_logits(self, which, v):
        p = f"{which}_entropy_bottleneck."
        logits = v
        for i in range(5):
            logits = torch.matmul(F.softplus(self.sd[p + f"_matrix{i}"]), logits) + self.sd[p + f"_bias{i}"]
            if i < 4:
                logits = logits + torch.tanh(self.sd[p + f"_factor{i}"]) * torch.tanh(logits)
        return logits

    def eb_forward(self, which, z):
        """-> (z_hat via ste_round, likelihood)  (elic_united.py:239-245, entropy_models.py:391-428)"""
        med = self.sd[f"{which}_entropy_bottleneck.quantiles"][:, :, 1:2]  # [C,1,1]
        B, C, H, W = z.shape
        v = z.permute(1, 2, 3, 0).reshape(C, 1, -1)
        out = torch.round(v - med) + med
        lower, upper = self._eb_logits(which, out - 0.5), self._eb_logits(which, out + 0.5)
        sign = -torch.sign(lower + upper)
        lik = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower)).clamp_min(1e-9)
        lik = lik.reshape(C, H, W, B).permute(3, 0, 1, 2).contiguous()
        m4 = med.reshape(1, C, 1, 1)
        d = z - m4
        return (torch.round(d) - d + d) + m4, lik

    # ------------------------------------------------------------ the Bi-CEE chain
    def _chain(self, hyper_r, hyper_d, step):
        """Runs the 5 x 4 coding steps; `step(which, idx, parity, scales, means)` returns the
        (masked) y_hat part of that step.  Returns full y_hat per modality."""
        sc = self.slice_ch
        yh_r, yh_d = [], []
        for idx, g in enumerate(sc):
            if self.cross:
                base_r = base_d = [hyper_r, hyper_d]
            else:
                base_r, base_d = [hyper_r], [hyper_r, hyper_d]
            if idx:
                ch_r = self._chctx(f"rgb_channel_context.{idx}", torch.cat(yh_r, 1))
                ch_d = self._chctx(f"depth_channel_context.{idx}", torch.cat(yh_d, 1))
                base_r = base_r + ([ch_r, ch_d] if self.cross else [ch_r])
                base_d = base_d + [ch_r, ch_d]
            cat = torch.cat
            sr, mr = self._ep(f"rgb_entropy_parameters_anchor.{idx}", cat(base_r, 1))
            ra = step("rgb", idx, 0, sr, mr)
            loc_r = self._c(f"rgb_local_context.{idx}", ra, pad=2)
            sd_, md = self._ep(f"depth_entropy_parameters_anchor.{idx}", cat([loc_r] + base_d, 1))
            da = step("depth", idx, 0, sd_, md)
            loc_d = self._c(f"depth_local_context.{idx}", da, pad=2)
            rin = [loc_r, loc_d] + base_r if self.cross else [loc_r] + base_r
            sr, mr = self._ep(f"rgb_entropy_parameters_nonanchor.{idx}", cat(rin, 1))
            rn = step("rgb", idx, 1, sr, mr)
            r_full = rn + ra
            loc_r = self._c(f"rgb_local_context_anchor_with_nonanchor.{idx}", r_full, pad=2)
            sd_, md = self._ep(f"depth_entropy_parameters_nonanchor.{idx}", cat([loc_r, loc_d] + base_d, 1))
            dn = step("depth", idx, 1, sd_, md)
            yh_r.append(r_full)
            yh_d.append(dn + da)
        return torch.cat(yh_r, 1), torch.cat(yh_d, 1)

    def _slice(self, y, idx):
        a = sum(self.slice_ch[:idx])
        return y[:, a:a + self.slice_ch[idx]]

    # ------------------------------------------------------------ public
    @torch.no_grad()
    def forward(self, rgb, depth):
        yr, yd = self.g_a(rgb, depth)
        zr, zd = self.h_a(yr, yd)
        zr_hat, lzr = self.eb_forward("rgb", zr)
        zd_hat, lzd = self.eb_forward("depth", zd)
        hr, hd = self.h_s(zr_hat, zd_hat)
        y = {"rgb": yr, "depth": yd}
        lik = {"rgb": torch.zeros_like(yr), "depth": torch.zeros_like(yd)}
        H, W = yr.shape[2:]

        def step(which, idx, parity, scales, means):
            m = self._mask(H, W, parity)
            ys = self._slice(y[which], idx)
            d = (ys - means) * m
            part = ((torch.round(d) - d + d) + means * m)
            a = sum(self.slice_ch[:idx])
            l = self._gauss_likelihood(which, ys, scales, means)
            lik[which][:, a:a + self.slice_ch[idx]] += l * m
            return part

        yr_hat, yd_hat = self._chain(hr, hd, step)
        xr, xd = self.g_s(yr_hat, yd_hat)
        return {"x_hat": {"r": xr, "d": xd}, "r_likelihoods": {"y": lik["rgb"], "z": lzr},
                "d_likelihoods": {"y": lik["depth"], "z": lzd},
                "_trace": {"y_r": yr, "y_d": yd, "z_r": zr, "z_d": zd, "hyper_r": hr, "hyper_d": hd,
                           "yhat_r": yr_hat, "yhat_d": yd_hat}}

    def _z_code(self, which, z):
        med = self.sd[f"{which}_entropy_bottleneck.quantiles"][:, 0, 1].reshape(1, -1, 1, 1)
        sym = torch.round(z - med).int()
        B, C, H, W = z.shape
        idx = torch.arange(C, dtype=torch.int32).view(1, -1, 1, 1).expand(B, C, H, W)
        t = self.eb_tables(which)
        strings = [self.coder.encode_with_indexes(sym[i].reshape(-1).numpy(), idx[i].reshape(-1).numpy(), t)
                   for i in range(B)]
        return strings, sym.float() + med, sym

    @torch.no_grad()
    def compress(self, rgb, depth, trace=False):
        yr, yd = self.g_a(rgb, depth)
        zr, zd = self.h_a(yr, yd)
        rz, zr_hat, zsym_r = self._z_code("rgb", zr)
        dz, zd_hat, zsym_d = self._z_code("depth", zd)
        hr, hd = self.h_s(zr_hat, zd_hat)
        y = {"rgb": yr, "depth": yd}
        B = rgb.shape[0]
        syms = {"rgb": [[] for _ in range(B)], "depth": [[] for _ in range(B)]}
        idxs = {"rgb": [[] for _ in range(B)], "depth": [[] for _ in range(B)]}
        steps = []

        def step(which, idx, parity, scales, means):
            ys = self._squeeze(self._slice(y[which], idx), parity)
            ss, mm = self._squeeze(scales, parity), self._squeeze(means, parity)
            ind = self._indexes(which, ss)
            sym = torch.round(ys - mm).int()
            for i in range(B):
                syms[which][i].append(sym[i].reshape(-1).numpy())
                idxs[which][i].append(ind[i].reshape(-1).numpy())
            if trace:
                steps.append((which, idx, parity, scales.clone(), means.clone()))
            return self._unsqueeze(sym.float() + mm, parity)

        yr_hat, yd_hat = self._chain(hr, hd, step)
        out = {"shape": tuple(zr.shape[-2:])}
        flat = {}
        for which, key in (("rgb", "r_strings"), ("depth", "d_strings")):
            t = self.gc_tables(which)
            ystr = []
            for i in range(B):
                s, x = np.concatenate(syms[which][i]), np.concatenate(idxs[which][i])
                flat[(which, i)] = (s, x)
                ystr.append(self.coder.encode_with_indexes(s, x, t))
            out[key] = [ystr, rz if which == "rgb" else dz]
        if trace:
            out["_trace"] = {"y_r": yr, "y_d": yd, "z_r": zr, "z_d": zd, "zsym_r": zsym_r, "zsym_d": zsym_d,
                             "hyper_r": hr, "hyper_d": hd, "yhat_r": yr_hat, "yhat_d": yd_hat, "symbols": flat,
                             "steps": steps}
        return out

    @torch.no_grad()
    def decompress(self, rgb_strings, depth_strings, shape):
        B = len(rgb_strings[1])
        hz, wz = int(shape[0]), int(shape[1])
        zhat = {}
        for which, strs in (("rgb", rgb_strings), ("depth", depth_strings)):
            t = self.eb_tables(which)
            C = self.sd[f"{which}_entropy_bottleneck._quantized_cdf"].shape[0]
            med = self.sd[f"{which}_entropy_bottleneck.quantiles"][:, 0, 1].reshape(1, -1, 1, 1)
            idx = np.repeat(np.arange(C, dtype=np.int32), hz * wz)
            vals = [self.coder.decode_with_indexes(strs[1][i], idx, t).reshape(C, hz, wz) for i in range(B)]
            zhat[which] = torch.from_numpy(np.stack(vals)).float() + med
        hr, hd = self.h_s(zhat["rgb"], zhat["depth"])
        dec = {"rgb": [self.coder.Decoder(rgb_strings[0][i]) for i in range(B)],
               "depth": [self.coder.Decoder(depth_strings[0][i]) for i in range(B)]}
        tabs = {"rgb": self.gc_tables("rgb"), "depth": self.gc_tables("depth")}

        def step(which, idx, parity, scales, means):
            ss, mm = self._squeeze(scales, parity), self._squeeze(means, parity)
            ind = self._indexes(which, ss)
            vals = [dec[which][i].decode_stream(ind[i].reshape(-1).numpy(), tabs[which]).reshape(ind[i].shape)
                    for i in range(B)]
            sym = torch.from_numpy(np.stack(vals)).float()
            return self._unsqueeze(sym + mm, parity)

        yr_hat, yd_hat = self._chain(hr, hd, step)
        xr, xd = self.g_s(yr_hat, yd_hat)
        return {"x_hat": {"r": xr.clamp_(0, 1), "d": xd.clamp_(0, 1)}, "_trace": {"yhat_r": yr_hat, "yhat_d": yd_hat}}
