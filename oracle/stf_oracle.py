"""ORACLE (test infrastructure): torch-CPU fp32 restatement of SymmetricalTransFormerUnited (models/stf_united.py:15-678)
as functions over a state_dict.  The entropy side (h_a, h_s, Bi-CEE chain, coder) is ELIC_united's with N = 192, M = 384,
slices (24, 24, 48, 96, 192) (stf_united.py:639-642) and comes from oracle/model_oracle.py; only the two transforms differ.

What it follows:  PatchEmbed :374-405;  SwinTransformerBlock :118-212 (LayerNorm, cyclic shift, 4x4 windows, masked window
attention :83-115 with the relative position bias :97-101, MLP with exact GELU);  BasicLayer's shifted-window mask :334-352;
PatchMerging :215-248 (channel order x0 x1 x2 x3 = (0,0) (1,0) (0,1) (1,1));  PatchSplit :251-274;  the analysis / synthesis
walks with the residual bi_spf fusion :476-511, :584-613.
Parity pin: tests/test_oracle_stf.py against tests/golden/model_stf_united.npz, produced by the unmodified reference
(oracle/make_golden_stf.py)."""
import numpy as np
import torch
import torch.nn.functional as F

from .model_oracle import OracleCodec

DEPTHS, HEADS, WINDOW, EMBED = (2, 2, 6, 2), (3, 6, 12, 24), 4, 48


class StfOracle(OracleCodec):
    def __init__(self, state_dict, **kw):
        kw.setdefault("N", 192)
        kw.setdefault("M", 384)
        kw.setdefault("slice_ch", (24, 24, 48, 96, 192))
        super().__init__(state_dict, cross=True, **kw)

    # ------------------------------------------------------------ Swin pieces (tokens: [B, H, W, C])
    def _ln(self, p, x):
        return F.layer_norm(x, (x.shape[-1],), self.sd[p + ".weight"], self.sd[p + ".bias"], 1e-5)

    def _lin(self, p, x):
        return F.linear(x, self.sd[p + ".weight"], self.sd.get(p + ".bias"))

    @staticmethod
    def _add(a, b):
        """the residual adds (a hook: oracle/bf16_emulation.py rounds the sum, as a bf16 residual stream does)"""
        return a + b

    @staticmethod
    def _windows(x, ws):
        B, H, W, C = x.shape
        return x.view(B, H // ws, ws, W // ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws, C)

    @staticmethod
    def _unwindows(w, ws, B, H, W):
        return w.view(B, H // ws, W // ws, ws, ws, -1).permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, -1)

    @staticmethod
    def _shift_mask(H, W, ws, shift):
        img = torch.zeros((1, H, W, 1))
        cnt = 0
        for h in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            for w in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
                img[:, h, w, :] = cnt
                cnt += 1
        mw = StfOracle._windows(img, ws).squeeze(-1)
        m = mw.unsqueeze(1) - mw.unsqueeze(2)
        return m.masked_fill(m != 0, -100.0).masked_fill(m == 0, 0.0)

    def _block(self, p, x, heads, shift):
        B, H, W, C = x.shape
        ws = WINDOW
        assert H % ws == 0 and W % ws == 0, "inputs are multiples of 64: no window padding on this path"
        t = self._ln(p + ".norm1", x)
        if shift:
            t = torch.roll(t, shifts=(-shift, -shift), dims=(1, 2))
        win = self._windows(t, ws)                                             # [nW B, 16, C]
        qkv = self._lin(p + ".attn.qkv", win).reshape(win.shape[0], ws * ws, 3, heads, C // heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0] * (C // heads) ** -0.5, qkv[1], qkv[2]
        attn = q @ k.transpose(-2, -1)
        idx = self.sd[p + ".attn.relative_position_index"].view(-1).long()
        bias = self.sd[p + ".attn.relative_position_bias_table"][idx].view(ws * ws, ws * ws, -1).permute(2, 0, 1)
        attn = attn + bias.unsqueeze(0)
        if shift:
            m = self._shift_mask(H, W, ws, shift)
            nW = m.shape[0]
            attn = (attn.view(B, nW, heads, ws * ws, ws * ws) + m.unsqueeze(1).unsqueeze(0)).view(-1, heads, ws * ws, ws * ws)
        attn = attn.softmax(-1)
        o = (attn @ v).transpose(1, 2).reshape(win.shape[0], ws * ws, C)
        o = self._unwindows(self._lin(p + ".attn.proj", o), ws, B, H, W)
        if shift:
            o = torch.roll(o, shifts=(shift, shift), dims=(1, 2))
        x = self._add(x, o)
        h = F.gelu(self._lin(p + ".mlp.fc1", self._ln(p + ".norm2", x)))
        return self._add(x, self._lin(p + ".mlp.fc2", h))

    def _layer(self, p, x, depth, heads, down):
        for i in range(depth):
            x = self._block(f"{p}.blocks.{i}", x, heads, 0 if i % 2 == 0 else WINDOW // 2)
        if down == "merge":
            x = torch.cat([x[:, 0::2, 0::2], x[:, 1::2, 0::2], x[:, 0::2, 1::2], x[:, 1::2, 1::2]], -1)
            x = self._lin(p + ".downsample.reduction", self._ln(p + ".downsample.norm", x))
        elif down == "split":
            x = self._lin(p + ".downsample.reduction", self._ln(p + ".downsample.norm", x))
            x = F.pixel_shuffle(x.permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)
        return x.contiguous()

    def _fuse(self, p, r, d):
        """rgb_y + rgb_f, depth_y + depth_f around a bi_spf on NCHW views (stf_united.py:492-497)."""
        rc, dc = r.permute(0, 3, 1, 2).contiguous(), d.permute(0, 3, 1, 2).contiguous()
        rf, df = self._spf(p, rc, dc)
        return self._add(rc, rf).permute(0, 2, 3, 1).contiguous(), self._add(dc, df).permute(0, 2, 3, 1).contiguous()

    # ------------------------------------------------------------ transforms
    def g_a(self, rgb, depth):
        outs = []
        for name, x in (("rgb", rgb), ("depth", depth)):
            t = self._c(f"g_a.{name}_patch_embed.proj", x, stride=2).permute(0, 2, 3, 1)
            outs.append(self._ln(f"g_a.{name}_patch_embed.norm", t))
        r, d = outs
        li = 0
        for i in range(4):
            down = "merge" if i < 3 else None
            r = self._layer(f"g_a.rgb_ana_layers.{li}", r, DEPTHS[i], HEADS[i], down)
            d = self._layer(f"g_a.depth_ana_layers.{li}", d, DEPTHS[i], HEADS[i], down)
            li += 1
            if i < 3:
                r, d = self._fuse(f"g_a.rgb_ana_layers.{li}", r, d)
                li += 1
        return r.permute(0, 3, 1, 2).contiguous(), d.permute(0, 3, 1, 2).contiguous()

    def g_s(self, yr, yd):
        r, d = yr.permute(0, 2, 3, 1), yd.permute(0, 2, 3, 1)
        depths, heads = DEPTHS[::-1], HEADS[::-1]
        li = 0
        for i in range(4):
            down = "split" if i < 3 else None
            r = self._layer(f"g_s.rgb_syn_layers.{li}", r, depths[i], heads[i], down)
            d = self._layer(f"g_s.depth_syn_layers.{li}", d, depths[i], heads[i], down)
            li += 1
            if i < 3:
                r, d = self._fuse(f"g_s.rgb_syn_layers.{li}", r, d)
                li += 1
        outs = []
        for name, x in (("rgb", r), ("depth", d)):
            t = self._c(f"g_s.{name}_end_conv.0", x.permute(0, 3, 1, 2).contiguous(), pad=2)
            outs.append(self._c(f"g_s.{name}_end_conv.2", F.pixel_shuffle(t, 2), pad=1))
        return outs[0], outs[1]
