"""STF_united — SymmetricalTransFormerUnited (reference models/stf_united.py:616-678) on the B200 path.

The reference class is ELIC_united with N = 192, M = 384, slices (24, 24, 48, 96, 192) and its two transforms replaced by
Swin-transformer stacks (embed 48, depths 2-2-6-2, heads 3-6-12-24, 4x4 windows, patch size 2) with a residual bi_spf fusion
after every resolution change.  So is ours: the hyperprior, the Bi-CEE context chain, the coder and the whole program /
pipeline machinery are inherited from ELIC_united; this file only compiles the two transforms into launch plans:

    LayerNorm / PatchMerging gather / PixelShuffle / shifted-window attention   csrc/swin.cu
    every nn.Linear (qkv, proj, fc1 + GELU, fc2, reduction) as a 1x1 conv        conv_simt (fp32) / conv_halo (bf16, tcgen05)

Tokens "B, H*W, C" of the reference are the pixels of our NHWC views, so no layout change is needed anywhere.
Same state_dict keys in the same order as the reference (1244), same API, same bitstream.  H and W are multiples of 64, at
least 256 (the ESA at 1/16 scale needs 15 pixels a side, in the reference as well), so no window padding occurs.
"""
import torch
import torch.nn as nn

from . import lib as L
from .elic_united import ELIC_united, NONE
from .engine import PackedConv, _DT
from .modules import BiSpf
from .modules_stf import AnalysisTransformSTF, BasicLayer, PatchMerging, PatchSplit, SynthesisTransformSTF

GELU = L.ACT_GELU


class SymmetricalTransFormerUnited(ELIC_united):
    def __init__(self, config, patch_size=2, embed_dim=48, depths=(2, 2, 6, 2), num_heads=(3, 6, 12, 24), window_size=4, **kwargs):
        config.slice_ch = [24, 24, 48, 96, 192]          # stf_united.py:639-641 (the reference mutates the config too)
        config.N = 192
        config.M = 384
        kwargs.pop("channel", None)
        super().__init__(config=config, **kwargs)
        self.g_a = AnalysisTransformSTF(patch_size, embed_dim, tuple(depths), tuple(num_heads), window_size)
        self.g_s = SynthesisTransformSTF(patch_size, embed_dim, tuple(depths), tuple(num_heads), window_size)
        self.window_size = window_size

    def _check_inputs(self, rgb, depth):
        super()._check_inputs(rgb, depth)
        if rgb.shape[2] < 256 or rgb.shape[3] < 256:
            raise ValueError("H and W must be >= 256 (ESA max_pool2d(7, 3) on the 1/32-scale map of the 1/16-scale fusion)")

    # ------------------------------------------------------------------ pieces
    def _pc_linear(self, lin):
        """nn.Linear as a 1x1 conv (tokens are pixels)."""
        if self._packed is None:
            self._packed = {}
        key = ("linear", id(lin))
        if key not in self._packed:
            m = nn.Conv2d(lin.in_features, lin.out_features, 1, bias=lin.bias is not None)
            with torch.no_grad():
                m.weight.copy_(lin.weight.detach().float().cpu()[:, :, None, None])
                if lin.bias is not None:
                    m.bias.copy_(lin.bias.detach().float().cpu())
            self._packed[key] = PackedConv(m, self.device)
        return self._packed[key]

    def _ln(self, b, norm, x, out=None, out_dtype=None, gather=False):
        """nn.LayerNorm over the channels of every pixel; gather=True: PatchMerging's 2x2 concatenation first."""
        g = self._dev32(("ln_w", id(norm)), lambda: norm.weight)
        be = self._dev32(("ln_b", id(norm)), lambda: norm.bias)
        if gather:
            C = 4 * x.C
            if out is None:
                out = b.alloc(x.N, x.H // 2, x.W // 2, C, out_dtype)
            npix = x.N * (x.H // 2) * (x.W // 2)
        else:
            C = x.C
            if out is None:
                out = b.alloc(x.N, x.H, x.W, C, out_dtype)
            npix = x.N * x.H * x.W
        assert C == norm.normalized_shape[0]
        b.op("rgbd_layernorm", x.ptr(), _DT[x.dtype], out.ptr(), _DT[out.dtype], npix, C, x.cstride, x.coff, out.cstride, out.coff,
             g.data_ptr(), be.data_ptr(), float(norm.eps), int(gather), x.H, x.W)
        b.prog.keep.extend([g, be, x.buf, out.buf])
        return out

    def _block(self, b, m, x, out_dtype=None):
        """SwinTransformerBlock.forward (stf_united.py:162-212)."""
        t = self._ln(b, m.norm1, x)
        qkv = b.conv(self._pc_linear(m.attn.qkv), t)
        b.release(t)
        a = b.alloc(x.N, x.H, x.W, x.C)
        table = self._dev32(("rpb", id(m.attn)), lambda: m.attn.relative_position_bias_table)
        b.op("rgbd_window_attention", qkv.ptr(), a.ptr(), _DT[a.dtype], x.N, x.H, x.W, x.C, m.num_heads, m.window_size, m.shift_size,
             table.data_ptr(), float(m.attn.scale), qkv.cstride, qkv.coff, a.cstride, a.coff)
        b.prog.keep.extend([table, qkv.buf, a.buf])
        b.release(qkv)
        x1 = b.conv(self._pc_linear(m.attn.proj), a, res=x)            # shortcut + attention
        b.release(a)
        t = self._ln(b, m.norm2, x1)
        h = b.conv(self._pc_linear(m.mlp.fc1), t, act=GELU)
        b.release(t)
        y = b.conv(self._pc_linear(m.mlp.fc2), h, res=x1, out_dtype=out_dtype)      # x + mlp(norm2(x))
        b.release(h, x1)
        return y

    def _layer(self, b, layer, x, out_dtype=None, keep_input=False):
        """BasicLayer.forward (stf_united.py:328-371): the blocks, then PatchMerging / PatchSplit.  keep_input: x belongs
        to the caller (the y_hat buffers of the context chain) and is not returned to the plan's arena."""
        n = len(layer.blocks)
        for i, blk in enumerate(layer.blocks):
            y = self._block(b, blk, x, out_dtype=out_dtype if (i == n - 1 and layer.downsample is None) else None)
            if not (keep_input and i == 0):
                b.release(x)
            x = y
        ds = layer.downsample
        if isinstance(ds, PatchMerging):
            t = self._ln(b, ds.norm, x, gather=True)
            b.release(x)
            x = b.conv(self._pc_linear(ds.reduction), t)
            b.release(t)
        elif isinstance(ds, PatchSplit):
            t = self._ln(b, ds.norm, x)
            b.release(x)
            u = b.conv(self._pc_linear(ds.reduction), t)
            b.release(t)
            x = self._shuffle(b, u)
            b.release(u)
        return x

    def _shuffle(self, b, u):
        out = b.alloc(u.N, 2 * u.H, 2 * u.W, u.C // 4, u.dtype)
        b.op("rgbd_pixel_shuffle2", u.ptr(), out.ptr(), _DT[u.dtype], u.N, u.H, u.W, u.C // 4, u.cstride, u.coff, out.cstride, out.coff)
        b.prog.keep.extend([u.buf, out.buf])
        return out

    def _walk(self, b, rgb_layers, depth_layers, r, d, final_dtype=None, keep_inputs=False):
        n = len(rgb_layers)
        for i in range(n):
            rm, dm = rgb_layers[i], depth_layers[i]
            if isinstance(rm, BiSpf):
                # rgb_y + rgb_f, depth_y + depth_f (stf_united.py:492-497): the residual rides in the ESA gate conv's epilogue
                nr, nd = b.alloc(r.N, r.H, r.W, r.C), b.alloc(d.N, d.H, d.W, d.C)
                self._bispf(b, rm, r, d, nr, nd, res_r=r, res_d=d)
                b.release(r, d)
                r, d = nr, nd
                continue
            assert isinstance(rm, BasicLayer)
            last = i == n - 1
            r = self._layer(b, rm, r, out_dtype=final_dtype if last else None, keep_input=keep_inputs and i == 0)
            d = self._layer(b, dm, d, out_dtype=final_dtype if last else None, keep_input=keep_inputs and i == 0)
        return r, d

    # ------------------------------------------------------------------ the two hooks of ELIC_united
    def _analysis(self, b, B, H, W):
        p = b.prog
        b.stage = "io"
        in_r, in_d = b.raw((B, 3, H, W), torch.float32), b.raw((B, 1, H, W), torch.float32)
        p.io["rgb"], p.io["depth"] = in_r, in_d
        outs = []
        for src, Cc, pe in ((in_r, 3, self.g_a.rgb_patch_embed), (in_d, 1, self.g_a.depth_patch_embed)):
            x = b.alloc(B, H, W, Cc, torch.float32)
            b.op("rgbd_nchw_to_nhwc", src.data_ptr(), x.ptr(), _DT[x.dtype], B, Cc, H, W, x.cstride, x.coff, 0)
            b.stage = "g_a"
            # PatchEmbed (stf_united.py:374-405): 2x2 stride-2 conv on the fp32 image (CUDA cores: 12 / 4 inputs per output), LayerNorm
            t = b.conv(self._pc(pe.proj), x, out_dtype=torch.float32)
            b.release(x)
            outs.append(self._ln(b, pe.norm, t))
            b.release(t)
            b.stage = "io"
        b.stage = "g_a"
        return self._walk(b, self.g_a.rgb_ana_layers, self.g_a.depth_ana_layers, outs[0], outs[1], final_dtype=torch.float32)

    def _synthesis(self, b, yhat_r, yhat_d):
        # the y_hat buffers are owned by the chain (and read by the parity tests): the first block must not release them
        r, d = self._walk(b, self.g_s.rgb_syn_layers, self.g_s.depth_syn_layers, yhat_r, yhat_d, keep_inputs=True)
        outs = []
        for x, end in ((r, self.g_s.rgb_end_conv), (d, self.g_s.depth_end_conv)):
            t = b.conv(self._pc(end[0]), x)                       # 5x5, embed -> 4 embed
            b.release(x)
            s = self._shuffle(b, t)
            b.release(t)
            outs.append(b.conv(self._pc(end[2]), s))              # 3x3, embed -> 3 / 1
            b.release(s)
        return outs[0], outs[1]


STF_united = SymmetricalTransFormerUnited
