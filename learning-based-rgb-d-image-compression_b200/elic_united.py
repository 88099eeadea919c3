"""ELIC_united / ELIC_united_R2D — B200-native drop-in for the reference's RGB-D codec classes.

Same constructor, state_dict keys and API surface as the reference
(models/elic_united.py:15-620, models/elic_united_R2D.py:9-326):

    net = ELIC_united(config=Config(N, M, slice_num, slice_ch, quant), channel=4).eval()
    net.load_state_dict(ckpt["state_dict"]); net.update(force=True); net.to("cuda")
    out  = net(rgb, depth)                      # x_hat + likelihoods      (forward,   :234-263)
    code = net.compress(rgb, depth)             # r_strings / d_strings / shape (compress, :403-427)
    rec  = net.decompress(code["r_strings"], code["d_strings"], code["shape"])   # (:429-452)

Everything below the API is new: the layer graph is compiled (engine.py) into a static list of
launches of hand-written sm_100a kernels behind the C-ABI in include/rgbd_b200.h — implicit-GEMM
convs over NHWC pixels, fused checkerboard/quantise/index kernels and a GPU rANS coder that is
byte-identical to compressai.ans.  No torch arithmetic runs on the hot path, there is no CPU
fallback, and the module refuses to run without its CUDA library.

Bitstream layout: per image and modality one y stream covering the 5 channel groups x
{anchor, non-anchor} in the reference's order (elic_united.py:377-399, utils/ckbd.py:83-105) and one
z stream; for batch 1 the result is identical to the reference's.  For batch B > 1 the y entry holds
B per-image strings (the reference would interleave the batch into one stream, which its own
harness never does: utils/IOutils.py:84-86, config/args.py:66-68).
"""
import ctypes
import os
import time

import numpy as np
import torch
import torch.nn as nn

from . import lib as L
from .engine import Builder, PackedConv, View, _DT
from .entropy_models import EntropyBottleneck, GaussianConditional, get_scale_table
from .modules import (AnalysisTransform, AttentionBlock, BiSpf, BiSpfSingle, ChannelContextEX,
                      EntropyParametersEX, HyperAnalysis, HyperSynthesis, ResidualBottleneck,
                      SynthesisTransform)

RELU, LEAKY, NONE = L.ACT_RELU, L.ACT_LEAKY, L.ACT_NONE


class ELIC_united(nn.Module):
    cross = True  # bidirectional RGB<->depth (Bi-CPT / Bi-CEE); the R2D subclass clears it

    def __init__(self, config, **kwargs):
        super().__init__()
        N, M = config.N, config.M
        self.N, self.M = N, M
        self.quant = config.quant
        self.slice_num = config.slice_num
        self.slice_ch = list(config.slice_ch)
        assert sum(self.slice_ch) == M and len(self.slice_ch) == self.slice_num
        sc = self.slice_ch
        cross = self.cross
        self.g_a = AnalysisTransform(N, M, cross)
        self.g_s = SynthesisTransform(N, M, cross)
        self.h_a = HyperAnalysis(N, M)
        self.h_s = HyperSynthesis(N, M, cross)

        def local():
            return nn.ModuleList(nn.Conv2d(c, 2 * c, 5, 1, 2) for c in sc)

        self.rgb_local_context = local()
        self.rgb_local_context_anchor_with_nonanchor = local()
        self.depth_local_context = local()
        self.rgb_channel_context = nn.ModuleList(
            ChannelContextEX(sum(sc[:i]), sc[i] * 2) if i else None for i in range(self.slice_num))
        self.depth_channel_context = nn.ModuleList(
            ChannelContextEX(sum(sc[:i]), sc[i] * 2) if i else None for i in range(self.slice_num))
        # context widths: hyper (2M per modality) + channel ctx (2g per modality, idx > 0) + locals
        rh = 4 * M if cross else 2 * M        # hyper channels the rgb branch sees
        rc = 4 if cross else 2                # channel-ctx multiples of g for rgb
        self.rgb_entropy_parameters_anchor = nn.ModuleList(
            EntropyParametersEX(rh + (rc * c if i else 0), 2 * c) for i, c in enumerate(sc))
        self.depth_entropy_parameters_anchor = nn.ModuleList(
            EntropyParametersEX(4 * M + 2 * c + (4 * c if i else 0), 2 * c) for i, c in enumerate(sc))
        rl = 4 if cross else 2                # local-ctx channels for rgb non-anchor
        self.rgb_entropy_parameters_nonanchor = nn.ModuleList(
            EntropyParametersEX(rh + rl * c + (rc * c if i else 0), 2 * c) for i, c in enumerate(sc))
        self.depth_entropy_parameters_nonanchor = nn.ModuleList(
            EntropyParametersEX(4 * M + 4 * c + (4 * c if i else 0), 2 * c) for i, c in enumerate(sc))

        self.entropy_bottleneck = None
        self.rgb_entropy_bottleneck = EntropyBottleneck(N)
        self.depth_entropy_bottleneck = EntropyBottleneck(N)
        self.gaussianConditional = None
        self.rgb_gaussian_conditional = GaussianConditional(None)
        self.depth_gaussian_conditional = GaussianConditional(None)

        self.precision = kwargs.get("precision", "fp32")   # "fp32" | "bf16"
        self.use_cuda_graph = kwargs.get("cuda_graph", False)
        self.tensor_cores = kwargs.get("tensor_cores", True)   # bf16 mode: tcgen05 convs (False: CUDA cores)
        # bf16 tensor-core mode: ResidualBottleneck / ResidualUnit as ONE launch with on-chip intermediates
        self.fuse_blocks = kwargs.get("fuse_blocks", os.environ.get("RGBD_FUSE_BLOCKS", "1") != "0")
        # Bitstream layout of the y latents.  "single" (default) = the reference's: one rANS stream per image and modality
        # (decodable by the reference).  "multi" (opt-in) = that stream cut into equal sub-streams of
        # `sub_channels` channels of one checkerboard half, each a complete RansEncoder string, carried as further strings
        # of the same container entry: every coding step then decodes many short streams concurrently.
        self.stream_layout = kwargs.get("stream_layout", "single")
        self.sub_channels = int(kwargs.get("sub_channels", 4))
        self._packed = None      # id(module) -> PackedConv, rebuilt when weights change
        self._programs = {}
        self._aux = {}

    # ------------------------------------------------------------------ reference API
    def count_parameters(self, only_trainable=False):
        return sum(p.numel() for p in self.parameters() if p.requires_grad or not only_trainable)

    def aux_loss(self):
        raise NotImplementedError("training is out of scope of the B200 inference path")

    def update(self, scale_table=None, force=False):
        """models/elic_united.py:580-586 + CompressionModel.update (priors.py:73-92)."""
        if scale_table is None:
            scale_table = get_scale_table()
        r = self.rgb_gaussian_conditional.update_scale_table(scale_table, force=force)
        d = self.depth_gaussian_conditional.update_scale_table(scale_table, force=force)
        eb = False
        for m in (self.rgb_entropy_bottleneck, self.depth_entropy_bottleneck):
            eb |= m.update(force=force)
        self._invalidate()
        return (r & d) | eb

    def load_state_dict(self, state_dict, strict=False):
        """Resizes the CDF buffers to the checkpoint's sizes first (models/elic_united.py:588-620,
        utils/moduleFunc.py:42-88), then loads like the reference: a strict load is tried first; when it
        fails and the caller did not ask for strict=True, the mismatch is reported and a non-strict load
        follows (the reference's fallback branch, :612-620)."""
        for name in ("rgb_gaussian_conditional", "depth_gaussian_conditional", "rgb_entropy_bottleneck",
                     "depth_entropy_bottleneck"):
            mod = getattr(self, name)
            bufs = ["_quantized_cdf", "_offset", "_cdf_length"]
            if name.endswith("gaussian_conditional"):
                bufs.append("scale_table")
            for b in bufs:
                key = f"{name}.{b}"
                if key in state_dict:
                    cur = getattr(mod, b)
                    if cur.numel() == 0 or cur.shape != state_dict[key].shape:
                        setattr(mod, b, torch.empty(state_dict[key].shape, dtype=cur.dtype, device=cur.device))
            mod.invalidate()
        try:
            rv = super().load_state_dict(state_dict, strict=True)
        except RuntimeError as e:
            if strict:
                raise
            print("ELIC_united load state dict strict error:", str(e).splitlines()[0])
            rv = super().load_state_dict(state_dict, strict=False)
        self._invalidate()
        return rv

    def _apply(self, fn, *a, **k):
        rv = super()._apply(fn, *a, **k)
        self._invalidate()
        return rv

    def _invalidate(self):
        self._packed = None
        self._programs = {}
        self._aux = {}
        self.__dict__.pop("_slot_streams", None)   # streams belong to the device the module lived on
        for m in (self.rgb_gaussian_conditional, self.depth_gaussian_conditional, self.rgb_entropy_bottleneck,
                  self.depth_entropy_bottleneck):
            m.invalidate()

    def set_precision(self, precision):
        assert precision in ("fp32", "bf16")
        if precision != self.precision:
            self.precision = precision
            self._programs = {}

    # ------------------------------------------------------------------ weights on device
    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def act_dtype(self):
        return torch.float32 if self.precision == "fp32" else torch.bfloat16

    def _pc(self, mod, in_perm=None, split3=False, s2d=False):
        if self._packed is None:
            self._packed = {}
        key = (id(mod), split3, s2d)
        if key not in self._packed:
            self._packed[key] = PackedConv(mod, self.device, in_perm, split3, s2d)
        return self._packed[key]

    def _pc_concat(self, first, second):
        """One 1x1 conv over the channel concatenation [x_first | x_second]: weights [W_first | W_second] along K, biases
        summed.  y = W_first x_first + W_second x_second + (b_first + b_second) in ONE accumulation — the skip conv and the
        last conv of a widening / narrowing ResidualBottleneck (res_blk.py:7-27) become one launch, and their sum is formed
        in the fp32 accumulator instead of through a bf16 round trip."""
        if self._packed is None:
            self._packed = {}
        key = ("concat", id(first), id(second))
        if key not in self._packed:
            assert first.kernel_size == (1, 1) and second.kernel_size == (1, 1) and first.out_channels == second.out_channels
            m = nn.Conv2d(first.in_channels + second.in_channels, first.out_channels, 1)
            with torch.no_grad():
                m.weight.copy_(torch.cat([first.weight.detach().float().cpu(), second.weight.detach().float().cpu()], dim=1))
                m.bias.copy_(first.bias.detach().float().cpu() + second.bias.detach().float().cpu())
            self._packed[key] = PackedConv(m, self.device)
        return self._packed[key]

    def _dev32(self, key, make):
        if key not in self._aux:
            self._aux[key] = make().detach().to(device=self.device, dtype=torch.float32).contiguous()
        return self._aux[key]

    # ------------------------------------------------------------------ graph pieces
    def _rb(self, b, m, x, out=None):
        """ResidualBottleneck (res_blk.py:7-27): x (+skip) + 1x1(relu(3x3(relu(1x1 x))))"""
        pcs = [self._pc(m.branch[i]) for i in (0, 2, 4)]
        if self.fuse_blocks and b.can_fuse_block(*pcs, x):
            # one launch, intermediates on chip (csrc/conv_rb.cu); the optional 1x1 skip conv stays a launch of its own
            idn = x if m.skip is None else b.conv(self._pc(m.skip), x)
            y = b.fused_block(*pcs, x, res=idn, out=out)
            if idn is not x:
                b.release(idn)
            return y
        t1 = b.conv(self._pc(m.branch[0]), x, act=RELU)
        mid = m.branch[2].out_channels
        if (m.skip is not None and b.tensor_cores and getattr(x.buf, "_rgbd_spare", 0) >= mid and x.coff + x.C + mid <= x.cstride
                and x.dtype == torch.bfloat16 and (x.coff + x.C) % 8 == 0 and m.skip.bias is not None
                and m.branch[4].bias is not None):
            # x sits in a buffer with `mid` spare channels behind it (see _transform): the 3x3 writes t2 there, and the
            # skip conv + last conv + add become ONE 1x1 conv over [x | t2] (K = Cin + mid)
            t2 = b.conv(self._pc(m.branch[2]), t1, out=View(x.buf, x.coff + x.C, mid), act=RELU)
            b.release(t1)
            return b.conv(self._pc_concat(m.skip, m.branch[4]), View(x.buf, x.coff, x.C + mid), out=out)
        t2 = b.conv(self._pc(m.branch[2]), t1, act=RELU)
        b.release(t1)
        if m.skip is not None:
            idn = b.conv(self._pc(m.skip), x, out=out)
            y = b.conv(self._pc(m.branch[4]), t2, out=idn, res=idn)
        else:
            y = b.conv(self._pc(m.branch[4]), t2, out=out, res=x)
        b.release(t2)
        return y

    def _ru(self, b, m, x):
        """AttentionBlock.ResidualUnit (layers.py:178-197): relu(x + conv(x))"""
        pcs = [self._pc(m.conv[i]) for i in (0, 2, 4)]
        if self.fuse_blocks and b.can_fuse_block(*pcs, x):
            return b.fused_block(*pcs, x, res=x, final_relu=True)
        t1 = b.conv(self._pc(m.conv[0]), x, act=RELU)
        t2 = b.conv(self._pc(m.conv[2]), t1, act=RELU)
        b.release(t1)
        y = b.conv(self._pc(m.conv[4]), t2, res=x, act=RELU)
        b.release(t2)
        return y

    def _attention(self, b, m, x, out=None, out_dtype=None):
        """AttentionBlock (layers.py:162-213): x + a(x) * sigmoid(b(x))"""
        a = x
        for i in range(3):
            n = self._ru(b, m.conv_a[i], a)
            if a is not x:
                b.release(a)
            a = n
        t = x
        for i in range(3):
            n = self._ru(b, m.conv_b[i], t)
            if t is not x:
                b.release(t)
            t = n
        y = b.conv(self._pc(m.conv_b[3]), t, out=out, epi=L.EPI_GATE, mul=a, res=x, out_dtype=out_dtype)
        b.release(a, t)
        return y

    def _esa(self, b, m, x, out, res=None):
        """ESA (attention.py:70-97): x * sigmoid(conv4(upsample(small path) + conv_f(conv1 x))) (+ res: the residual
        fusion of STF_united, stf_united.py:494-497)"""
        c1_ = b.conv(self._pc(m.conv1), x)
        c1 = b.conv(self._pc(m.conv2), c1_)
        vmax = b.maxpool7s3(c1)
        b.release(c1)
        vr = b.conv(self._pc(m.conv_max), vmax, act=RELU)
        b.release(vmax)
        c3 = b.conv(self._pc(m.conv3), vr, act=RELU)
        b.release(vr)
        c3b = b.conv(self._pc(m.conv3_), c3)
        b.release(c3)
        s = b.conv(self._pc(m.conv_f), c1_, epi=L.EPI_BILERP, res=c3b)
        b.release(c1_, c3b)
        y = b.conv(self._pc(m.conv4), s, out=out, epi=L.EPI_GATE, mul=x, res=res)
        b.release(s)
        return y

    def _bispf(self, b, m, rgb, depth, rgb_out, depth_out, res_r=None, res_d=None):
        """bi_spf / bi_spf_single (attention.py:14-48). rgb/depth: N-channel views; *_out: the
        N-channel slot after them in the 2N-wide concat buffer (rgb_out None for the single form)."""
        Nh = m.r_ext.out_channels
        e = b.alloc(rgb.N, rgb.H, rgb.W, 3 * Nh)          # [r | d | r]
        b.conv(self._pc(m.r_ext), rgb, out=e.sub(0, Nh), act=RELU, y2=e.sub(2 * Nh, Nh))
        b.conv(self._pc(m.d_ext), depth, out=e.sub(Nh, Nh), act=RELU)
        if rgb_out is not None:
            self._esa(b, m.r_esa, e.sub(0, 2 * Nh), rgb_out, res=res_r)       # ESA(cat(r, d))
        self._esa(b, m.d_esa, e.sub(Nh, 2 * Nh), depth_out, res=res_d)        # ESA(cat(d, r))
        b.release(e)

    def _transform(self, b, rgb_seq, depth_seq, r, d, final_dtype=None):
        """Walks the paired nn.Sequential of g_a / g_s (analysis.py:168-181, synthesis.py:171-184)."""
        n = len(rgb_seq)
        N = self.N
        for i in range(n):
            rm, dm = rgb_seq[i], depth_seq[i]
            last = i == n - 1
            # does a bi_spf follow? then this stage writes into the first half of a 2N buffer
            nxt = rgb_seq[i + 1] if i + 1 < n else None
            feeds_spf = isinstance(nxt, BiSpfSingle)
            if isinstance(rm, BiSpfSingle):
                # r, d are views [0:N] of 2N-wide buffers (depth always; rgb only when bidirectional)
                if isinstance(rm, BiSpf):
                    self._bispf(b, rm, r, d, View(r.buf, N, N), View(d.buf, N, N))
                    r = View(r.buf, 0, 2 * N)
                else:
                    self._bispf(b, rm, r, d, None, View(d.buf, N, N))
                d = View(d.buf, 0, 2 * N)
                continue

            # ... and is the module after that bi_spf a ResidualBottleneck(2N -> N)?  Then the 2N-wide concat buffer gets
            # N / ... spare channels for that block's bottleneck tensor (see _rb: skip conv + last conv as one launch)
            after = rgb_seq[i + 2] if i + 2 < n else None
            spare_r = after.branch[2].out_channels if feeds_spf and isinstance(after, ResidualBottleneck) and after.skip is not None else 0
            after_d = depth_seq[i + 2] if i + 2 < n else None
            spare_d = after_d.branch[2].out_channels if feeds_spf and isinstance(after_d, ResidualBottleneck) and after_d.skip is not None else 0

            def out_for(is_rgb, Hh, Ww, Cc):
                wide = feeds_spf and (self.cross or not is_rgb)
                if wide:
                    spare = (spare_r if is_rgb else spare_d) if b.tensor_cores else 0
                    wide_buf = b.alloc(r.N, Hh, Ww, 2 * Cc + spare).buf
                    wide_buf._rgbd_spare = spare          # channels behind the 2N-wide concat that _rb may use
                    return View(wide_buf, 0, Cc)
                return None

            def step(mod, x, is_rgb):
                if isinstance(mod, (nn.Conv2d, nn.ConvTranspose2d)):
                    s2d = x.C == 12 * mod.in_channels          # image layer on the space-to-depth map
                    pc = self._pc(mod, split3=s2d or x.C == 3 * mod.in_channels, s2d=s2d)
                    Ho, Wo, _ = pc.launches(x.H, x.W)
                    return b.conv(pc, x, out=out_for(is_rgb, Ho, Wo, pc.Cout),
                                  out_dtype=final_dtype if last else None)
                if isinstance(mod, ResidualBottleneck):
                    return self._rb(b, mod, x, out=out_for(is_rgb, x.H, x.W, mod.branch[4].out_channels))
                if isinstance(mod, AttentionBlock):
                    return self._attention(b, mod, x, out=out_for(is_rgb, x.H, x.W, x.C),
                                           out_dtype=final_dtype if last else None)
                raise TypeError(type(mod))

            nr, nd = step(rm, r, True), step(dm, d, False)
            if i > 0:
                b.release(r, d)
            r, d = nr, nd
        return r, d

    def _h_a(self, b, y_r, y_d):
        outs = []
        for seq, y in ((self.h_a.rgb_reduction, y_r), (self.h_a.depth_reduction, y_d)):
            if b.tensor_cores and y.dtype == torch.float32:
                # y stays fp32 for the quantiser; h_a reads a bf16 copy on the tensor cores
                y16 = b.alloc(y.N, y.H, y.W, y.C, torch.bfloat16)
                b.op("rgbd_cast_view_bf16", y.ptr(), y16.ptr(), y.N * y.H * y.W, y.C, y.cstride, y.coff, y16.cstride,
                     y16.coff)
                t1 = b.conv(self._pc(seq[0]), y16, act=RELU)
                b.release(y16)
            else:
                t1 = b.conv(self._pc(seq[0]), y, act=RELU)
            t2 = b.conv(self._pc(seq[2]), t1, act=RELU)
            z = b.conv(self._pc(seq[4]), t2, out_dtype=torch.float32)
            b.release(t1, t2)
            outs.append(z)
        return outs

    def _se_weights(self, se, perm=None):
        w1 = self._dev32(("se1", id(se)), lambda: se.fc[0].weight if perm is None else se.fc[0].weight[:, perm])
        w2 = self._dev32(("se2", id(se)), lambda: se.fc[2].weight if perm is None else se.fc[2].weight[perm, :])
        return w1, w2

    def _hyper_block(self, b, m, x, out=None, y2=None):
        """hyper_transform_block[_single] (synthesis.py:345-380): deconv(SE(x)) (+LeakyReLU)"""
        w1, w2 = self._se_weights(m.se)
        s = b.se_scale(x, w1, w2, plus_one=False)
        return b.conv(self._pc(m.deconv), x, out=out, y2=y2, in_scale=s, act=NONE if m.is_last else LEAKY)

    def _h_s(self, b, zcat, ctx_r, ctx_d, ctx_r2=None):
        """HyperSynthesisEXcross / EXSingle (synthesis.py:305-343).  zcat = [z_r | z_d | z_r] so both
        cat orders are views; the last stage writes straight into the context buffers."""
        Nz, M = self.N, self.M
        hs = self.h_s
        B, h, w = zcat.N, zcat.H, zcat.W
        c1 = b.alloc(B, 2 * h, 2 * w, 3 * M)            # [r1 | d1 | r1]
        c2 = b.alloc(B, 4 * h, 4 * w, 3 * (M * 3 // 2))  # [r2 | d2 | r2]
        M2 = M * 3 // 2
        if self.cross:
            self._hyper_block(b, hs.r_h_s1, zcat.sub(0, 2 * Nz), out=c1.sub(0, M), y2=c1.sub(2 * M, M))
            self._hyper_block(b, hs.d_h_s1, zcat.sub(Nz, 2 * Nz), out=c1.sub(M, M))
            self._hyper_block(b, hs.r_h_s2, c1.sub(0, 2 * M), out=c2.sub(0, M2), y2=c2.sub(2 * M2, M2))
            self._hyper_block(b, hs.d_h_s2, c1.sub(M, 2 * M), out=c2.sub(M2, M2))
            self._hyper_block(b, hs.r_h_s3, c2.sub(0, 2 * M2), out=ctx_r)
            self._hyper_block(b, hs.d_h_s3, c2.sub(M2, 2 * M2), out=ctx_d)
        else:
            self._hyper_block(b, hs.r_h_s1, zcat.sub(0, Nz), out=c1.sub(0, M), y2=c1.sub(2 * M, M))
            self._hyper_block(b, hs.d_h_s1, zcat.sub(Nz, 2 * Nz), out=c1.sub(M, M))
            self._hyper_block(b, hs.r_h_s2, c1.sub(0, M), out=c2.sub(0, M2), y2=c2.sub(2 * M2, M2))
            self._hyper_block(b, hs.d_h_s2, c1.sub(M, 2 * M), out=c2.sub(M2, M2))
            self._hyper_block(b, hs.r_h_s3, c2.sub(0, M2), out=ctx_r, y2=ctx_r2)
            self._hyper_block(b, hs.d_h_s3, c2.sub(M2, 2 * M2), out=ctx_d)
        b.release(c1, c2)

    def _dup(self, b, src, dst):
        b.op("rgbd_copy_view", src.ptr(), dst.ptr(), _DT[src.dtype], src.N * src.H * src.W, src.C, src.cstride,
             src.coff, dst.cstride, dst.coff)

    # ------------------------------------------------------------------ Bi-CEE context model
    def _ctx_layout(self, idx):
        """Channel layout of the shared context buffer for slice idx.
        ours:  [hyper_r 2M | hyper_d 2M | ch_r 2g | ch_d 2g | loc_r 2g | loc_d 2g]   (ch_* only idx>0)
        R2D additionally keeps an rgb-only buffer [hyper_r 2M | ch_r 2g | loc_r 2g] because its rgb
        branch must not see depth (elic_united_R2D.py:82-90).
        Returns (offsets, perms, rgb_offsets): per EntropyParametersEX perm[j] = index into the
        reference's torch.cat order (elic_united.py:288,302,317,333) of our channel j."""
        M, g = self.M, self.slice_ch[idx]
        has_ch = idx > 0
        o = {"hyper_r": 0, "hyper_d": 2 * M}
        p = 4 * M
        if has_ch:
            o["ch_r"], o["ch_d"] = p, p + 2 * g
            p += 4 * g
        o["loc_r"], o["loc_d"] = p, p + 2 * g
        o["end"] = p + 4 * g

        def perm(ref_order, ours):
            start, q = {}, 0
            for name, width in ref_order:
                start[name] = q
                q += width
            idxs = []
            for name, width in ours:
                idxs.extend(range(start[name], start[name] + width))
            assert len(idxs) == q
            return torch.tensor(idxs, dtype=torch.long, device="cpu")

        base = [("hyper_r", 2 * M), ("hyper_d", 2 * M)] + ([("ch_r", 2 * g), ("ch_d", 2 * g)] if has_ch else [])
        lr, ld = ("loc_r", 2 * g), ("loc_d", 2 * g)
        plans = {
            "d_anchor": perm([lr] + base, base + [lr]),
            "d_nonanchor": perm([lr, ld] + base, base + [lr, ld]),
        }
        o_r = None
        if self.cross:
            plans["r_anchor"] = perm(base, base)
            plans["r_nonanchor"] = perm([lr, ld] + base, base + [lr, ld])
        else:
            rbase = [("hyper_r", 2 * M)] + ([("ch_r", 2 * g)] if has_ch else [])
            plans["r_anchor"] = perm(rbase, rbase)
            plans["r_nonanchor"] = perm([lr] + rbase, rbase + [lr])
            o_r = {"hyper_r": 0, "ch_r": 2 * M, "loc_r": 2 * M + (2 * g if has_ch else 0)}
        return o, plans, o_r

    def _ep(self, b, m, x, perm, out):
        """EntropyParametersEX (entropy.py:56-78): fusion(x + se(x)) -> fp32 (scales | means)"""
        w1, w2 = self._se_weights(m.se, perm)
        s = b.se_scale(x, w1, w2, plus_one=True)
        t1 = b.conv(self._pc(m.fusion[0], in_perm=perm), x, in_scale=s, act=RELU)
        t2 = b.conv(self._pc(m.fusion[2]), t1, act=RELU)
        b.release(t1)
        y = b.conv(self._pc(m.fusion[4]), t2, out=out)
        b.release(t2)
        return y

    def _channel_ctx(self, b, m, x, out, y2=None):
        t1 = b.conv(self._pc(m.fushion[0]), x, act=RELU)
        t2 = b.conv(self._pc(m.fushion[2]), t1, act=RELU)
        b.release(t1)
        b.conv(self._pc(m.fushion[4]), t2, out=out, y2=y2)
        b.release(t2)

    def _context_chain(self, b, ctx, ctx_rgb, yhat_r, yhat_d, code_step):
        """The 20-stage serial chain (elic_united.py:265-348 / 454-541; R2D: elic_united_R2D.py:73-326).
        code_step(which, idx, parity, params_view, g, coff) emits the kernels that turn Gaussian
        params into y_hat at the parity sites (quantise in the encoder, rANS-decode in the decoder,
        ste + likelihood in forward).  ctx_rgb is the rgb-only context buffer of the R2D variant."""
        cross = self.cross
        if b.tensor_cores:
            # one scratch buffer for the per-image SE-folded 1x1 filters of every EntropyParametersEX stage
            eps = [m for lst in (self.rgb_entropy_parameters_anchor, self.depth_entropy_parameters_anchor,
                                 self.rgb_entropy_parameters_nonanchor, self.depth_entropy_parameters_nonanchor)
                   for m in lst]
            b.reserve_wscratch(ctx.N * max((m.fusion[0].out_channels + 15) // 16 * 16 *
                                           ((m.fusion[0].in_channels + 63) // 64 * 64) for m in eps))
        for idx, g in enumerate(self.slice_ch):
            coff = sum(self.slice_ch[:idx])
            o, perms, o_r = self._ctx_layout(idx)
            params = b.alloc(ctx.N, ctx.H, ctx.W, 2 * g, torch.float32)

            def rgb_copy(name):
                return None if cross else ctx_rgb.sub(o_r[name], 2 * g)

            if idx > 0:
                self._channel_ctx(b, self.rgb_channel_context[idx], yhat_r.sub(0, coff), ctx.sub(o["ch_r"], 2 * g),
                                  y2=rgb_copy("ch_r"))
                self._channel_ctx(b, self.depth_channel_context[idx], yhat_d.sub(0, coff), ctx.sub(o["ch_d"], 2 * g))
            pre = o["loc_r"]
            # (1) rgb anchor
            x = ctx.sub(0, pre) if cross else ctx_rgb.sub(0, o_r["loc_r"])
            self._ep(b, self.rgb_entropy_parameters_anchor[idx], x, perms["r_anchor"], params)
            code_step("r", idx, 0, params, g, coff)
            b.conv(self._pc(self.rgb_local_context[idx]), yhat_r.sub(coff, g), out=ctx.sub(o["loc_r"], 2 * g),
                   y2=rgb_copy("loc_r"))
            # (2) depth anchor
            self._ep(b, self.depth_entropy_parameters_anchor[idx], ctx.sub(0, pre + 2 * g), perms["d_anchor"], params)
            code_step("d", idx, 0, params, g, coff)
            b.conv(self._pc(self.depth_local_context[idx]), yhat_d.sub(coff, g), out=ctx.sub(o["loc_d"], 2 * g))
            # (3) rgb non-anchor
            x = ctx.sub(0, pre + 4 * g) if cross else ctx_rgb.sub(0, o_r["loc_r"] + 2 * g)
            self._ep(b, self.rgb_entropy_parameters_nonanchor[idx], x, perms["r_nonanchor"], params)
            code_step("r", idx, 1, params, g, coff)
            b.conv(self._pc(self.rgb_local_context_anchor_with_nonanchor[idx]), yhat_r.sub(coff, g),
                   out=ctx.sub(o["loc_r"], 2 * g))
            # (4) depth non-anchor
            self._ep(b, self.depth_entropy_parameters_nonanchor[idx], ctx.sub(0, pre + 4 * g), perms["d_nonanchor"], params)
            code_step("d", idx, 1, params, g, coff)
            b.release(params)

    def _ctx_buffers(self, b, B, h, w):
        M, gm = self.M, max(self.slice_ch)
        ctx = b.alloc(B, h, w, 4 * M + 8 * gm)
        ctx_rgb = None if self.cross else b.alloc(B, h, w, 2 * M + 4 * gm)
        # the SE gates of the 20 EntropyParametersEX stages read growing prefixes of these buffers: keep the per-channel
        # sums and refresh only what a stage's producers rewrote (the 2M / 4M hyper-prior channels never change)
        b.se_cache(ctx)
        if ctx_rgb is not None:
            b.se_cache(ctx_rgb)
        return ctx, ctx_rgb

    # ------------------------------------------------------------------ programs
    def _gc(self, which):
        return self.rgb_gaussian_conditional if which == "r" else self.depth_gaussian_conditional

    def _eb(self, which):
        return self.rgb_entropy_bottleneck if which == "r" else self.depth_entropy_bottleneck

    def _chunk_offsets(self, h, w):
        """Symbol offsets of the 10 chunks inside one image's y stream."""
        offs, p = {}, 0
        for idx, g in enumerate(self.slice_ch):
            for parity in (0, 1):
                offs[(idx, parity)] = p
                p += g * h * (w // 2)
        return offs, p

    def _sub_streams(self, h, w, ny, sub_channels=None):
        """(sub-streams per image and modality, symbols per sub-stream) of the y latents."""
        if self.stream_layout == "single" and sub_channels is None:
            return 1, ny
        c = int(sub_channels or self.sub_channels)
        if c < 1 or any(g % c for g in self.slice_ch):
            raise ValueError(f"sub_channels = {c} must divide every channel group {self.slice_ch}")
        sublen = c * h * (w // 2)
        return ny // sublen, sublen

    def _analysis(self, b, B, H, W):
        """images -> g_a.  Registers the fp32 NCHW input buffers in the program and returns the latents (fp32 views)."""
        p = b.prog
        # bf16 tensor-core mode: the images enter as a two-term bf16 expansion [hi | lo | hi] so the first
        # layer keeps fp32-like accuracy (16-bit depth is represented exactly) at no extra MMA cost
        # ... and as a space-to-depth map (2x2 pixel blocks -> channels), which turns the 5x5 stride-2 first conv
        # into a 3x3 stride-1 conv with a single tap group (one halo load per tile instead of four parity loads)
        split = 2 if b.tensor_cores else 0
        b.stage = "io"
        x_r = b.alloc(B, H // 2, W // 2, 36) if split else b.alloc(B, H, W, 3)
        x_d = b.alloc(B, H // 2, W // 2, 12) if split else b.alloc(B, H, W, 1)
        in_r = b.raw((B, 3, H, W), torch.float32)
        in_d = b.raw((B, 1, H, W), torch.float32)
        p.io["rgb"], p.io["depth"] = in_r, in_d
        b.op("rgbd_nchw_to_nhwc", in_r.data_ptr(), x_r.ptr(), _DT[x_r.dtype], B, 3, H, W, x_r.cstride, x_r.coff, split)
        b.op("rgbd_nchw_to_nhwc", in_d.data_ptr(), x_d.ptr(), _DT[x_d.dtype], B, 1, H, W, x_d.cstride, x_d.coff, split)
        b.stage = "g_a"
        return self._transform(b, self.g_a.rgb_analysis_transform, self.g_a.depth_analysis_transform, x_r, x_d,
                               final_dtype=torch.float32)

    def _synthesis(self, b, yhat_r, yhat_d):
        """y_hat -> g_s: the reconstructions as NHWC views (3 and 1 channels)."""
        return self._transform(b, self.g_s.rgb_synthesis_transform, self.g_s.depth_synthesis_transform, yhat_r, yhat_d)

    def _common_front(self, b, B, H, W):
        """image -> g_a -> h_a. Returns io views."""
        y_r, y_d = self._analysis(b, B, H, W)
        b.stage = "h_a"
        z_r, z_d = self._h_a(b, y_r, y_d)
        b.stage = "coder"
        return y_r, y_d, z_r, z_d

    def _tables(self, which_model, which):
        m = self._gc(which) if which_model == "gc" else self._eb(which)
        return m.device_tables(self.device)

    def _build_encoder(self, B, H, W):
        b = Builder(self.device, self.act_dtype, self.tensor_cores and self.precision == "bf16")
        p = b.prog
        y_r, y_d, z_r, z_d = self._common_front(b, B, H, W)
        h, w = y_r.H, y_r.W
        hz, wz = z_r.H, z_r.W
        Nz, M = self.N, self.M
        nz = Nz * hz * wz
        zcat = b.alloc(B, hz, wz, 3 * Nz)
        ys = {"r": y_r, "d": y_d}
        zs = {"r": z_r, "d": z_d}
        offs, ny = self._chunk_offsets(h, w)
        st = {}
        # both modalities' y streams live in one [2B, ...] buffer so one launch can code all of them
        # y streams per image and modality: 1 (reference layout) or ny / sublen equal sub-streams (multi-stream layout)
        nsub, sublen = self._sub_streams(h, w, ny)
        ycap = sublen + sublen // 2 + 64
        zcap = nz + nz // 2 + 64
        ysym_all, yidx_all = b.raw((2 * B, ny), torch.int32), b.raw((2 * B, ny), torch.uint8)
        yout_all, zout_all = b.raw((2 * B * nsub, ycap), torch.int32), b.raw((2 * B, zcap), torch.int32)
        # word counts of all streams in gather order [y_r | y_d | z_r | z_d] (rgbd_gather_streams)
        counts = b.raw((2 * B * nsub + 2 * B,), torch.int32)
        ynw_all = counts[:2 * B * nsub]
        for k, which in enumerate(("r", "d")):
            eb = self._eb(which)
            med = self._dev32(("med", which), eb.medians)
            s = dict(
                zsym=b.raw((B, nz), torch.int32), zidx=b.raw((B, nz), torch.uint8),
                ysym=ysym_all[k * B:(k + 1) * B], yidx=yidx_all[k * B:(k + 1) * B],
                zcap=zcap, ycap=ycap)
            s["zout"] = zout_all[k * B:(k + 1) * B]
            s["yout"] = yout_all[k * B * nsub:(k + 1) * B * nsub]
            s["znw"] = counts[2 * B * nsub + k * B:2 * B * nsub + (k + 1) * B]
            s["ynw"] = ynw_all[k * B * nsub:(k + 1) * B * nsub]
            st[which] = s
            zh = zcat.sub(0 if which == "r" else Nz, Nz)
            b.op("rgbd_eb_quantize", zs[which].ptr(), zs[which].cstride, B, hz * wz, Nz, med.data_ptr(),
                 s["zsym"].data_ptr(), s["zidx"].data_ptr(), zh.ptr(), _DT[zh.dtype], zh.cstride, zh.coff)
            t = self._tables("eb", which)
            b.op("rgbd_rans_encode", s["zsym"].data_ptr(), s["zidx"].data_ptr(), nz, nz, B, ctypes.byref(t.struct),
                 s["zout"].data_ptr(), s["zcap"], s["znw"].data_ptr())
            p.keep.append(t)
        self._dup(b, zcat.sub(0, Nz), zcat.sub(2 * Nz, Nz))
        ctx, ctx_rgb = self._ctx_buffers(b, B, h, w)
        b.stage = "h_s"
        self._h_s(b, zcat, ctx.sub(0, 2 * M), ctx.sub(2 * M, 2 * M), None if ctx_rgb is None else ctx_rgb.sub(0, 2 * M))
        b.stage = "chain"
        yhat = {"r": b.alloc(B, h, w, M, zero=True), "d": b.alloc(B, h, w, M, zero=True)}
        for which in ("r", "d"):
            t = yhat[which].buf
            b.op("rgbd_zero", t.data_ptr(), t.numel() * t.element_size())
        table = {k: self._dev32(("scale_table", k), lambda k=k: self._gc(k).scale_table) for k in ("r", "d")}
        bound = {k: float(self._gc(k).lower_bound_scale.bound) for k in ("r", "d")}

        def code_step(which, idx, parity, params, g, coff):
            s = st[which]
            b.stage = "coder"
            b.op("rgbd_ckbd_quantize_index", ys[which].ptr(), ys[which].cstride, ys[which].coff + coff,
                 params.ptr(), table[which].data_ptr(), table[which].numel(), bound[which], B, h, w, g, parity,
                 s["ysym"].data_ptr(), s["yidx"].data_ptr(), ny, offs[(idx, parity)],
                 yhat[which].ptr(), _DT[yhat[which].dtype], yhat[which].cstride, coff)
            b.stage = "chain"

        self._context_chain(b, ctx, ctx_rgb, yhat["r"], yhat["d"], code_step)
        b.stage = "coder"
        gr, gd = self._gc("r"), self._gc("d")
        same_tables = all(torch.equal(getattr(gr, n), getattr(gd, n)) for n in ("_quantized_cdf", "_cdf_length", "_offset"))
        if same_tables:   # the usual case: both Gaussian conditionals use get_scale_table()
            t = self._tables("gc", "r")
            # (the symbols of all images and both modalities are contiguous: 2B * nsub streams of sublen symbols)
            b.op("rgbd_rans_encode", ysym_all.data_ptr(), yidx_all.data_ptr(), sublen, sublen, 2 * B * nsub,
                 ctypes.byref(t.struct), yout_all.data_ptr(), ycap, ynw_all.data_ptr())
            p.keep.append(t)
        else:
            for which in ("r", "d"):
                s = st[which]
                t = self._tables("gc", which)
                b.op("rgbd_rans_encode", s["ysym"].data_ptr(), s["yidx"].data_ptr(), sublen, sublen, B * nsub,
                     ctypes.byref(t.struct), s["yout"].data_ptr(), s["ycap"], s["ynw"].data_ptr())
                p.keep.append(t)
        # all stream tails packed into one pinned host buffer by one kernel: [4B counts | words ...]; sized for
        # 8 bits per symbol on average (a longer job falls back to per-stream copies in _collect_strings)
        gcap = 2 * B * (ny + nz) // 4 + 64 + 4 * B * nsub
        p.io.update(st=st, shape=(hz, wz), y=ys, z=zs, yhat=yhat, ny=ny, nz=nz, counts=counts, gather_cap=gcap, nsub=nsub,
                    sublen=sublen, n_streams=2 * B * nsub + 2 * B,
                    gather_args=(yout_all.data_ptr(), ycap, 2 * B * nsub, zout_all.data_ptr(), zcap, 2 * B,
                                 counts.data_ptr()))
        return p

    def _build_decoder(self, B, hz, wz, sub_channels=0):
        """sub_channels = 0: the reference's single-stream layout; else the multi-stream layout with that many channels
        per sub-stream."""
        b = Builder(self.device, self.act_dtype, self.tensor_cores and self.precision == "bf16")
        p = b.prog
        b.stage = "coder"
        Nz, M = self.N, self.M
        h, w = hz * 4, wz * 4
        H, W = h * 16, w * 16
        nz = Nz * hz * wz
        offs, ny = self._chunk_offsets(h, w)
        nsub, sublen = self._sub_streams(h, w, ny, sub_channels) if sub_channels else (1, ny)
        zcat = b.alloc(B, hz, wz, 3 * Nz)
        # stream table: the order is [z_r | z_d | y_r | y_d]; z: B streams per modality, y: B * nsub
        n_streams = 2 * B + 2 * B * nsub
        words_cap = 2 * B * ((nz + nz // 2 + 64) + nsub * (sublen + sublen // 2 + 64))
        words = b.raw((words_cap,), torch.int32)
        word_off = b.raw((n_streams,), torch.int64)
        word_len = b.raw((n_streams,), torch.int64)
        state = b.raw((n_streams, 2), torch.int64)
        p.io.update(words=words, word_off=word_off, word_len=word_len, words_cap=words_cap, state=state, nsub=nsub,
                    n_streams=n_streams)
        b.op("rgbd_rans_decode_init", words.data_ptr(), word_off.data_ptr(), n_streams, state.data_ptr())
        st = {}
        for k, which in enumerate(("r", "d")):
            eb = self._eb(which)
            med = self._dev32(("med", which), eb.medians)
            s = dict(zsym=b.raw((B, nz), torch.int32), zidx=b.raw((B, nz), torch.uint8),
                     ysym=b.raw((B, ny), torch.int32), yidx=b.raw((B, ny), torch.uint8),
                     zslot=k * B, yslot=2 * B + k * B * nsub)
            st[which] = s
            chan = torch.arange(Nz, device=self.device, dtype=torch.uint8).repeat_interleave(hz * wz).repeat(B, 1)
            s["zidx"].copy_(chan)   # EntropyBottleneck._build_indexes (entropy_models.py:430-435)
            t = self._tables("eb", which)
            b.op("rgbd_rans_decode_chunk", words.data_ptr(), word_off[s["zslot"]:].data_ptr(),
                 word_len[s["zslot"]:].data_ptr(), B, state[s["zslot"]:].data_ptr(), s["zidx"].data_ptr(),
                 s["zsym"].data_ptr(), nz, 0, nz, ctypes.byref(t.struct))
            p.keep.append(t)
            zh = zcat.sub(0 if which == "r" else Nz, Nz)
            b.op("rgbd_eb_dequantize", s["zsym"].data_ptr(), B, hz * wz, Nz, med.data_ptr(), zh.ptr(),
                 _DT[zh.dtype], zh.cstride, zh.coff)
        self._dup(b, zcat.sub(0, Nz), zcat.sub(2 * Nz, Nz))
        ctx, ctx_rgb = self._ctx_buffers(b, B, h, w)
        b.stage = "h_s"
        self._h_s(b, zcat, ctx.sub(0, 2 * M), ctx.sub(2 * M, 2 * M), None if ctx_rgb is None else ctx_rgb.sub(0, 2 * M))
        b.stage = "chain"
        yhat = {"r": b.alloc(B, h, w, M, zero=True), "d": b.alloc(B, h, w, M, zero=True)}
        for which in ("r", "d"):
            t = yhat[which].buf
            b.op("rgbd_zero", t.data_ptr(), t.numel() * t.element_size())
        table = {k: self._dev32(("scale_table", k), lambda k=k: self._gc(k).scale_table) for k in ("r", "d")}
        bound = {k: float(self._gc(k).lower_bound_scale.bound) for k in ("r", "d")}

        def code_step(which, idx, parity, params, g, coff):
            s = st[which]
            n = g * h * (w // 2)
            off = offs[(idx, parity)]
            t = self._tables("gc", which)
            b.stage = "coder"
            b.op("rgbd_ckbd_index", params.ptr(), table[which].data_ptr(), table[which].numel(), bound[which],
                 B, h, w, g, parity, s["yidx"].data_ptr(), ny, off)
            if nsub == 1:
                b.op("rgbd_rans_decode_chunk", words.data_ptr(), word_off[s["yslot"]:].data_ptr(),
                     word_len[s["yslot"]:].data_ptr(), B, state[s["yslot"]:].data_ptr(), s["yidx"].data_ptr(),
                     s["ysym"].data_ptr(), ny, off, n, ctypes.byref(t.struct))
            else:
                # this step's n / sublen sub-streams of every image, all concurrently (fresh decoder state each)
                b.op("rgbd_rans_decode_streams", words.data_ptr(), word_off.data_ptr(), word_len.data_ptr(), B, n // sublen,
                     s["yslot"] + off // sublen, nsub, state.data_ptr(), s["yidx"].data_ptr(), s["ysym"].data_ptr(), ny,
                     sublen, off, sublen, ctypes.byref(t.struct))
            b.op("rgbd_ckbd_dequant_scatter", s["ysym"].data_ptr(), ny, off, params.ptr(), B, h, w, g, parity,
                 yhat[which].ptr(), _DT[yhat[which].dtype], yhat[which].cstride, coff)
            p.keep.append(t)
            b.stage = "chain"

        self._context_chain(b, ctx, ctx_rgb, yhat["r"], yhat["d"], code_step)
        b.stage = "g_s"
        x_r, x_d = self._synthesis(b, yhat["r"], yhat["d"])
        b.stage = "io"
        out_r = b.raw((B, 3, H, W), torch.float32)
        out_d = b.raw((B, 1, H, W), torch.float32)
        b.op("rgbd_nhwc_to_nchw", x_r.ptr(), _DT[x_r.dtype], out_r.data_ptr(), B, 3, H, W, x_r.cstride, x_r.coff, 1)
        b.op("rgbd_nhwc_to_nchw", x_d.ptr(), _DT[x_d.dtype], out_d.data_ptr(), B, 1, H, W, x_d.cstride, x_d.coff, 1)
        # (x_nhwc: the synthesis transform's output before the [0, 1] clamp of decompress(), for the parity tests)
        p.io.update(st=st, out_r=out_r, out_d=out_d, yhat=yhat, ny=ny, nz=nz, x_nhwc={"r": x_r, "d": x_d})
        return p

    def _build_forward(self, B, H, W):
        b = Builder(self.device, self.act_dtype, self.tensor_cores and self.precision == "bf16")
        p = b.prog
        y_r, y_d, z_r, z_d = self._common_front(b, B, H, W)
        h, w = y_r.H, y_r.W
        hz, wz = z_r.H, z_r.W
        Nz, M = self.N, self.M
        zcat = b.alloc(B, hz, wz, 3 * Nz)
        ys = {"r": y_r, "d": y_d}
        zs = {"r": z_r, "d": z_d}
        lik = {}
        for which in ("r", "d"):
            eb = self._eb(which)
            ebp = eb.packed_params(self.device)
            p.keep.append(ebp)
            lz = b.raw((B, Nz, hz, wz), torch.float32)
            ly = b.raw((B, M, h, w), torch.float32)
            lik[which] = (ly, lz)
            zh = zcat.sub(0 if which == "r" else Nz, Nz)
            b.op("rgbd_eb_likelihood", zs[which].ptr(), zs[which].cstride, B, hz * wz, Nz, ebp.data_ptr(),
                 eb.likelihood_bound, zh.ptr(), _DT[zh.dtype], zh.cstride, zh.coff, lz.data_ptr())
        self._dup(b, zcat.sub(0, Nz), zcat.sub(2 * Nz, Nz))
        ctx, ctx_rgb = self._ctx_buffers(b, B, h, w)
        b.stage = "h_s"
        self._h_s(b, zcat, ctx.sub(0, 2 * M), ctx.sub(2 * M, 2 * M), None if ctx_rgb is None else ctx_rgb.sub(0, 2 * M))
        b.stage = "chain"
        yhat = {"r": b.alloc(B, h, w, M, zero=True), "d": b.alloc(B, h, w, M, zero=True)}
        for which in ("r", "d"):
            t = yhat[which].buf
            b.op("rgbd_zero", t.data_ptr(), t.numel() * t.element_size())
        bound = {k: float(self._gc(k).lower_bound_scale.bound) for k in ("r", "d")}

        def code_step(which, idx, parity, params, g, coff):
            gc = self._gc(which)
            b.op("rgbd_ckbd_ste_likelihood", ys[which].ptr(), ys[which].cstride, ys[which].coff + coff,
                 params.ptr(), bound[which], gc.likelihood_bound, B, h, w, g, parity, yhat[which].ptr(),
                 _DT[yhat[which].dtype], yhat[which].cstride, coff, lik[which][0].data_ptr(), M, coff)

        self._context_chain(b, ctx, ctx_rgb, yhat["r"], yhat["d"], code_step)
        b.stage = "g_s"
        x_r, x_d = self._synthesis(b, yhat["r"], yhat["d"])
        b.stage = "io"
        out_r = b.raw((B, 3, H, W), torch.float32)
        out_d = b.raw((B, 1, H, W), torch.float32)
        b.op("rgbd_nhwc_to_nchw", x_r.ptr(), _DT[x_r.dtype], out_r.data_ptr(), B, 3, H, W, x_r.cstride, x_r.coff, 0)
        b.op("rgbd_nhwc_to_nchw", x_d.ptr(), _DT[x_d.dtype], out_d.data_ptr(), B, 1, H, W, x_d.cstride, x_d.coff, 0)
        p.io.update(out_r=out_r, out_d=out_d, lik=lik, y=ys, z=zs, yhat=yhat)
        return p

    def _program(self, kind, *dims, slot=0):
        if kind == "decoder" and len(dims) == 3:
            dims = tuple(dims) + (0,)          # (B, hz, wz, sub_channels): 0 = the reference's single-stream layout
        layout = (self.stream_layout, self.sub_channels) if kind == "encoder" else None
        key = (kind, self.precision, self.tensor_cores, self.fuse_blocks, layout, slot) + tuple(dims)
        if key not in self._programs:
            self._require_cuda()
            with torch.cuda.device(self.device), torch.no_grad():
                self._programs[key] = getattr(self, "_build_" + kind)(*dims)
        return self._programs[key]

    def _require_cuda(self):
        L.load()  # raises if the CUDA extension is missing — there is no fallback path
        if self.device.type != "cuda":
            raise L.RgbdError("ELIC_united runs on a CUDA device only (no CPU fallback); call .to('cuda')")
        if self.rgb_gaussian_conditional.quantized_cdf.numel() == 0:
            raise ValueError("Uninitialized CDFs. Run update() first")

    def _check_inputs(self, rgb, depth):
        if rgb.dim() != 4 or depth.dim() != 4 or rgb.shape[1] != 3 or depth.shape[1] != 1:
            raise ValueError("expected rgb [B,3,H,W] and depth [B,1,H,W]")
        if rgb.shape[0] != depth.shape[0] or rgb.shape[2:] != depth.shape[2:]:
            raise ValueError("rgb and depth must have the same batch and spatial size")
        if rgb.shape[2] % 64 or rgb.shape[3] % 64:
            raise ValueError("H and W must be multiples of 64 (pad first, dataset/utils.py:58-67)")
        if rgb.shape[2] < 128 or rgb.shape[3] < 128:
            raise ValueError("H and W must be >= 128 (ESA max_pool2d(7, 3) on the 1/8-scale map)")

    # ------------------------------------------------------------------ public API
    @torch.no_grad()
    def forward(self, rgb, depth):
        if self.quant != "ste":
            # codeOnePart (models/elic_united.py:99-103) quantises without the means for any other setting
            raise NotImplementedError(f"forward() implements quant='ste' only (got {self.quant!r}); "
                                      "compress()/decompress() do not depend on it")
        self._check_inputs(rgb, depth)
        B, _, H, W = rgb.shape
        p = self._program("forward", B, H, W)
        with torch.cuda.device(self.device):
            p.io["rgb"].copy_(rgb)
            p.io["depth"].copy_(depth)
            p.run(self.use_cuda_graph)
        lik = p.io["lik"]
        return {
            "x_hat": {"r": p.io["out_r"].clone(), "d": p.io["out_d"].clone()},
            "r_likelihoods": {"y": lik["r"][0].clone(), "z": lik["r"][1].clone()},
            "d_likelihoods": {"y": lik["d"][0].clone(), "z": lik["d"][1].clone()},
        }

    @torch.no_grad()
    def compress(self, rgb, depth):
        return self.compress_async(rgb, depth).result()

    @torch.no_grad()
    def compress_async(self, rgb, depth, slot=0):
        """Enqueue one compress() on pipeline slot `slot` (its own program instance and CUDA stream)
        and return a handle; `.result()` waits for that slot only and returns the compress() dict.
        Several slots in flight let the serial rANS kernels of one batch overlap the convolutions of
        another (images are independent: SURVEY §8e)."""
        self._check_inputs(rgb, depth)
        B, _, H, W = rgb.shape
        p = self._program("encoder", B, H, W, slot=slot)
        stream = self._slot_stream(slot)
        self._order_after_producer(stream, rgb, depth)
        with torch.cuda.device(self.device), torch.cuda.stream(stream):
            for name, src in (("rgb", rgb), ("depth", depth)):
                if src.is_cuda or B == 1:
                    p.io[name].copy_(src, non_blocking=True)
                else:
                    # host images: one H2D copy per image, so that a big batch does not hold the copy engine for
                    # milliseconds while other pipeline slots' small transfers (stream words, decoder states) wait behind it
                    for i in range(B):
                        p.io[name][i].copy_(src[i], non_blocking=True)
            p.run(self.use_cuda_graph)
            if "gather_host" not in p.io:
                p.io["gather_host"] = torch.empty((p.io["n_streams"] + p.io["gather_cap"],), dtype=torch.int32,
                                                  device="cpu", pin_memory=True)
            gh = p.io["gather_host"]
            # one kernel packs the counts and every stream's tail straight into pinned host memory (zero-copy stores over
            # PCIe): the strings are complete on the host when `done` fires.  Deliberately NOT a DMA copy: the copy engine
            # serves its queue in order, and behind other slots' bulk transfers (a job's reconstruction is 126 MB) this small,
            # latency-critical transfer waited milliseconds (measured on B200: 304 -> 259 pairs/s end to end).
            L.call("rgbd_gather_streams", *p.io["gather_args"], gh.data_ptr(), p.io["gather_cap"],
                   ctypes.c_void_p(stream.cuda_stream))
            done = torch.cuda.Event()
            done.record(stream)
        return _CompressHandle(self, p, B, stream, done, gh)

    def _order_after_producer(self, stream, *tensors):
        """A slot stream reads tensors the caller produced on ITS current stream: wait for that stream, and tell the
        caching allocator the slot stream uses the storage (it must not be recycled while the copy is pending)."""
        cur = torch.cuda.current_stream(self.device)
        if stream != cur:
            stream.wait_stream(cur)
            for t in tensors:
                if t.is_cuda:
                    t.record_stream(stream)

    def _slot_stream(self, slot):
        if slot == 0:
            return torch.cuda.current_stream(self.device)
        streams = self.__dict__.setdefault("_slot_streams", {})
        if slot not in streams:
            streams[slot] = torch.cuda.Stream(self.device)
        return streams[slot]

    def _collect_strings(self, p, B, gh):
        """Cut the packed host buffer [4B counts | words] (rgbd_gather_streams) into the per-stream strings."""
        st = p.io["st"]
        host = gh.numpy()
        ns, nsub = p.io["n_streams"], p.io["nsub"]
        counts = host[:ns].copy()
        if (counts < 0).any():
            raise L.RgbdError("rANS output buffer overflow (stream longer than 48 bits/symbol)")
        res = {"ry": [], "rz": [], "dy": [], "dz": []}
        order = [(k, i) for k in ("ry", "dy") for i in range(B * nsub)] + [(k, i) for k in ("rz", "dz") for i in range(B)]
        if int(counts.sum()) <= p.io["gather_cap"]:
            words = host[ns:]
            pos = 0
            for (key, _), n in zip(order, counts):
                res[key].append(words[pos:pos + n].tobytes())
                pos += int(n)
            return res
        # rare: more than 8 bits per symbol on average — the packed buffer was too small, fetch stream by stream
        for (key, i), n in zip(order, counts):
            s_ = st[key[0]]
            out, cap = s_[key[1] + "out"], s_[key[1] + "cap"]
            res[key].append(out[i, cap - int(n):].cpu().numpy().tobytes())
        return res

    @torch.no_grad()
    def decompress(self, rgb_strings, depth_strings, shape):
        self._require_cuda()
        torch.cuda.synchronize(self.device)
        t0 = time.process_time()
        out = self.decompress_async(rgb_strings, depth_strings, shape).result()
        torch.cuda.synchronize(self.device)
        out["cost_time"] = time.process_time() - t0
        return out

    @torch.no_grad()
    def decompress_async(self, rgb_strings, depth_strings, shape, slot=0):
        """Enqueue one decompress() on pipeline slot `slot`; `.result()` returns {"x_hat": ...}."""
        self._require_cuda()
        rz, dz = list(rgb_strings[1]), list(depth_strings[1])
        B = len(rz)
        ry, dy = list(rgb_strings[0]), list(depth_strings[0])
        if B == 0 or len(dz) != B or len(ry) != len(dy) or len(ry) % B:
            raise ValueError(f"expected the same number of y strings per modality, a multiple of the {B} z strings; "
                             f"got {len(ry)} / {len(dy)}")
        hz, wz = int(shape[0]), int(shape[1])
        if not (2 <= hz <= 1024 and 2 <= wz <= 1024):
            raise ValueError(f"latent shape {(hz, wz)} out of range (images of 128 .. 65536 pixels per side)")
        # one y string per image = the reference's layout; more = the multi-stream layout, whose sub-stream size follows
        # from the count (2 M / sub_channels strings per image: both checkerboard halves of every channel group)
        sub_channels = 0
        if len(ry) != B:
            per_image = len(ry) // B
            if (2 * self.M) % per_image:
                raise ValueError(f"{per_image} y strings per image do not form a multi-stream layout of {self.M} channels")
            sub_channels = 2 * self.M // per_image
        p = self._program("decoder", B, hz, wz, sub_channels, slot=slot)
        streams = rz + dz + ry + dy
        lens = np.array([len(s) // 4 for s in streams], dtype=np.int64)
        if any(len(s) % 4 or len(s) < 8 for s in streams):
            raise ValueError("corrupt stream: rANS payloads are whole 32-bit words, at least two")
        offs = np.zeros_like(lens)
        offs[1:] = np.cumsum(lens[:-1])
        total = int(lens.sum())
        if total > p.io["words_cap"]:
            raise ValueError("streams larger than the decoder's word buffer")
        if "words_host" not in p.io:   # pinned staging so the H2D copies are asynchronous
            p.io["words_host"] = torch.empty(p.io["words_cap"], dtype=torch.int32, device="cpu", pin_memory=True)
            p.io["meta_host"] = torch.empty((2, p.io["n_streams"]), dtype=torch.int64, device="cpu", pin_memory=True)
        stream = self._slot_stream(slot)
        with torch.cuda.device(self.device), torch.cuda.stream(stream):
            prev = p.io.get("h2d_done")
            if prev is not None:
                prev.synchronize()     # the staging buffers are free again
            wh = p.io["words_host"].numpy()
            pos = 0
            for sbytes in streams:
                n = len(sbytes) // 4
                wh[pos:pos + n] = np.frombuffer(sbytes, dtype=np.int32)
                pos += n
            p.io["meta_host"][0].copy_(torch.from_numpy(offs))
            p.io["meta_host"][1].copy_(torch.from_numpy(lens))
            p.io["words"][:total].copy_(p.io["words_host"][:total], non_blocking=True)
            p.io["word_off"].copy_(p.io["meta_host"][0], non_blocking=True)
            p.io["word_len"].copy_(p.io["meta_host"][1], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(stream)
            p.io["h2d_done"] = ev
            p.run(self.use_cuda_graph)
            if "state_host" not in p.io:
                p.io["state_host"] = torch.empty((p.io["n_streams"], 2), dtype=torch.int64, device="cpu", pin_memory=True)
            # decoder end states -> pinned host memory by a tiny kernel (zero-copy stores), not by the copy engine: see
            # compress_async — `done` must not wait behind other slots' bulk D2H
            L.call("rgbd_copy_view", p.io["state"].data_ptr(), p.io["state_host"].data_ptr(), L.DT_F32, p.io["n_streams"], 4, 4, 0,
                   4, 0, ctypes.c_void_p(stream.cuda_stream))
            done = torch.cuda.Event()
            done.record(stream)
        return _DecompressHandle(p, stream, done, lens)


class _CompressHandle:
    def __init__(self, net, prog, B, stream, done, gather_host):
        self.net, self.prog, self.B, self.stream, self.done, self.gather_host = net, prog, B, stream, done, gather_host

    def result(self):
        self.done.synchronize()
        p = self.prog
        with torch.cuda.device(self.net.device), torch.cuda.stream(self.stream):
            strings = self.net._collect_strings(p, self.B, self.gather_host)
        return {"r_strings": [strings["ry"], strings["rz"]], "d_strings": [strings["dy"], strings["dz"]],
                "shape": torch.Size(p.io["shape"])}


class _DecompressHandle:
    def __init__(self, prog, stream, done, word_lens):
        self.prog, self.stream, self.done, self.word_lens = prog, stream, done, word_lens

    def result(self, clone=True):
        self.done.synchronize()
        # a valid stream is consumed exactly: the decoder ends on its last word (reads past the end return zeros and
        # would otherwise decode a truncated or foreign stream silently to garbage)
        pos = self.prog.io["state_host"].numpy()[:, 1]
        # (RGBD_RANS_SKIP: timing experiments of development builds that skip the coder kernels, profiles/tools/skip_probe.py)
        if (pos != self.word_lens).any() and not os.environ.get("RGBD_RANS_SKIP"):
            bad = int(np.nonzero(pos != self.word_lens)[0][0])
            raise ValueError(f"corrupt stream {bad}: decoder stopped at word {int(pos[bad])} of {int(self.word_lens[bad])}")
        r, d = self.prog.io["out_r"], self.prog.io["out_d"]
        if clone:
            with torch.cuda.stream(self.stream):
                r, d = r.clone(), d.clone()
            self.stream.synchronize()
        return {"x_hat": {"r": r, "d": d}}
