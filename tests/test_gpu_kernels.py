"""Kernel-level parity through the C-ABI: entropy elementwise kernels are integer/bit exact vs a
numpy restatement of the reference formulas; the conv family is checked against torch CPU fp32
(tolerance 2e-4 relative to the tensor scale: different but fixed fp32 summation order)."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def sp():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def squeeze_order(x_nchw, parity):
    """[B,C,H,W] -> [B, C*H*W/2] in the reference's symbol order (utils/ckbd.py:51-64, 90-91)"""
    B, Cc, H, W = x_nchw.shape
    out = np.empty((B, Cc, H, W // 2), dtype=x_nchw.dtype)
    a, b = (1, 0) if parity == 0 else (0, 1)
    out[:, :, 0::2] = x_nchw[:, :, 0::2, a::2]
    out[:, :, 1::2] = x_nchw[:, :, 1::2, b::2]
    return out.reshape(B, -1)


def ref_indexes(scales, table, bound=np.float32(0.11)):
    s = np.maximum(scales, bound)
    idx = np.full(s.shape, len(table) - 1, dtype=np.int32)
    for t in table[:-1]:
        idx -= (s <= t).astype(np.int32)
    return idx


@pytest.mark.parametrize("B,H,W,g", [(1, 8, 8, 16), (2, 32, 40, 64), (1, 36, 48, 192), (3, 6, 10, 5)])
def test_ckbd_quantize_index_scatter(golden_dir, B, H, W, g):
    from rgbd_b200 import lib as L
    rng = np.random.default_rng(B * 1000 + g)
    table = np.load(f"{golden_dir}/gauss_tables.npz")["scale_table"]
    kat = np.load(f"{golden_dir}/index_kat.npz")
    Ct, coff = g + 7, 3
    y = (rng.standard_normal((B, H, W, Ct)) * 4).astype(np.float32)
    scales = np.exp(rng.uniform(np.log(0.05), np.log(300), (B, H, W, g))).astype(np.float32)
    flat = scales.reshape(-1)
    n = min(flat.size, kat["scales"].size)
    flat[:n] = kat["scales"][:n]          # every table boundary +-1 ulp, negatives, huge values
    means = (rng.standard_normal((B, H, W, g)) * 2).astype(np.float32)
    y[..., coff:coff + g].reshape(-1)[::7] = (means.reshape(-1)[::7] + 0.5)   # exact .5 ties
    params = np.concatenate([scales, means], axis=-1)
    ny = 2 * g * H * (W // 2) + 11
    for parity in (0, 1):
        d_y, d_p, d_t = (torch.from_numpy(a).to(DEV) for a in (y, params, table))
        sym = torch.full((B, ny), -7, dtype=torch.int32, device=DEV)
        idx = torch.full((B, ny), 255, dtype=torch.uint8, device=DEV)
        yhat = torch.zeros(B, H, W, g + 2, device=DEV)
        off = 5 + parity * g * H * (W // 2)
        L.call("rgbd_ckbd_quantize_index", d_y.data_ptr(), Ct, coff, d_p.data_ptr(), d_t.data_ptr(), len(table),
               0.11, B, H, W, g, parity, sym.data_ptr(), idx.data_ptr(), ny, off, yhat.data_ptr(), L.DT_F32, g + 2, 1,
               sp())
        ys = squeeze_order(np.transpose(y[..., coff:coff + g], (0, 3, 1, 2)), parity)
        ss = squeeze_order(np.transpose(scales, (0, 3, 1, 2)), parity)
        ms = squeeze_order(np.transpose(means, (0, 3, 1, 2)), parity)
        want_sym = np.rint(ys - ms).astype(np.int32)       # np.rint = half-to-even = torch.round
        want_idx = ref_indexes(ss, table)
        n_s = want_sym.shape[1]
        got_sym, got_idx = sym.cpu().numpy(), idx.cpu().numpy()
        assert np.array_equal(got_sym[:, off:off + n_s], want_sym)
        assert np.array_equal(got_idx[:, off:off + n_s].astype(np.int32), want_idx)
        assert (got_sym[:, :off] == -7).all() and (got_sym[:, off + n_s:] == -7).all()   # nothing else touched
        # y_hat = float(sym) + mean at the parity sites, untouched elsewhere
        hh, ww = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
        mask = ((hh + ww) % 2 == 1 - parity)
        want_hat = np.where(mask[None, :, :, None], np.rint(y[..., coff:coff + g] - means) + means, 0).astype(np.float32)
        got_hat = yhat.cpu().numpy()
        assert np.array_equal(got_hat[..., 1:1 + g], want_hat)
        assert (got_hat[..., 0] == 0).all() and (got_hat[..., -1] == 0).all()
        # decoder side: index-only kernel and dequant-scatter agree with the encoder side
        idx2 = torch.full((B, ny), 255, dtype=torch.uint8, device=DEV)
        L.call("rgbd_ckbd_index", d_p.data_ptr(), d_t.data_ptr(), len(table), 0.11, B, H, W, g, parity,
               idx2.data_ptr(), ny, off, sp())
        assert torch.equal(idx2, idx)
        yhat2 = torch.zeros(B, H, W, g + 2, device=DEV)
        L.call("rgbd_ckbd_dequant_scatter", sym.data_ptr(), ny, off, d_p.data_ptr(), B, H, W, g, parity,
               yhat2.data_ptr(), L.DT_F32, g + 2, 1, sp())
        assert torch.equal(yhat2, yhat)


def test_eb_quantize_dequantize():
    from rgbd_b200 import lib as L
    rng = np.random.default_rng(0)
    B, HW, Cc = 2, 80, 192
    z = (rng.standard_normal((B, HW, Cc)) * 9).astype(np.float32)
    med = rng.uniform(-1, 1, Cc).astype(np.float32)
    z[0, :10, :] = med + 1.5
    d_z, d_m = torch.from_numpy(z).to(DEV), torch.from_numpy(med).to(DEV)
    sym = torch.zeros(B, Cc * HW, dtype=torch.int32, device=DEV)
    idx = torch.zeros(B, Cc * HW, dtype=torch.uint8, device=DEV)
    zhat = torch.zeros(B, HW, 3 * Cc, device=DEV)
    L.call("rgbd_eb_quantize", d_z.data_ptr(), Cc, B, HW, Cc, d_m.data_ptr(), sym.data_ptr(), idx.data_ptr(),
           zhat.data_ptr(), L.DT_F32, 3 * Cc, Cc, sp())
    want = np.rint(np.transpose(z, (0, 2, 1)) - med[None, :, None]).astype(np.int32)      # [B, C, HW]
    assert np.array_equal(sym.cpu().numpy().reshape(B, Cc, HW), want)
    assert np.array_equal(idx.cpu().numpy().reshape(B, Cc, HW), np.broadcast_to(np.arange(Cc)[None, :, None], want.shape))
    want_hat = np.transpose(want.astype(np.float32) + med[None, :, None], (0, 2, 1))
    assert np.array_equal(zhat.cpu().numpy()[..., Cc:2 * Cc], want_hat)
    zhat2 = torch.zeros(B, HW, 3 * Cc, device=DEV)
    L.call("rgbd_eb_dequantize", sym.data_ptr(), B, HW, Cc, d_m.data_ptr(), zhat2.data_ptr(), L.DT_F32, 3 * Cc, Cc, sp())
    assert torch.equal(zhat, zhat2)


def run_conv(mod, x_nchw, act=0, epi=0, res=None, mul=None, in_scale=None, dtype=torch.float32, out_dtype=None,
             pad_c=0):
    """Runs one nn.Conv2d / ConvTranspose2d through engine.Builder (-> rgbd_conv_simt)."""
    from rgbd_b200.engine import Builder, PackedConv, View
    b = Builder(torch.device(DEV), dtype, tensor_cores=False)

    def view(t):
        if t is None:
            return None
        nhwc = t.permute(0, 2, 3, 1).contiguous()
        buf = torch.zeros(*nhwc.shape[:3], nhwc.shape[3] + 2 * pad_c, device=DEV, dtype=dtype)
        buf[..., pad_c:pad_c + nhwc.shape[3]] = nhwc.to(DEV).to(dtype)
        return View(buf, pad_c, nhwc.shape[3])

    pc = PackedConv(mod, torch.device(DEV))
    out = b.conv(pc, view(x_nchw), act=act, epi=epi, res=view(res), mul=view(mul),
                 in_scale=None if in_scale is None else in_scale.to(DEV), out_dtype=out_dtype)
    b.prog.run()
    torch.cuda.synchronize()
    return out.torch().float().cpu().permute(0, 3, 1, 2)


CONVS = [
    ("conv5s2_3", lambda: nn.Conv2d(3, 192, 5, 2, 2), (2, 3, 64, 80)),
    ("conv5s2_1", lambda: nn.Conv2d(1, 192, 5, 2, 2), (1, 1, 64, 64)),
    ("conv1x1", lambda: nn.Conv2d(192, 96, 1), (2, 192, 17, 23)),
    ("conv3x3", lambda: nn.Conv2d(96, 96, 3, 1, 1), (2, 96, 19, 21)),
    ("conv3s2p0", lambda: nn.Conv2d(48, 48, 3, 2, 0), (1, 48, 33, 41)),
    ("conv5x5_odd", lambda: nn.Conv2d(85, 64, 5, 1, 2), (1, 85, 16, 20)),
    ("conv1x1_odd", lambda: nn.Conv2d(1280, 213, 1), (1, 1280, 8, 10)),
    ("deconv5s2", lambda: nn.ConvTranspose2d(192, 192, 5, 2, padding=2, output_padding=1), (1, 192, 9, 11)),
    ("deconv5s2_rgb", lambda: nn.ConvTranspose2d(192, 3, 5, 2, padding=2, output_padding=1), (2, 192, 16, 16)),
    ("deconv3s1", lambda: nn.ConvTranspose2d(96, 64, 3, 1, padding=1), (1, 96, 8, 10)),
]


def rel_err(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


@pytest.mark.parametrize("name,make,shape", CONVS, ids=[c[0] for c in CONVS])
def test_conv_simt_fp32_vs_torch(name, make, shape):
    torch.manual_seed(1)
    mod = make().eval()
    x = torch.randn(shape)
    with torch.no_grad():
        want = mod(x)
    got = run_conv(mod, x, pad_c=4 if shape[1] % 4 == 0 else 0)
    assert got.shape == want.shape
    assert rel_err(got, want) < 2e-4, name      # fp32 tolerance: summation order differs from mkldnn


def test_conv_epilogues_vs_torch():
    torch.manual_seed(2)
    mod = nn.Conv2d(96, 192, 1).eval()
    x, res, mul = torch.randn(2, 96, 10, 12), torch.randn(2, 192, 10, 12), torch.randn(2, 192, 10, 12)
    with torch.no_grad():
        v = mod(x)
        assert rel_err(run_conv(mod, x, act=1), F.relu(v)) < 2e-4
        assert rel_err(run_conv(mod, x, act=2), F.leaky_relu(v)) < 2e-4
        assert rel_err(run_conv(mod, x, res=res), v + res) < 2e-4
        assert rel_err(run_conv(mod, x, act=1, res=res), F.relu(v + res)) < 2e-4
        assert rel_err(run_conv(mod, x, epi=1, mul=mul, res=res), res + mul * torch.sigmoid(v)) < 2e-4
        assert rel_err(run_conv(mod, x, epi=1, mul=mul), mul * torch.sigmoid(v)) < 2e-4
        scale = torch.rand(2, 96) + 0.5
        assert rel_err(run_conv(mod, x, in_scale=scale), mod(x * scale[:, :, None, None])) < 2e-4
        # bilinear-upsample epilogue == F.interpolate(align_corners=False) + add (ESA, attention.py:92-94)
        mod2 = nn.Conv2d(48, 48, 1).eval()
        x2 = torch.randn(2, 48, 37, 45)
        for small in ((2, 48, 5, 7), (2, 48, 1, 1), (2, 48, 11, 13)):
            sm = torch.randn(small)
            want = mod2(x2) + F.interpolate(sm, (37, 45), mode="bilinear", align_corners=False)
            assert rel_err(run_conv(mod2, x2, epi=2, res=sm), want) < 2e-4, small


def test_conv_bf16_io_vs_torch():
    """bf16 activations, fp32 accumulate: error bounded by bf16 rounding of inputs/outputs."""
    torch.manual_seed(3)
    mod = nn.Conv2d(64, 96, 3, 1, 1).eval()
    x = torch.randn(1, 64, 12, 12)
    with torch.no_grad():
        want = mod(x.bfloat16().float())
    got = run_conv(mod, x, dtype=torch.bfloat16)
    assert rel_err(got, want) < 1e-2
    got32 = run_conv(mod, x, dtype=torch.bfloat16, out_dtype=torch.float32)
    assert rel_err(got32, want) < 2e-4


def test_se_maxpool_layout():
    from rgbd_b200 import lib as L
    from rgbd_b200.engine import Builder, View
    torch.manual_seed(4)
    b = Builder(torch.device(DEV), torch.float32)
    x = torch.randn(2, 30, 34, 100)
    w1, w2 = torch.randn(6, 96) * 0.2, torch.randn(96, 6) * 0.2
    xv = View(x.to(DEV), 2, 96)
    s0 = b.se_scale(xv, w1.to(DEV), w2.to(DEV), plus_one=False)
    s1 = b.se_scale(xv, w1.to(DEV), w2.to(DEV), plus_one=True)
    mp = b.maxpool7s3(View(x.to(DEV)))
    b.prog.run()
    torch.cuda.synchronize()
    want = torch.sigmoid(F.linear(F.relu(F.linear(x[..., 2:98].mean(dim=(1, 2)), w1)), w2))
    assert torch.allclose(s0.cpu(), want, atol=1e-5)
    assert torch.allclose(s1.cpu(), want + 1, atol=1e-5)
    want_mp = F.max_pool2d(x.permute(0, 3, 1, 2), 7, 3).permute(0, 2, 3, 1)
    assert torch.equal(mp.torch().cpu(), want_mp)
    # NCHW <-> NHWC boundary + clamp
    img = torch.rand(2, 3, 16, 20) * 1.4 - 0.2
    nhwc = torch.zeros(2, 16, 20, 3, device=DEV)
    back = torch.zeros(2, 3, 16, 20, device=DEV)
    L.call("rgbd_nchw_to_nhwc", img.to(DEV).data_ptr(), nhwc.data_ptr(), L.DT_F32, 2, 3, 16, 20, 3, 0, 0, sp())
    L.call("rgbd_nhwc_to_nchw", nhwc.data_ptr(), L.DT_F32, back.data_ptr(), 2, 3, 16, 20, 3, 0, 1, sp())
    assert torch.equal(nhwc.cpu(), img.permute(0, 2, 3, 1))
    assert torch.equal(back.cpu(), img.clamp(0, 1))


def test_se_cached_table_equals_full_recomputation():
    """se_scale on a buffer registered with Builder.se_cache: growing prefixes, channel ranges rewritten by convs in
    between — every gate is bit-identical to the uncached kernel run on the same buffer contents."""
    import torch.nn as nn
    from rgbd_b200.engine import Builder, PackedConv, View
    torch.manual_seed(14)
    dev = torch.device(DEV)
    outs = {}
    for cached in (True, False):
        torch.manual_seed(15)
        b = Builder(dev, torch.float32, tensor_cores=False)
        buf = b.alloc(2, 12, 20, 96)
        buf.buf.copy_(torch.randn(2, 12, 20, 96))
        src = b.alloc(2, 12, 20, 8)
        src.buf.copy_(torch.randn(2, 12, 20, 8))
        if cached:
            b.se_cache(buf)
        convs = [PackedConv(nn.Conv2d(8, 16, 3, 1, 1), dev) for _ in range(3)]
        gates = []
        for width, (pc, off) in zip((40, 64, 96, 96), [(None, 0), (convs[0], 24), (convs[1], 70), (convs[2], 0)]):
            if pc is not None:
                b.conv(pc, src, out=buf.sub(off, 16))          # rewrites 16 channels: their sums go stale
            w1, w2 = (torch.randn(5, width) * 0.2).to(dev), (torch.randn(width, 5) * 0.2).to(dev)
            gates.append(b.se_scale(buf.sub(0, width), w1, w2, plus_one=True))
        b.prog.run()
        torch.cuda.synchronize()
        outs[cached] = [g.cpu() for g in gates]
        if cached:
            n_partial = sum(1 for op in b.prog.ops if getattr(op, "label", "") == "rgbd_se_partial")
            assert n_partial == 4      # [0, 40) | [24, 64) | [64, 96) | [0, 16): one refresh per call, never the whole table
    for a, c in zip(outs[False], outs[True]):
        assert torch.equal(a, c)


def test_likelihood_kernels_vs_torch():
    """forward()-only kernels: erfc / sigmoid chains, fp32 tolerance 1e-5 absolute."""
    import rgbd_b200
    from rgbd_b200 import lib as L
    from oracle.model_oracle import OracleCodec
    net = rgbd_b200.ELIC_united(config=rgbd_b200.model_config(), channel=4)
    net.load_state_dict(rgbd_b200.synthetic.synthetic_state_dict(net, 0, "mid"))
    net.update(force=True)
    orc = OracleCodec(net.state_dict())
    rng = np.random.default_rng(8)
    B, H, W, g = 2, 8, 10, 32
    y = torch.from_numpy((rng.standard_normal((B, g, H, W)) * 3).astype(np.float32))
    scales = torch.from_numpy(np.exp(rng.uniform(np.log(0.05), np.log(20), (B, g, H, W))).astype(np.float32))
    means = torch.from_numpy(rng.standard_normal((B, g, H, W)).astype(np.float32))
    want = orc._gauss_likelihood("rgb", y, scales, means)
    params = torch.cat([scales, means], 1).permute(0, 2, 3, 1).contiguous().to(DEV)
    y_d = y.permute(0, 2, 3, 1).contiguous().to(DEV)
    lik = torch.zeros(B, g + 3, H, W, device=DEV)
    yhat = torch.zeros(B, H, W, g, device=DEV)
    for parity in (0, 1):
        L.call("rgbd_ckbd_ste_likelihood", y_d.data_ptr(), g, 0, params.data_ptr(), 0.11, 1e-9, B, H, W, g, parity,
               yhat.data_ptr(), L.DT_F32, g, 0, lik.data_ptr(), g + 3, 2, sp())
    torch.cuda.synchronize()
    assert torch.allclose(lik[:, 2:2 + g].cpu(), want, atol=1e-5, rtol=1e-4)
    d = y - means
    assert torch.allclose(yhat.cpu().permute(0, 3, 1, 2), (torch.round(d) - d + d) + means, atol=1e-6)
    # factorised prior
    z = torch.from_numpy((rng.standard_normal((B, 192, 4, 5)) * 6).astype(np.float32))
    zh_want, lz_want = orc.eb_forward("rgb", z)
    ebp = net.rgb_entropy_bottleneck.packed_params(DEV)
    z_d = z.permute(0, 2, 3, 1).contiguous().to(DEV)
    zhat = torch.zeros(B, 20, 192, device=DEV)
    lz = torch.zeros(B, 192, 4, 5, device=DEV)
    L.call("rgbd_eb_likelihood", z_d.data_ptr(), 192, B, 20, 192, ebp.data_ptr(), 1e-9, zhat.data_ptr(), L.DT_F32, 192,
           0, lz.data_ptr(), sp())
    torch.cuda.synchronize()
    assert torch.allclose(lz.cpu(), lz_want, atol=1e-5, rtol=1e-4)
    assert torch.allclose(zhat.cpu().reshape(B, 4, 5, 192).permute(0, 3, 1, 2), zh_want, atol=1e-5)


# ------------------------------------------------------------------------------------------------
# tcgen05 tensor-core conv path (bf16 operands, fp32 TMEM accumulators) vs torch CPU on the same
# bf16-rounded inputs and weights.  fp32 outputs: tolerance 1e-3 of the tensor scale (accumulation
# order / tensor-core fp32 accumulate); bf16 outputs: + one bf16 rounding (2^-8).
# ------------------------------------------------------------------------------------------------
def run_conv_tc(mod, x_nchw, tensor_cores=True, out_dtype=torch.float32, **kw):
    from rgbd_b200.engine import Builder, PackedConv, View
    b = Builder(torch.device(DEV), torch.bfloat16, tensor_cores=tensor_cores)

    def view(t, pad=8):
        if t is None:
            return None
        nhwc = t.permute(0, 2, 3, 1).contiguous()
        Cp = (nhwc.shape[3] + 7) // 8 * 8 + 2 * pad
        buf = torch.zeros(*nhwc.shape[:3], Cp, device=DEV, dtype=torch.bfloat16)
        buf[..., pad:pad + nhwc.shape[3]] = nhwc.to(DEV).to(torch.bfloat16)
        return View(buf, pad, nhwc.shape[3])

    pc = PackedConv(mod, torch.device(DEV))
    args = {k: (view(v) if k in ("res", "mul") else v) for k, v in kw.items()}
    if "in_scale" in args:
        args["in_scale"] = args["in_scale"].to(DEV)
    out = b.conv(pc, view(x_nchw), out_dtype=out_dtype, **args)
    assert (b.prog.n_tc > 0) == tensor_cores, (b.prog.n_tc, b.prog.n_simt)
    b.prog.run()
    torch.cuda.synchronize()
    return out.torch().float().cpu().permute(0, 3, 1, 2)


def bf16_ref(mod, x):
    m = type(mod)(mod.in_channels, mod.out_channels, mod.kernel_size, mod.stride, mod.padding,
                  **({"output_padding": mod.output_padding} if isinstance(mod, nn.ConvTranspose2d) else {})).eval()
    with torch.no_grad():
        m.weight.copy_(mod.weight.bfloat16().float())
        m.bias.copy_(mod.bias)
        return m(x.bfloat16().float())


TC_CONVS = [
    ("5x5s2_rgb_3_192", lambda: nn.Conv2d(3, 192, 5, 2, 2), (2, 3, 64, 80)),
    ("5x5s2_depth_1_192", lambda: nn.Conv2d(1, 192, 5, 2, 2), (1, 1, 64, 64)),
    ("1x1_192_96", lambda: nn.Conv2d(192, 96, 1), (2, 192, 32, 40)),
    ("3x3_96_96_ragged", lambda: nn.Conv2d(96, 96, 3, 1, 1), (2, 96, 19, 21)),
    ("5x5s2_384_192", lambda: nn.Conv2d(384, 192, 5, 2, 2), (1, 384, 32, 48)),
    ("5x5s2_odd_hw", lambda: nn.Conv2d(64, 64, 5, 2, 2), (1, 64, 17, 23)),
    ("3x3s2p0_48", lambda: nn.Conv2d(48, 48, 3, 2, 0), (1, 48, 33, 41)),
    ("5x5_ctx_16_32", lambda: nn.Conv2d(16, 32, 5, 1, 2), (2, 16, 32, 40)),
    ("1x1_1344_224", lambda: nn.Conv2d(1344, 224, 1), (1, 1344, 8, 12)),
    ("3x3_213_42_oddC", lambda: nn.Conv2d(213, 42, 3, 1, 1), (1, 213, 16, 20)),
    ("5x5_512_384", lambda: nn.Conv2d(512, 384, 5, 1, 2), (1, 512, 8, 10)),
    ("deconv5s2_192", lambda: nn.ConvTranspose2d(192, 192, 5, 2, padding=2, output_padding=1), (1, 192, 9, 11)),
    ("deconv5s2_to3", lambda: nn.ConvTranspose2d(192, 3, 5, 2, padding=2, output_padding=1), (2, 192, 16, 16)),
    ("deconv5s2_to1_ragged", lambda: nn.ConvTranspose2d(192, 1, 5, 2, padding=2, output_padding=1), (2, 192, 19, 45)),
    ("deconv3s1_960_640", lambda: nn.ConvTranspose2d(960, 640, 3, 1, padding=1), (1, 960, 8, 10)),
    # two-accumulator super-tiles, several tiles per row, halo rows shared between the M tiles
    ("3x3_96_96_wide", lambda: nn.Conv2d(96, 96, 3, 1, 1), (2, 96, 64, 80)),
    ("1x1_192_96_wide", lambda: nn.Conv2d(192, 96, 1), (1, 192, 64, 80)),
    ("5x5_224_128_latent", lambda: nn.Conv2d(224, 128, 5, 1, 2), (2, 224, 32, 40)),
    ("5x5s2_192_192_wide", lambda: nn.Conv2d(192, 192, 5, 2, 2), (1, 192, 64, 96)),
    ("deconv5s2_320_192_latent", lambda: nn.ConvTranspose2d(320, 192, 5, 2, padding=2, output_padding=1), (1, 320, 32, 40)),
]


@pytest.mark.parametrize("name,make,shape", TC_CONVS, ids=[c[0] for c in TC_CONVS])
def test_conv_tc_vs_torch(name, make, shape):
    torch.manual_seed(11)
    mod = make().eval()
    x = torch.randn(shape)
    want = bf16_ref(mod, x)
    got = run_conv_tc(mod, x)
    assert got.shape == want.shape
    assert rel_err(got, want) < 1e-3, (name, rel_err(got, want))
    # the CUDA-core kernel on the same bf16 activations (it keeps fp32 weights)
    simt = run_conv_tc(mod, x, tensor_cores=False)
    with torch.no_grad():
        assert rel_err(simt, mod(x.bfloat16().float())) < 2e-4
    got16 = run_conv_tc(mod, x, out_dtype=torch.bfloat16)
    assert rel_err(got16, want) < 6e-3, name


def test_conv_tc_epilogues():
    torch.manual_seed(12)
    mod = nn.Conv2d(96, 192, 1).eval()
    x, res, mul = torch.randn(2, 96, 10, 12), torch.randn(2, 192, 10, 12), torch.randn(2, 192, 10, 12)
    rb, mb = res.bfloat16().float(), mul.bfloat16().float()
    v = bf16_ref(mod, x)
    tol = 1e-3
    assert rel_err(run_conv_tc(mod, x, act=1), F.relu(v)) < tol
    assert rel_err(run_conv_tc(mod, x, act=2), F.leaky_relu(v)) < tol
    assert rel_err(run_conv_tc(mod, x, act=1, res=res), F.relu(v + rb)) < tol
    assert rel_err(run_conv_tc(mod, x, epi=1, mul=mul, res=res), rb + mb * torch.sigmoid(v)) < tol
    assert rel_err(run_conv_tc(mod, x, epi=1, mul=mul), mb * torch.sigmoid(v)) < tol
    # residual riding on the tensor core (identity B tile): several N tiles, ragged last residual block
    for cin, cout, hw in ((96, 192, (40, 48)), (160, 320, (32, 40)), (48, 40, (9, 11))):
        m3 = nn.Conv2d(cin, cout, 1).eval()
        x3, r3 = torch.randn(2, cin, *hw), torch.randn(2, cout, *hw)
        assert rel_err(run_conv_tc(m3, x3, act=1, res=r3), F.relu(bf16_ref(m3, x3) + r3.bfloat16().float())) < tol
        assert rel_err(run_conv_tc(m3, x3, res=r3, out_dtype=torch.bfloat16), bf16_ref(m3, x3) + r3.bfloat16().float()) < 6e-3
    m4 = nn.Conv2d(96, 96, 3, 1, 1).eval()
    x4, r4 = torch.randn(1, 96, 30, 44), torch.randn(1, 96, 30, 44)
    assert rel_err(run_conv_tc(m4, x4, res=r4), bf16_ref(m4, x4) + r4.bfloat16().float()) < tol
    # SE gate folded through the separate scale pass (bf16 rounding of the scaled input)
    scale = torch.rand(2, 96) + 0.5
    want = bf16_ref(mod, (x.bfloat16().float() * scale[:, :, None, None]))
    assert rel_err(run_conv_tc(mod, x, in_scale=scale), want) < 1e-2
    # larger map: the gate is folded into per-image filter copies instead (rgbd_scale_weights, one bf16
    # rounding of w * scale from fp32)
    xl = torch.randn(3, 96, 32, 40)
    sl = torch.rand(3, 96) + 0.5
    m_ref = nn.Conv2d(96, 192, 1).eval()
    outs = []
    for n in range(3):
        with torch.no_grad():
            m_ref.weight.copy_((mod.weight * sl[n][None, :, None, None]).bfloat16().float())
            m_ref.bias.copy_(mod.bias)
            outs.append(m_ref(xl[n:n + 1].bfloat16().float()))
    assert rel_err(run_conv_tc(mod, xl, in_scale=sl), torch.cat(outs)) < 1e-3
    mod2 = nn.Conv2d(48, 48, 1).eval()
    x2 = torch.randn(2, 48, 37, 45)
    for small in ((2, 48, 5, 7), (2, 48, 11, 13)):
        sm = torch.randn(small)
        want = bf16_ref(mod2, x2) + F.interpolate(sm.bfloat16().float(), (37, 45), mode="bilinear", align_corners=False)
        assert rel_err(run_conv_tc(mod2, x2, epi=2, res=sm), want) < tol, small


def test_conv_tc_is_deterministic_and_batch_invariant():
    """Same pixels -> same bits, whatever the batch size / tile position."""
    torch.manual_seed(13)
    mod = nn.Conv2d(128, 64, 3, 1, 1).eval()
    x = torch.randn(3, 128, 24, 40)
    full = run_conv_tc(mod, x, out_dtype=torch.float32)
    again = run_conv_tc(mod, x, out_dtype=torch.float32)
    assert torch.equal(full, again)
    for i in range(3):
        one = run_conv_tc(mod, x[i:i + 1], out_dtype=torch.float32)
        assert torch.equal(one[0], full[i])


def test_split_bf16_first_layer_is_fp32_accurate():
    """Image layers on the tensor cores: [hi | lo | hi] input x [w_hi | w_hi | w_lo] weights keeps the
    16-bit depth exact and the product accurate to ~2^-16 (vs 2^-9 for plain bf16)."""
    from rgbd_b200 import lib as L
    from rgbd_b200.engine import Builder, PackedConv, View
    torch.manual_seed(21)
    for cin in (3, 1):
        mod = nn.Conv2d(cin, 192, 5, 2, 2).eval()
        img = torch.round(torch.rand(2, cin, 64, 96) * 65535) / 65535
        b = Builder(torch.device(DEV), torch.bfloat16, tensor_cores=True)
        x = b.alloc(2, 64, 96, 3 * cin)
        src = img.to(DEV)
        b.op("rgbd_nchw_to_nhwc", src.data_ptr(), x.ptr(), L.DT_BF16, 2, cin, 64, 96, x.cstride, x.coff, 1)
        out = b.conv(PackedConv(mod, torch.device(DEV), split3=True), x, out_dtype=torch.float32)
        assert b.prog.n_tc == 1
        b.prog.run()
        torch.cuda.synchronize()
        with torch.no_grad():
            want = mod(img)
        got = out.torch().float().cpu().permute(0, 3, 1, 2)
        assert rel_err(got, want) < 5e-5, (cin, rel_err(got, want))
        # the same layer on the space-to-depth map (what the codec uses): 3x3 stride 1, one tap group
        b2 = Builder(torch.device(DEV), torch.bfloat16, tensor_cores=True)
        x2 = b2.alloc(2, 32, 48, 12 * cin)
        b2.op("rgbd_nchw_to_nhwc", src.data_ptr(), x2.ptr(), L.DT_BF16, 2, cin, 64, 96, x2.cstride, x2.coff, 2)
        out2 = b2.conv(PackedConv(mod, torch.device(DEV), split3=True, s2d=True), x2, out_dtype=torch.float32)
        assert b2.prog.n_tc == 1 and (out2.H, out2.W) == (32, 48)
        b2.prog.run()
        torch.cuda.synchronize()
        got2 = out2.torch().float().cpu().permute(0, 3, 1, 2)
        assert rel_err(got2, want) < 5e-5, (cin, rel_err(got2, want))
        assert abs(b2.prog.flops - b.prog.flops) < 1
