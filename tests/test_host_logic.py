"""Host-side logic that needs no GPU: launch decomposition of the conv family (checked by a
torch-CPU emulation of the kernel's documented semantics), context-buffer permutations,
CDF tables, the exact-reciprocal identity the rANS encoder relies on, and error behaviour."""
import json
import random

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

import rgbd_b200
from rgbd_b200.engine import PackedConv


def emulate(pc, x_nhwc):
    """The kernel contract of include/rgbd_b200.h in slow torch: per launch, per tap."""
    N, H, W, Cin = x_nhwc.shape
    Ho, Wo, launches = pc.launches(H, W)
    y = torch.zeros(N, Ho, Wo, pc.Cout)
    for ln in launches:
        oy = torch.arange(ln["Hs"])
        ox = torch.arange(ln["Ws"])
        acc = torch.zeros(N, ln["Hs"], ln["Ws"], pc.cout_pad)
        for dy, dx, wt in ln["taps"]:
            iy, ix = oy * ln["i_step"] + dy, ox * ln["i_step"] + dx
            vy, vx = (iy >= 0) & (iy < H), (ix >= 0) & (ix < W)
            g = x_nhwc[:, iy.clamp(0, H - 1)][:, :, ix.clamp(0, W - 1)]
            g = g * (vy[:, None] & vx[None, :])[None, :, :, None]
            acc += g @ pc.w32[wt]
        acc = acc[..., :pc.Cout] + pc.bias
        y[:, ln["o_off_y"]::ln["o_step"], ln["o_off_x"]::ln["o_step"]] = acc
    return y


@pytest.mark.parametrize("mod", [
    nn.Conv2d(5, 7, 5, 2, 2), nn.Conv2d(4, 6, 3, 1, 1), nn.Conv2d(6, 3, 1), nn.Conv2d(4, 4, 3, 2, 0),
    nn.Conv2d(3, 8, 5, 1, 2),
    nn.ConvTranspose2d(5, 6, 5, 2, padding=2, output_padding=1),
    nn.ConvTranspose2d(4, 3, 3, 1, padding=1, output_padding=0),
])
def test_launch_decomposition_equals_torch(mod):
    torch.manual_seed(0)
    x = torch.randn(2, mod.in_channels, 9, 12)
    with torch.no_grad():
        want = mod(x)
        got = emulate(PackedConv(mod, "cpu"), x.permute(0, 2, 3, 1).contiguous()).permute(0, 3, 1, 2)
    assert got.shape == want.shape
    assert torch.allclose(got, want, atol=1e-5, rtol=1e-5)


def test_fused_parity_deconv_weights_equal_torch():
    """ConvTranspose2d(5, 2, 2, output_padding=1) with <= 4 output channels as ONE 3x3 conv over the input lattice
    whose 16 columns are (output parity, channel) + a pixel shuffle (RGBD_EPI_SHUFFLE2): emulated in torch."""
    torch.manual_seed(1)
    for cout in (3, 1):
        mod = nn.ConvTranspose2d(6, cout, 5, 2, padding=2, output_padding=1)
        pc = PackedConv(mod, "cpu")
        w = pc.wtc_shuffle.float()                       # [9 taps][16 cols][cin_pad]
        x = torch.randn(2, 6, 7, 9).bfloat16().float()
        xh = x.permute(0, 2, 3, 1)
        N, H, W, Cin = xh.shape
        acc = torch.zeros(N, H, W, 16)
        for u, (dy, dx) in enumerate((a, b) for a in (-1, 0, 1) for b in (-1, 0, 1)):
            iy, ix = torch.arange(H) + dy, torch.arange(W) + dx
            vy, vx = (iy >= 0) & (iy < H), (ix >= 0) & (ix < W)
            g = xh[:, iy.clamp(0, H - 1)][:, :, ix.clamp(0, W - 1)] * (vy[:, None] & vx[None, :])[None, :, :, None]
            acc += g @ w[u, :, :Cin].t()
        y = torch.zeros(N, 2 * H, 2 * W, cout)
        for q in range(4):
            y[:, (q >> 1)::2, (q & 1)::2] = acc[..., 4 * q:4 * q + cout] + mod.bias
        with torch.no_grad():
            m2 = nn.ConvTranspose2d(6, cout, 5, 2, padding=2, output_padding=1)
            m2.weight.copy_(mod.weight.bfloat16().float())
            m2.bias.copy_(mod.bias)
            want = m2(x)
        assert torch.allclose(y.permute(0, 3, 1, 2), want, atol=1e-4, rtol=1e-4)


def test_space_to_depth_first_layer_equals_torch():
    """Conv2d(k5, s2, p2) over an image == the packed 3x3 s1 conv over its space-to-depth map (split3 = 2 layout:
    channel block 2*ry+rx of [hi | lo | hi]); here with lo = 0, hi = x to check the tap / channel bookkeeping."""
    torch.manual_seed(2)
    for cin in (3, 1):
        mod = nn.Conv2d(cin, 8, 5, 2, 2)
        pc = PackedConv(mod, "cpu", split3=True, s2d=True)
        assert (pc.k, pc.stride, pc.pad, pc.Cin, pc.flop_taps) == (3, 1, 1, 12 * cin, 25)
        x = torch.randn(2, cin, 8, 12)
        hi = x.bfloat16().float()
        lo = x - hi
        # [N, H/2, W/2, 4 blocks x (hi | lo | hi)]
        blocks = []
        for ry in range(2):
            for rx in range(2):
                blocks += [hi[:, :, ry::2, rx::2], lo[:, :, ry::2, rx::2], hi[:, :, ry::2, rx::2]]
        s2d = torch.cat(blocks, dim=1).permute(0, 2, 3, 1).contiguous()
        got = emulate(pc, s2d).permute(0, 3, 1, 2)
        with torch.no_grad():
            want = mod(x)
        # hi*w_hi + lo*w_hi + hi*w_lo = x*w - lo*w_lo: the dropped term is ~2^-16 relative
        assert got.shape == want.shape
        assert torch.allclose(got, want, atol=2e-4, rtol=1e-4)


def test_deconv_phase_flops_equal_dense_count():
    pc = PackedConv(nn.ConvTranspose2d(4, 4, 5, 2, padding=2, output_padding=1), "cpu")
    _, _, launches = pc.launches(8, 8)
    assert sorted(len(l["taps"]) for l in launches) == [4, 6, 6, 9]   # 25 taps in total


def test_context_buffer_permutation():
    """Our concat order [hyper, ch, loc_r, loc_d] with permuted weights == the reference's order."""
    net = rgbd_b200.ELIC_united(config=rgbd_b200.model_config(), channel=4)
    M = net.M
    for idx in (0, 2):
        g = net.slice_ch[idx]
        o, perms, _ = net._ctx_layout(idx)
        parts = {"hyper_r": torch.randn(1, 2 * M, 2, 2), "hyper_d": torch.randn(1, 2 * M, 2, 2),
                 "loc_r": torch.randn(1, 2 * g, 2, 2), "loc_d": torch.randn(1, 2 * g, 2, 2)}
        base = ["hyper_r", "hyper_d"]
        if idx:
            parts["ch_r"], parts["ch_d"] = torch.randn(1, 2 * g, 2, 2), torch.randn(1, 2 * g, 2, 2)
            base += ["ch_r", "ch_d"]
        ours_all = torch.cat([parts[k] for k in base + ["loc_r", "loc_d"]], 1)
        ref_orders = {"r_anchor": base, "d_anchor": ["loc_r"] + base, "r_nonanchor": ["loc_r", "loc_d"] + base,
                      "d_nonanchor": ["loc_r", "loc_d"] + base}
        mods = {"r_anchor": net.rgb_entropy_parameters_anchor, "d_anchor": net.depth_entropy_parameters_anchor,
                "r_nonanchor": net.rgb_entropy_parameters_nonanchor, "d_nonanchor": net.depth_entropy_parameters_nonanchor}
        for name, order in ref_orders.items():
            ref_x = torch.cat([parts[k] for k in order], 1)
            perm = perms[name]
            width = ref_x.shape[1]
            assert mods[name][idx].fusion[0].in_channels == width
            ours = ours_all[:, :width]
            assert torch.equal(ours, ref_x[:, perm])
            conv = mods[name][idx].fusion[0]
            with torch.no_grad():
                want = conv(ref_x)
                got = F.conv2d(ours, conv.weight[:, perm], conv.bias)
            assert torch.allclose(got, want, atol=1e-4)


def test_gaussian_tables_equal_reference(golden_dir, built_lib):
    g = np.load(f"{golden_dir}/gauss_tables.npz")
    net = rgbd_b200.ELIC_united(config=rgbd_b200.model_config(), channel=4)
    assert net.update(force=True)
    gc = net.depth_gaussian_conditional
    assert np.array_equal(gc.quantized_cdf.numpy(), g["cdf"])
    assert np.array_equal(gc.cdf_length.numpy(), g["lengths"])
    assert np.array_equal(gc.offset.numpy(), g["offsets"])
    assert np.array_equal(gc.scale_table.numpy(), g["scale_table"])
    assert int(g["lengths"].sum()) == 27256


def test_load_state_dict_resizes_tables():
    a = rgbd_b200.ELIC_united(config=rgbd_b200.model_config(), channel=4)
    a.update(force=True)
    b = rgbd_b200.ELIC_united(config=rgbd_b200.model_config(), channel=4)
    b.load_state_dict(a.state_dict())          # checkpoint saved after update() into a fresh model
    assert torch.equal(b.rgb_gaussian_conditional.quantized_cdf, a.rgb_gaussian_conditional.quantized_cdf)
    b.load_state_dict(a.state_dict())          # and into an already-updated one
    bad = dict(a.state_dict())
    bad.pop("g_a.rgb_analysis_transform.0.weight")
    with pytest.raises(RuntimeError):
        b.load_state_dict(bad, strict=True)
    # the reference's default (models/elic_united.py:588-620): strict is tried first, then a non-strict load
    rv = b.load_state_dict(bad)
    assert "g_a.rgb_analysis_transform.0.weight" in rv.missing_keys


def test_update_does_not_depend_on_the_default_device():
    """playground/test.py:20 makes CUDA the default tensor type: every factory call inside update() has to name its
    device.  Emulated here with the `meta` default device (a device-less factory call would make a meta tensor
    and the mix with the CPU tensors raises)."""
    a = rgbd_b200.ELIC_united(config=rgbd_b200.model_config(), channel=4)
    a.update(force=True)
    want = {k: v.clone() for k, v in a.state_dict().items() if "_quantized_cdf" in k or "_offset" in k or "_cdf_length" in k}
    b = rgbd_b200.ELIC_united(config=rgbd_b200.model_config(), channel=4)
    b.load_state_dict({k: v for k, v in a.state_dict().items()})
    torch.set_default_device("meta")
    try:
        b.update(force=True)
        perm = b._ctx_layout(1)[1]["d_nonanchor"]
    finally:
        torch.set_default_device("cpu")
    assert perm.device.type == "cpu"
    for k, v in want.items():
        assert torch.equal(b.state_dict()[k], v), k


def test_forward_rejects_non_ste_quant():
    cfg = rgbd_b200.model_config()
    cfg["quant"] = "noise"
    net = rgbd_b200.ELIC_united(config=cfg, channel=4)
    with pytest.raises(NotImplementedError):
        net(torch.zeros(1, 3, 128, 128), torch.zeros(1, 1, 128, 128))


def test_exact_reciprocal_division():
    """q = mulhi(x, rcp) >> shift == x // f for every x < 2^63 the encoder can see
    (ryg_rans rans64.h:167-247; csrc/rans.cu make_reciprocal)."""
    rnd = random.Random(5)
    for f in list(range(2, 70)) + [rnd.randrange(2, 65536) for _ in range(400)] + [65535, 32768, 32769]:
        s = (f - 1).bit_length()
        rcp = ((1 << (s + 63)) + f - 1) // f
        assert rcp < (1 << 64)
        for x in [1 << 31, (1 << 31) + 1, (f << 47) - 1, f << 31] + [rnd.randrange(1 << 31, f << 47) for _ in range(60)]:
            assert ((x * rcp) >> 64) >> (s - 1) == x // f, (f, x)


def test_cpu_device_is_refused_without_fallback(built_lib):
    net = rgbd_b200.ELIC_united(config=rgbd_b200.model_config(), channel=4).eval()
    net.update(force=True)
    with pytest.raises(rgbd_b200.lib.RgbdError, match="no CPU fallback"):
        net.compress(torch.zeros(1, 3, 128, 128), torch.zeros(1, 1, 128, 128))
    with pytest.raises(ValueError, match="multiples of 64"):
        net.compress(torch.zeros(1, 3, 100, 128), torch.zeros(1, 1, 100, 128))


def test_model_zoo_lookup_order():
    # substring lookup in dict order must find R2D before the bidirectional model
    name = "ELIC_united_R2D_q2"
    hit = next(v for k, v in rgbd_b200.modelZoo.items() if name.find(k) != -1)
    assert hit is rgbd_b200.ELIC_united_R2D


def test_plan_arena_reuses_and_coalesces_storage():
    """Builder.alloc / release: a small heap over the plan's byte arena (engine.py) — released ranges are reused by
    later tensors of any shape, neighbours coalesce, two views of one buffer release it once."""
    from rgbd_b200.engine import Builder, View
    b = Builder(torch.device("cpu"), torch.bfloat16, tensor_cores=False)
    a = b.alloc(1, 8, 8, 16)                      # 2 KB inside a fresh 64 MB chunk
    c = b.alloc(1, 8, 8, 32)                      # 4 KB right behind it
    assert b.prog.bytes == 64 << 20 and len(b.prog.pool["chunks"]) == 1
    pa, pc_ = a.ptr(), c.ptr()
    assert pc_ == pa + 2048
    b.release(a, View(a.buf, 0, 8))               # second view of the same buffer: released once
    assert b.prog.pool["chunks"][0][1][0] == (0, 2048)
    d = b.alloc(1, 4, 4, 16, torch.float32)       # 1 KB fp32 tensor reuses the head of the freed range
    assert d.ptr() == pa and d.dtype == torch.float32 and d.cstride == 16
    b.release(c, d)
    # everything free again and coalesced into one range
    assert b.prog.pool["chunks"][0][1] == [(0, 64 << 20)]
    e = b.alloc(1, 16, 16, 24)                    # bf16 rows padded to 8 channels: 24 -> 24, 12 KB
    assert e.ptr() == pa and e.cstride == 24 and b.prog.bytes == 64 << 20
    z = b.alloc(1, 8, 8, 16, zero=True)           # zero-initialised buffers get a private, exact chunk
    assert len(b.prog.pool["chunks"]) == 2 and float(z.buf.float().abs().sum()) == 0.0


def test_oracle_tables_equal_product_update(built_lib):
    """oracle/tables.py (the CPU arm's update(force=True), built on oracle/_ref/_CXX or the C restatement) produces the
    tables the product's update() produces."""
    from oracle.tables import updated_state_dict
    net = rgbd_b200.ELIC_united(config=rgbd_b200.model_config(), channel=4)
    sd0 = rgbd_b200.synthetic.synthetic_state_dict(net, 0, "mid")
    net.load_state_dict(sd0)
    net.update(force=True)
    want = net.state_dict()
    got = updated_state_dict(sd0)
    assert list(got) == list(want)
    for k in want:
        assert got[k].shape == want[k].shape and torch.equal(got[k], want[k]), k


def test_cpu_arm_does_not_load_the_product_library():
    """bench.py --impl reference / cpu_baseline must time the oracle only (the arm used to map librgbd_b200.so)."""
    import subprocess
    import sys
    code = ("import sys, types; sys.argv=['bench.py','--height','128','--width','128','--preset','mid'];"
            "import bench; a=bench.parse(); r=bench.cpu_arm(a,1,1,0);"
            "maps=open('/proc/self/maps').read(); print('LOADED' if 'librgbd_b200' in maps else 'CLEAN', r['value']>0)")
    root = __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.strip().splitlines()[-1] == "CLEAN True", out.stdout


def test_se_cache_range_bookkeeping():
    """engine._missing / _union / _subtract: which channels of the SE partial-sum table need refreshing."""
    from rgbd_b200.engine import _missing, _subtract, _union
    rng = random.Random(7)
    for _ in range(200):
        truth = [False] * 64
        ranges = []
        for _ in range(12):
            lo = rng.randrange(0, 63)
            hi = rng.randrange(lo + 1, 65)
            if rng.random() < 0.5:
                ranges = _union(ranges, lo, hi)
                truth[lo:hi] = [True] * (hi - lo)
            else:
                ranges = _subtract(ranges, lo, hi)
                truth[lo:hi] = [False] * (hi - lo)
            assert all(a < b for a, b in ranges) and all(ranges[i][1] <= ranges[i + 1][0] for i in range(len(ranges) - 1))
            got = [any(a <= c < b for a, b in ranges) for c in range(64)]
            assert got == truth
            qlo = rng.randrange(0, 63)
            qhi = rng.randrange(qlo + 1, 65)
            miss = _missing(ranges, qlo, qhi)
            assert [any(a <= c < b for a, b in miss) for c in range(64)] == [qlo <= c < qhi and not truth[c] for c in range(64)]


def test_concat_packing_of_skip_and_last_conv_equals_their_sum():
    """ELIC_united._pc_concat: one 1x1 conv over [x | t2] with weights [W_skip | W3] and summed biases == skip(x) + conv3(t2)
    (the widening ResidualBottleneck's last two launches as one)."""
    torch.manual_seed(3)
    net = rgbd_b200.ELIC_united(config=rgbd_b200.model_config(), channel=4)
    skip, last = nn.Conv2d(24, 16, 1), nn.Conv2d(8, 16, 1)
    pc = net._pc_concat(skip, last)
    assert pc.Cin == 32 and pc.Cout == 16 and pc.k == 1
    x, t2 = torch.randn(2, 24, 5, 7), torch.randn(2, 8, 5, 7)
    want = skip(x) + last(t2)
    w = pc.w32[0, :, :16]                                   # [Cin, Cout] of the single tap
    got = torch.einsum("nchw,co->nohw", torch.cat([x, t2], 1), w) + pc.bias.view(1, -1, 1, 1)
    assert torch.allclose(got, want, atol=1e-5)
    assert net._pc_concat(skip, last) is pc                 # cached per pair of modules


def test_linear_packing_for_the_token_layers_equals_torch():
    """STF_united._pc_linear: an nn.Linear over tokens == the 1x1 conv the launch plan runs over the pixels."""
    torch.manual_seed(4)
    net = rgbd_b200.STF_united(config=rgbd_b200.model_config(), channel=4)
    for lin in (nn.Linear(48, 144), nn.Linear(96, 48, bias=False)):
        pc = net._pc_linear(lin)
        assert (pc.Cin, pc.Cout, pc.k, pc.stride, pc.pad) == (lin.in_features, lin.out_features, 1, 1, 0)
        tok = torch.randn(3, 10, lin.in_features)
        got = tok @ pc.w32[0, :, :lin.out_features] + (pc.bias if pc.bias is not None else 0)
        assert torch.allclose(got, lin(tok), atol=1e-5)
    blk = net.g_a.rgb_ana_layers[0].blocks[1]
    assert blk.shift_size == 2 and blk.window_size == 4 and blk.num_heads == 3 and abs(blk.attn.scale - 16 ** -0.5) < 1e-12
