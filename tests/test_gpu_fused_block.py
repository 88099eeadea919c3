"""rgbd_rb_* (csrc/conv_rb.cu): ResidualBottleneck / ResidualUnit as one tcgen05 launch, against torch CPU on the same
bf16-rounded inputs / weights with the intermediates rounded to bf16 where the kernel rounds them (t1, t2 live in shared
memory as bf16).  Tolerance 6e-3 of the tensor scale = fp32 accumulation-order differences + one bf16 output rounding
(2^-8); the same block run as three unfused tensor-core launches must agree to the same tolerance."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rel_err(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def _bf(t):
    return t.bfloat16().float()


def _view(t_nchw, pad=8, wide=0):
    from rgbd_b200.engine import View
    nhwc = t_nchw.permute(0, 2, 3, 1).contiguous()
    Cp = nhwc.shape[3] + 2 * pad + wide
    buf = torch.zeros(*nhwc.shape[:3], Cp, device=DEV, dtype=torch.bfloat16)
    buf[..., pad:pad + nhwc.shape[3]] = nhwc.to(DEV).to(torch.bfloat16)
    return View(buf, pad, nhwc.shape[3])


def _reference(c1, c2, c3, x, res, final_relu):
    with torch.no_grad():
        t1 = _bf(F.relu(F.conv2d(_bf(x), _bf(c1.weight), c1.bias)))
        t2 = _bf(F.relu(F.conv2d(t1, _bf(c2.weight), c2.bias, padding=1)))
        y = F.conv2d(t2, _bf(c3.weight), c3.bias) + _bf(res)
        return F.relu(y) if final_relu else y


def _run(c1, c2, c3, x, res, final_relu, fused=True, same_res=False, pads=(8, 16, 24)):
    """pads: channel offsets of the x / res / y views inside wider buffers ((16, 32, 32): every pixel row 32-byte aligned,
    the kernel's 256-bit load / store path; the default exercises the 128-bit path)"""
    from rgbd_b200.engine import Builder, PackedConv
    dev = torch.device(DEV)
    b = Builder(dev, torch.bfloat16, tensor_cores=True)
    pcs = [PackedConv(m, dev) for m in (c1, c2, c3)]
    xv = _view(x, pad=pads[0])
    rv = xv if same_res else _view(res, pad=pads[1])
    out = _view(torch.zeros(x.shape[0], c3.out_channels, *x.shape[2:]), pad=pads[2], wide=40 if pads[2] == 24 else 32)
    if fused:
        assert b.can_fuse_block(*pcs, xv)
        b.fused_block(*pcs, xv, res=rv, out=out, final_relu=final_relu)
    else:
        t1 = b.conv(pcs[0], xv, act=1)
        t2 = b.conv(pcs[1], t1, act=1)
        b.conv(pcs[2], t2, out=out, res=rv, act=1 if final_relu else 0)
    b.prog.run()
    b.prog.run()     # a second launch of the same plan (barrier phases / scheduler state start clean every launch)
    torch.cuda.synchronize()
    full = out.buf.float().cpu()
    assert float(full[..., :pads[2]].abs().max()) == 0 and float(full[..., pads[2] + c3.out_channels:].abs().max()) == 0, \
        "the kernel wrote outside its channel view"
    return out.torch().float().cpu().permute(0, 3, 1, 2)


CASES = [
    # name, Cin, N, H, W, final_relu, residual is x itself
    ("rb192_small", 192, 2, 16, 20, False, True),
    ("rb192_ragged", 192, 1, 37, 53, False, True),           # tiles hang over the right / bottom border
    ("ru192_relu", 192, 2, 32, 40, True, True),
    ("rb384_skipres", 384, 1, 24, 44, False, False),         # RB(2N -> N): residual = output of the 1x1 skip conv
    ("rb192_many_tiles", 192, 3, 64, 80, False, True),       # > 148 tiles: every CTA walks several tiles
    ("rb192_multi_tile_per_cta", 192, 8, 64, 80, False, True),   # 192 tiles on 148 CTAs: barrier phases across tiles
    ("rb192_five_tiles_per_cta", 192, 2, 256, 320, True, True),   # 704 tiles
    ("rb192_tiny", 192, 1, 2, 2, False, True),
    ("rb192_tall_narrow", 192, 1, 70, 5, True, False),
]


@pytest.mark.parametrize("name,cin,n,h,w,final_relu,same_res", CASES, ids=[c[0] for c in CASES])
def test_fused_block_vs_torch(name, cin, n, h, w, final_relu, same_res):
    torch.manual_seed(21)
    c1, c2, c3 = nn.Conv2d(cin, 96, 1).eval(), nn.Conv2d(96, 96, 3, 1, 1).eval(), nn.Conv2d(96, 192, 1).eval()
    with torch.no_grad():
        c1.bias.add_(0.3)       # relu(b1) != 0: a wrong border (t1 not zero-padded) would show
    x = torch.randn(n, cin, h, w)
    res = x if same_res else torch.randn(n, 192, h, w)
    want = _reference(c1, c2, c3, x, res, final_relu)
    got = _run(c1, c2, c3, x, res, final_relu, fused=True, same_res=same_res)
    assert got.shape == want.shape
    assert rel_err(got, want) < 6e-3, (name, rel_err(got, want))
    wide = _run(c1, c2, c3, x, res, final_relu, fused=True, same_res=same_res, pads=(16, 32, 32))
    assert torch.equal(wide, got), "the 256-bit and the 128-bit load / store paths must give the same bits"
    unfused = _run(c1, c2, c3, x, res, final_relu, fused=False, same_res=same_res)
    assert rel_err(got, unfused) < 6e-3, (name, rel_err(got, unfused))


def test_fused_block_is_deterministic_and_batch_invariant():
    torch.manual_seed(22)
    c1, c2, c3 = nn.Conv2d(192, 96, 1).eval(), nn.Conv2d(96, 96, 3, 1, 1).eval(), nn.Conv2d(96, 192, 1).eval()
    x = torch.randn(3, 192, 40, 48)
    a = _run(c1, c2, c3, x, x, False, same_res=True)
    b = _run(c1, c2, c3, x, x, False, same_res=True)
    assert torch.equal(a, b)
    one = _run(c1, c2, c3, x[1:2], x[1:2], False, same_res=True)
    assert torch.equal(one[0], a[1])


def test_model_with_fused_blocks_matches_unfused_model():
    """The whole codec with fused bottleneck blocks: same symbols rate (within 0.5 %) as the unfused launch list, exact
    round trip, reconstruction as close to the oracle's g_s on the same y_hat as bf16 arithmetic allows, and it really
    runs fewer launches."""
    import rgbd_b200
    from gpu_utils import check_recon_fidelity, make_model, nchw
    from oracle.bf16_emulation import Bf16OracleCodec
    from oracle.model_oracle import OracleCodec
    from rgbd_b200.synthetic import synthetic_pairs
    rgb, depth = synthetic_pairs(2, 128, 192, seed=5)
    outs = {}
    for fuse in (True, False):
        net, sd = make_model(rgbd_b200.ELIC_united, "mid", 0, precision="bf16", fuse_blocks=fuse)
        c = net.compress(rgb.to(DEV), depth.to(DEV))
        enc = net._program("encoder", 2, 128, 192)
        sym = {k: enc.io["st"][k]["ysym"].clone() for k in ("r", "d")}
        r = net.decompress(c["r_strings"], c["d_strings"], c["shape"])
        dec = net._program("decoder", 2, 2, 3)
        for k in ("r", "d"):
            assert torch.equal(dec.io["st"][k]["ysym"], sym[k]), (fuse, k)
        yh = [nchw(dec.io["yhat"][k]) for k in ("r", "d")]
        gs, gs_emu = (dict(zip(("r", "d"), o.g_s(*yh))) for o in (OracleCodec(sd), Bf16OracleCodec(sd)))
        for k in ("r", "d"):
            check_recon_fidelity(f"fused={fuse} {k}", nchw(dec.io["x_nhwc"][k]), gs[k], gs_emu[k])
        outs[fuse] = (c, r, len(enc.ops) + len(dec.ops), len(enc.rb_plans) + len(dec.rb_plans))
    assert outs[True][3] > 0 and outs[False][3] == 0 and outs[True][2] < outs[False][2]
    for key in ("r_strings", "d_strings"):
        a = sum(len(s) for g in outs[True][0][key] for s in g)
        b = sum(len(s) for g in outs[False][0][key] for s in g)
        assert abs(a - b) <= 0.005 * b, (key, a, b)
