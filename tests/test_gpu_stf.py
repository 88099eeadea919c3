"""STF_united (models/stf_united.py) on the CUDA path: the token-side kernels against torch, the two
transforms against the oracle, and the three parity levels of the codec against the reference golden."""
import ctypes as C
import json

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import rgbd_b200
from oracle import coder
from oracle.stf_oracle import StfOracle
from rgbd_b200 import lib as L
from rgbd_b200.synthetic import synthetic_pairs, synthetic_state_dict

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def sp():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def rel_err(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


def _net(precision="fp32", **kw):
    net = rgbd_b200.STF_united(config=rgbd_b200.model_config(), channel=4, precision=precision, **kw).eval()
    net.load_state_dict(synthetic_state_dict(net, 0, "mid"))
    net.update(force=True)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    return net.to(DEV), sd


def test_layernorm_and_patch_merging_gather():
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 6, 10, 48 + 8, generator=g) * 3 + 1
    gamma, beta = torch.rand(48, generator=g) + 0.5, torch.randn(48, generator=g)
    xd, gd, bd = x.to(DEV), gamma.to(DEV), beta.to(DEV)        # (named: a temporary would be freed before the launch)
    y = torch.zeros(2, 6, 10, 64, device=DEV)
    L.call("rgbd_layernorm", xd.data_ptr(), L.DT_F32, y.data_ptr(), L.DT_F32, 2 * 6 * 10, 48, 56, 8, 64, 16, gd.data_ptr(),
           bd.data_ptr(), 1e-5, 0, 6, 10, sp())
    want = F.layer_norm(x[..., 8:], (48,), gamma, beta, 1e-5)
    assert torch.allclose(y[..., 16:].cpu(), want, atol=2e-5) and float(y[..., :16].abs().max()) == 0
    # PatchMerging: [x(2i,2j) | x(2i+1,2j) | x(2i,2j+1) | x(2i+1,2j+1)] then LayerNorm(4C)
    g4, b4 = torch.rand(192, generator=g) + 0.5, torch.randn(192, generator=g)
    xm = x[..., 8:]
    cat = torch.cat([xm[:, 0::2, 0::2], xm[:, 1::2, 0::2], xm[:, 0::2, 1::2], xm[:, 1::2, 1::2]], -1)
    ym = torch.zeros(2, 3, 5, 192, device=DEV, dtype=torch.bfloat16)
    g4d, b4d = g4.to(DEV), b4.to(DEV)
    L.call("rgbd_layernorm", xd.data_ptr(), L.DT_F32, ym.data_ptr(), L.DT_BF16, 2 * 3 * 5, 192, 56, 8, 192, 0, g4d.data_ptr(),
           b4d.data_ptr(), 1e-5, 1, 6, 10, sp())
    assert rel_err(ym.float().cpu(), F.layer_norm(cat, (192,), g4, b4, 1e-5)) < 6e-3        # one bf16 rounding


def test_pixel_shuffle_matches_torch():
    x = torch.randn(2, 5, 7, 4 * 24)
    y = torch.zeros(2, 10, 14, 24, device=DEV)
    xd = x.to(DEV)
    L.call("rgbd_pixel_shuffle2", xd.data_ptr(), y.data_ptr(), L.DT_F32, 2, 5, 7, 24, 96, 0, 24, 0, sp())
    want = F.pixel_shuffle(x.permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1)
    assert torch.equal(y.cpu(), want)


@pytest.mark.parametrize("shift", [0, 2])
def test_window_attention_matches_reference_arithmetic(shift):
    """q k^T * scale + relative position bias (+ the -100 region mask on the rolled map), softmax, times v — against the
    oracle's restatement of WindowAttention inside SwinTransformerBlock (stf_united.py:83-115, 162-212)."""
    g = torch.Generator().manual_seed(7 + shift)
    B, H, W, Cc, heads, ws = 2, 8, 12, 48, 3, 4
    qkv = torch.randn(B, H, W, 3 * Cc, generator=g)
    table = torch.randn((2 * ws - 1) ** 2, heads, generator=g) * 0.5
    out = torch.zeros(B, H, W, Cc, device=DEV)
    qd, td = qkv.to(DEV), table.to(DEV)
    L.call("rgbd_window_attention", qd.data_ptr(), out.data_ptr(), L.DT_F32, B, H, W, Cc, heads, ws, shift,
           td.data_ptr(), float((Cc // heads) ** -0.5), 3 * Cc, 0, Cc, 0, sp())
    # reference arithmetic on the CPU
    from rgbd_b200.modules_stf import WindowAttention
    idx = WindowAttention(Cc, ws, heads).relative_position_index.view(-1)
    t = torch.roll(qkv, shifts=(-shift, -shift), dims=(1, 2)) if shift else qkv
    win = StfOracle._windows(t, ws)
    q, k, v = win.reshape(win.shape[0], ws * ws, 3, heads, Cc // heads).permute(2, 0, 3, 1, 4)
    attn = (q * (Cc // heads) ** -0.5) @ k.transpose(-2, -1) + table[idx].view(ws * ws, ws * ws, -1).permute(2, 0, 1).unsqueeze(0)
    if shift:
        m = StfOracle._shift_mask(H, W, ws, shift)
        attn = (attn.view(B, m.shape[0], heads, ws * ws, ws * ws) + m.unsqueeze(1).unsqueeze(0)).view(-1, heads, ws * ws, ws * ws)
    o = (attn.softmax(-1) @ v).transpose(1, 2).reshape(win.shape[0], ws * ws, Cc)
    want = StfOracle._unwindows(o, ws, B, H, W)
    if shift:
        want = torch.roll(want, shifts=(shift, shift), dims=(1, 2))
    assert rel_err(out.cpu(), want) < 2e-5


def test_stf_transforms_match_oracle_fp32():
    net, sd = _net()
    orc = StfOracle(sd)
    rgb, depth = synthetic_pairs(1, 256, 320, seed=9)
    net.compress(rgb.to(DEV), depth.to(DEV))
    prog = net._program("encoder", 1, 256, 320)
    yr, yd = orc.g_a(rgb, depth)
    for m, want in (("r", yr), ("d", yd)):
        got = prog.io["y"][m].torch().float().cpu().permute(0, 3, 1, 2)
        assert rel_err(got, want) < 2e-4, (m, rel_err(got, want))


def test_stf_matches_reference_golden(golden_dir):
    g = np.load(f"{golden_dir}/model_stf_united.npz")
    meta = json.loads(str(g["meta"]))
    net, sd = _net()
    orc = StfOracle(sd)
    rgb, depth = synthetic_pairs(1, meta["H"], meta["W"], seed=meta["input_seed"])
    out = net.compress(rgb.to(DEV), depth.to(DEV))
    assert tuple(out["shape"]) == tuple(g["shape"])
    prog = net._program("encoder", 1, meta["H"], meta["W"])
    tr = orc.compress(rgb, depth, trace=True)["_trace"]
    for which, key, name in (("r", "r_strings", "rgb"), ("d", "d_strings", "depth")):
        st = prog.io["st"][which]
        ysym, yidx = st["ysym"][0].cpu().numpy(), st["yidx"][0].cpu().numpy().astype(np.int32)
        # level 1: bytes == the oracle coder on the GPU's own symbols; the z string equals the reference's byte for byte
        assert out[key][0][0] == coder.encode_with_indexes(ysym, yidx, orc.gc_tables(name))
        assert out[key][1][0] == g[which + "z_bytes"].tobytes()
        # level 2: symbols equal the reference's except at rounding boundaries, bpp within 0.5 %
        want_sym, want_idx = tr["symbols"][(name, 0)]
        assert (ysym != want_sym).mean() < 2e-3 and np.abs(ysym - want_sym).max() <= 1 and (yidx != want_idx).mean() < 2e-3
        ref_bytes = len(g[which + "y_bytes"]) + len(g[which + "z_bytes"])
        got_bytes = len(out[key][0][0]) + len(out[key][1][0])
        assert abs(got_bytes - ref_bytes) <= 0.005 * ref_bytes
    # level 3 + reconstruction against the reference's
    rec = net.decompress(out["r_strings"], out["d_strings"], out["shape"])
    dec = net._program("decoder", 1, int(out["shape"][0]), int(out["shape"][1]))
    for which in ("r", "d"):
        assert torch.equal(dec.io["st"][which]["ysym"], prog.io["st"][which]["ysym"])
        ref = torch.from_numpy(g["xhat_" + which])
        assert float(((rec["x_hat"][which].cpu() - ref) ** 2).mean()) < 1e-5
    # forward
    f = net(rgb.to(DEV), depth.to(DEV))
    for which in ("r", "d"):
        assert float(((f["x_hat"][which].cpu() - torch.from_numpy(g["fwd_xhat_" + which])) ** 2).mean()) < 1e-5
        got, want = f[which + "_likelihoods"]["y"].cpu(), g["lik_y_" + which]
        bits_g, bits_w = float(-torch.log2(got).sum()), float(-np.log2(want).sum())
        assert abs(bits_g - bits_w) / bits_w < 0.005


def test_stf_bf16_roundtrip_and_fidelity():
    """bf16 tensor-core mode: exact round trip, rate within 0.5 % of the fp32 oracle, reconstruction as close to the oracle's
    g_s on the same y_hat as bf16 arithmetic allows (gpu_utils.check_recon_fidelity), batch of 2."""
    from gpu_utils import check_recon_fidelity, nchw
    from oracle.bf16_emulation import Bf16StfOracle
    net, sd = _net("bf16")
    orc = StfOracle(sd)
    rgb, depth = synthetic_pairs(2, 256, 256, seed=31)
    out = net.compress(rgb.to(DEV), depth.to(DEV))
    enc = net._program("encoder", 2, 256, 256)
    assert enc.n_tc > 0
    rec = net.decompress(out["r_strings"], out["d_strings"], out["shape"])
    dec = net._program("decoder", 2, 4, 4)
    ref_c = orc.compress(rgb[:1], depth[:1])
    for which, key in (("r", "r_strings"), ("d", "d_strings")):
        assert torch.equal(dec.io["st"][which]["ysym"], enc.io["st"][which]["ysym"])
        want = len(ref_c[key][0][0]) + len(ref_c[key][1][0])
        got = len(out[key][0][0]) + len(out[key][1][0])
        assert abs(got - want) <= 0.005 * want, (which, got, want)
    yh = [nchw(dec.io["yhat"][k])[:1] for k in ("r", "d")]
    gs, gs_emu = orc.g_s(*yh), Bf16StfOracle(sd).g_s(*yh)
    for i, k in enumerate(("r", "d")):
        got_db, emu_db = check_recon_fidelity("STF " + k, nchw(dec.io["x_nhwc"][k])[:1], gs[i], gs_emu[i], floor_db=None)
        print(f"[STF fidelity {k}] {got_db:.2f} dB (bf16 emulation {emu_db:.2f} dB)")
    assert rec["x_hat"]["r"].shape == (2, 3, 256, 256) and torch.isfinite(rec["x_hat"]["d"]).all()
