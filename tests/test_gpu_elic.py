"""Single-modality ELIC (models/elic.py, testing/tester_single.py) on the CUDA path against the oracle
and the reference golden: the three parity levels of the united model, for channel = 3 and channel = 1."""
import json

import numpy as np
import pytest
import torch

import rgbd_b200
from oracle import coder
from oracle.elic_oracle import ElicOracle
from rgbd_b200.synthetic import synthetic_pairs, synthetic_state_dict

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _net(channel, precision="fp32", **kw):
    net = rgbd_b200.ELIC(config=rgbd_b200.model_config(), channel=channel, precision=precision, **kw).eval()
    net.load_state_dict(synthetic_state_dict(net, 0, "mid"))
    net.update(force=True)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    return net.to(DEV), ElicOracle(sd)


def _psnr(a, b):
    mse = float(((a.double() - b.double()) ** 2).mean())
    return 99.0 if mse == 0 else 10 * np.log10(1.0 / mse)


@pytest.mark.parametrize("name,channel", [("rgb", 3), ("depth", 1)])
def test_elic_matches_reference_golden(golden_dir, name, channel):
    g = np.load(f"{golden_dir}/model_elic_{name}.npz")
    meta = json.loads(str(g["meta"]))
    net, orc = _net(channel)
    rgb, depth = synthetic_pairs(1, meta["H"], meta["W"], seed=meta["input_seed"])
    x = rgb if channel == 3 else depth
    out = net.compress(x.to(DEV))
    assert tuple(out["shape"]) == tuple(g["shape"]) and len(out["strings"][0]) == 1 and len(out["strings"][1]) == 1
    prog = net._program("encoder", 1, meta["H"], meta["W"])
    # level 1: bytes == the oracle coder on the GPU's own symbols; the z string equals the reference's byte for byte
    ysym, yidx = prog.io["ysym"][0].cpu().numpy(), prog.io["yidx"][0].cpu().numpy().astype(np.int32)
    assert out["strings"][0][0] == coder.encode_with_indexes(ysym, yidx, orc.gc_tables("x"))
    zsym, zidx = prog.io["zsym"][0].cpu().numpy(), prog.io["zidx"][0].cpu().numpy().astype(np.int32)
    assert out["strings"][1][0] == coder.encode_with_indexes(zsym, zidx, orc.eb_tables("x"))
    # level 2: symbols equal the reference's except at rounding boundaries, bpp within 0.5 %
    tr = orc.compress(x, trace=True)["_trace"]
    want_sym, want_idx = tr["symbols"][0]
    assert (ysym != want_sym).mean() < 2e-3 and (yidx != want_idx).mean() < 2e-3 and np.abs(ysym - want_sym).max() <= 1
    ref_bytes = len(g["y_bytes"]) + len(g["z_bytes"])
    got_bytes = len(out["strings"][0][0]) + len(out["strings"][1][0])
    assert abs(got_bytes - ref_bytes) <= 0.005 * ref_bytes
    y = prog.io["y"].torch().float().cpu().permute(0, 3, 1, 2)
    assert float((y - torch.from_numpy(g["y"])).abs().max() / np.abs(g["y"]).max()) < 1e-4
    # level 3 + reconstruction
    rec = net.decompress(out["strings"], out["shape"])
    dec = net._program("decoder", 1, int(out["shape"][0]), int(out["shape"][1]))
    assert torch.equal(dec.io["ysym"], prog.io["ysym"]) and torch.equal(dec.io["yhat"].torch(), prog.io["yhat"].torch())
    ref = torch.from_numpy(g["xhat"])
    assert abs(_psnr(rec["x_hat"].cpu().clamp(0, 1), x) - _psnr(ref.clamp(0, 1), x)) < 0.05
    assert float(((rec["x_hat"].cpu() - ref) ** 2).mean()) < 1e-5
    # forward
    f = net(x.to(DEV))
    assert float(((f["x_hat"].cpu() - torch.from_numpy(g["fwd_xhat"])) ** 2).mean()) < 1e-5
    for k, want in (("y_likelihoods", g["lik_y"]), ("z_likelihoods", g["lik_z"])):
        got = f["likelihoods"][k].cpu()
        bits_g, bits_w = float(-torch.log2(got).sum()), float(-np.log2(want).sum())
        assert got.shape == want.shape and abs(bits_g - bits_w) / bits_w < 0.005, k


@pytest.mark.parametrize("channel", [3, 1])
def test_elic_bf16_batch_and_multistream(channel):
    net, orc = _net(channel, precision="bf16")
    rgb, depth = synthetic_pairs(3, 128, 128, seed=12)
    x = rgb if channel == 3 else depth
    out = net.compress(x.to(DEV))
    assert len(out["strings"][0]) == 3 and len(out["strings"][1]) == 3
    rec = net.decompress(out["strings"], out["shape"])
    enc, dec = net._program("encoder", 3, 128, 128), net._program("decoder", 3, 2, 2)
    assert torch.equal(dec.io["ysym"], enc.io["ysym"])
    assert rec["x_hat"].shape == (3, channel, 128, 128) and torch.isfinite(rec["x_hat"]).all()
    # rate and reconstruction close to the fp32 oracle
    ref_c = orc.compress(x[:1])
    want = len(ref_c["strings"][0][0]) + len(ref_c["strings"][1][0])
    got = len(out["strings"][0][0]) + len(out["strings"][1][0])
    assert abs(got - want) <= 0.005 * want, (got, want)
    # reconstruction: against the oracle's g_s on the GPU's own y_hat (gpu_utils.recon_fidelity_db says why)
    from gpu_utils import check_recon_fidelity, nchw
    from oracle.bf16_emulation import Bf16ElicOracle
    yh = nchw(dec.io["yhat"])
    check_recon_fidelity(f"ELIC channel {channel}", rec["x_hat"], orc.g_s1(yh), Bf16ElicOracle(orc.sd).g_s1(yh))
    # per-image bytes do not depend on the batch
    one = net.compress(x[1:2].to(DEV))
    assert one["strings"][0][0] == out["strings"][0][1] and one["strings"][1][0] == out["strings"][1][1]
    # multi-stream layout: every sub-stream is the reference coder's string for its symbols; same reconstruction
    net2, _ = _net(channel, precision="bf16", stream_layout="multi", sub_channels=8)
    m = net2.compress(x.to(DEV))
    nsub = 2 * 320 // 8
    assert len(m["strings"][0]) == 3 * nsub
    e2 = net2._program("encoder", 3, 128, 128)
    ysym, yidx = e2.io["ysym"].cpu().numpy(), e2.io["yidx"].cpu().numpy().astype(np.int32)
    sublen = e2.io["sublen"]
    t = orc.gc_tables("x")
    for i in (0, 2):
        for k in range(0, nsub, 7):
            sl = slice(k * sublen, (k + 1) * sublen)
            assert m["strings"][0][i * nsub + k] == coder.encode_with_indexes(ysym[i, sl], yidx[i, sl], t), (i, k)
    rec2 = net2.decompress(m["strings"], m["shape"])
    assert torch.equal(rec2["x_hat"], rec["x_hat"])
