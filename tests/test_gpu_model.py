"""End-to-end parity of the CUDA ELIC_united path against the oracle (BASELINE north_star's three
levels) at sizes the oracle finishes in seconds, plus size-independent properties at the
benchmark size (480x640 -> 512x640)."""
import json

import numpy as np
import pytest
import torch

import rgbd_b200
from oracle import coder
from oracle.model_oracle import OracleCodec
from rgbd_b200.synthetic import pad_to_multiple, synthetic_pairs

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def bpp(strings, npix):
    return sum(len(s) for grp in strings for s in grp) * 8.0 / npix


def nhwc_to_nchw(view):
    return view.torch().float().cpu().permute(0, 3, 1, 2)


@pytest.fixture(scope="module")
def united():
    from gpu_utils import make_model
    net, sd = make_model(rgbd_b200.ELIC_united, "mid", 0)
    return net, OracleCodec(sd)


def _check_bytes_against_oracle_coder(net, orc, out, B, H, W):
    """Level 1: every stream is byte-identical to the reference coder run on the symbols and
    indexes the GPU path produced."""
    prog = net._program("encoder", B, H, W)
    for which, key, name in (("r", "r_strings", "rgb"), ("d", "d_strings", "depth")):
        st = prog.io["st"][which]
        ysym, yidx = st["ysym"].cpu().numpy(), st["yidx"].cpu().numpy().astype(np.int32)
        zsym, zidx = st["zsym"].cpu().numpy(), st["zidx"].cpu().numpy().astype(np.int32)
        gt, et = orc.gc_tables(name), orc.eb_tables(name)
        for i in range(B):
            assert out[key][0][i] == coder.encode_with_indexes(ysym[i], yidx[i], gt), (which, "y", i)
            assert out[key][1][i] == coder.encode_with_indexes(zsym[i], zidx[i], et), (which, "z", i)


def test_transforms_match_oracle(united):
    net, orc = united
    rgb, depth = synthetic_pairs(1, 128, 128, seed=4321)
    net.compress(rgb.to(DEV), depth.to(DEV))
    prog = net._program("encoder", 1, 128, 128)
    yr, yd = orc.g_a(rgb, depth)
    zr, zd = orc.h_a(yr, yd)
    for name, view, want in (("y_r", prog.io["y"]["r"], yr), ("y_d", prog.io["y"]["d"], yd),
                             ("z_r", prog.io["z"]["r"], zr), ("z_d", prog.io["z"]["d"], zd)):
        got = nhwc_to_nchw(view)
        err = float((got - want).abs().max() / want.abs().max())
        assert err < 1e-4, (name, err)     # fp32: ~100 layers of differently-ordered fp32 sums


def test_compress_matches_reference_golden(united, golden_dir):
    """Levels 1+2 against the UNMODIFIED reference's output on the same weights and input."""
    net, orc = united
    g = np.load(f"{golden_dir}/model_united.npz")
    meta = json.loads(str(g["meta"]))
    rgb, depth = synthetic_pairs(1, meta["H"], meta["W"], seed=meta["input_seed"])
    out = net.compress(rgb.to(DEV), depth.to(DEV))
    assert tuple(out["shape"]) == tuple(g["shape"])
    _check_bytes_against_oracle_coder(net, orc, out, 1, meta["H"], meta["W"])
    npix = meta["H"] * meta["W"]
    for key, tag in (("r_strings", "r"), ("d_strings", "d")):
        ref_bpp = (len(g[f"{tag}_y"]) + len(g[f"{tag}_z"])) * 8.0 / npix
        got_bpp = bpp(out[key], npix)
        assert abs(got_bpp - ref_bpp) / ref_bpp < 0.005, (tag, got_bpp, ref_bpp)      # bpp within 0.5 %
    # symbols equal the reference's except where y - mu sits on a rounding boundary
    tr = orc.compress(rgb, depth, trace=True)["_trace"]
    prog = net._program("encoder", 1, meta["H"], meta["W"])
    for which, name in (("r", "rgb"), ("d", "depth")):
        want_sym, want_idx = tr["symbols"][(name, 0)]
        got_sym = prog.io["st"][which]["ysym"][0].cpu().numpy()
        got_idx = prog.io["st"][which]["yidx"][0].cpu().numpy()
        assert (got_sym != want_sym).mean() < 2e-3, (which, (got_sym != want_sym).mean())
        assert (got_idx != want_idx).mean() < 2e-3
        assert np.abs(got_sym - want_sym).max() <= 1
    rec = net.decompress(out["r_strings"], out["d_strings"], out["shape"])
    from gpu_utils import psnr
    for m, tag in (("r", "xhat_r"), ("d", "xhat_d")):
        ref = torch.from_numpy(g[tag])
        x = rgb if m == "r" else depth
        assert abs(psnr(rec["x_hat"][m].cpu(), x) - psnr(ref, x)) < 0.05, m           # PSNR within 0.05 dB
        assert rec["x_hat"][m].min() >= 0 and rec["x_hat"][m].max() <= 1


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_roundtrip_every_symbol_and_batch_invariance(precision):
    """Level 3: decompress(compress(x)) reproduces every symbol; a batch of 3 gives the same bytes per
    image as 3 single-image calls (determinism across batch size)."""
    from gpu_utils import make_model
    net, sd = make_model(rgbd_b200.ELIC_united, "mid", 0, precision=precision)
    orc = OracleCodec(sd)
    rgb, depth = synthetic_pairs(3, 128, 192, seed=77)
    out = net.compress(rgb.to(DEV), depth.to(DEV))
    assert [len(g) for g in out["r_strings"]] == [3, 3]
    _check_bytes_against_oracle_coder(net, orc, out, 3, 128, 192)
    enc = net._program("encoder", 3, 128, 192)
    enc_hat = {k: enc.io["yhat"][k].torch().clone() for k in ("r", "d")}
    enc_sym = {k: enc.io["st"][k]["ysym"].clone() for k in ("r", "d")}
    rec = net.decompress(out["r_strings"], out["d_strings"], out["shape"])
    dec = net._program("decoder", 3, int(out["shape"][0]), int(out["shape"][1]))
    for k in ("r", "d"):
        assert torch.equal(dec.io["st"][k]["ysym"], enc_sym[k]), k
        assert torch.equal(dec.io["yhat"][k].torch(), enc_hat[k]), k
    assert rec["x_hat"]["r"].shape == (3, 3, 128, 192) and rec["x_hat"]["d"].shape == (3, 1, 128, 192)
    assert torch.isfinite(rec["x_hat"]["r"]).all()
    for i in range(3):
        one = net.compress(rgb[i:i + 1].to(DEV), depth[i:i + 1].to(DEV))
        for key in ("r_strings", "d_strings"):
            assert one[key][0][0] == out[key][0][i] and one[key][1][0] == out[key][1][i], (i, key)
        rec1 = net.decompress(one["r_strings"], one["d_strings"], one["shape"])
        assert torch.equal(rec1["x_hat"]["r"][0], rec["x_hat"]["r"][i])
        assert torch.equal(rec1["x_hat"]["d"][0], rec["x_hat"]["d"][i])


def test_bf16_rate_and_quality_close_to_fp32_oracle():
    from gpu_utils import make_model, psnr
    net, sd = make_model(rgbd_b200.ELIC_united, "mid", 0, precision="bf16")
    orc = OracleCodec(sd)
    rgb, depth = synthetic_pairs(1, 128, 128, seed=4321)
    out = net.compress(rgb.to(DEV), depth.to(DEV))
    rec = net.decompress(out["r_strings"], out["d_strings"], out["shape"])
    ref_c = orc.compress(rgb, depth)
    ref = orc.decompress(ref_c["r_strings"], ref_c["d_strings"], ref_c["shape"])
    npix = 128 * 128
    for key, m, x in (("r_strings", "r", rgb), ("d_strings", "d", depth)):
        b0, b1 = bpp(ref_c[key], npix), bpp(out[key], npix)
        assert abs(b1 - b0) / b0 < 0.005, (m, b0, b1)
        assert abs(psnr(rec["x_hat"][m].cpu(), x) - psnr(ref["x_hat"][m], x)) < 0.05, m


def test_forward_matches_oracle(united):
    net, orc = united
    rgb, depth = synthetic_pairs(2, 128, 128, seed=9)
    got = net(rgb.to(DEV), depth.to(DEV))
    want = orc.forward(rgb, depth)
    for m in ("r", "d"):
        assert got["x_hat"][m].shape == want["x_hat"][m].shape
        mse = float(((got["x_hat"][m].cpu() - want["x_hat"][m]) ** 2).mean())
        assert mse < 1e-5, (m, mse)
    for side in ("r_likelihoods", "d_likelihoods"):
        for k in ("y", "z"):
            g_, w_ = got[side][k].cpu(), want[side][k]
            assert g_.shape == w_.shape and float(g_.min()) > 0
            bits_g, bits_w = float(-torch.log2(g_).sum()), float(-torch.log2(w_).sum())
            assert abs(bits_g - bits_w) / bits_w < 0.005, (side, k, bits_g, bits_w)


def test_r2d_variant_roundtrip_and_golden(golden_dir):
    from gpu_utils import make_model, psnr
    net, sd = make_model(rgbd_b200.ELIC_united_R2D, "mid", 0)
    orc = OracleCodec(sd, cross=False)
    g = np.load(f"{golden_dir}/model_r2d.npz")
    meta = json.loads(str(g["meta"]))
    rgb, depth = synthetic_pairs(1, meta["H"], meta["W"], seed=meta["input_seed"])
    out = net.compress(rgb.to(DEV), depth.to(DEV))
    _check_bytes_against_oracle_coder(net, orc, out, 1, meta["H"], meta["W"])
    npix = meta["H"] * meta["W"]
    for key, tag in (("r_strings", "r"), ("d_strings", "d")):
        ref_bpp = (len(g[f"{tag}_y"]) + len(g[f"{tag}_z"])) * 8.0 / npix
        assert abs(bpp(out[key], npix) - ref_bpp) / ref_bpp < 0.005, tag
    rec = net.decompress(out["r_strings"], out["d_strings"], out["shape"])
    for m, tag, x in (("r", "xhat_r", rgb), ("d", "xhat_d", depth)):
        assert abs(psnr(rec["x_hat"][m].cpu(), x) - psnr(torch.from_numpy(g[tag]), x)) < 0.05, m


def test_nyuv2_size_roundtrip_properties():
    """BASELINE configs[1]: one 480x640 pair padded to 512x640; size-independent checks."""
    from gpu_utils import make_model
    net, sd = make_model(rgbd_b200.ELIC_united, "realistic", 0)
    orc = OracleCodec(sd)
    rgb, depth = synthetic_pairs(1, 480, 640, seed=5)
    rgb, depth = pad_to_multiple(rgb), pad_to_multiple(depth)
    assert rgb.shape[-2:] == (512, 640)
    out = net.compress(rgb.to(DEV), depth.to(DEV))
    assert tuple(out["shape"]) == (8, 10)
    _check_bytes_against_oracle_coder(net, orc, out, 1, 512, 640)
    enc = net._program("encoder", 1, 512, 640)
    assert enc.io["ny"] == 409600 and enc.io["nz"] == 15360
    sym = {k: enc.io["st"][k]["ysym"].clone() for k in ("r", "d")}
    rec = net.decompress(out["r_strings"], out["d_strings"], out["shape"])
    dec = net._program("decoder", 1, 8, 10)
    for k in ("r", "d"):
        assert torch.equal(dec.io["st"][k]["ysym"], sym[k])
    # stream length is a whole number of words and >= 8 bytes; streams are not degenerate
    for key in ("r_strings", "d_strings"):
        for grp in out[key]:
            assert all(len(s) % 4 == 0 and len(s) >= 8 for s in grp)
        assert len(out[key][0][0]) > 1000
    # idempotence: compressing again gives the same bytes
    out2 = net.compress(rgb.to(DEV), depth.to(DEV))
    assert out2["r_strings"] == out["r_strings"] and out2["d_strings"] == out["d_strings"]
    # dense conv flop count of the plan == the survey's algorithmic figure (721.2 + 785.1 GFLOP)
    assert abs(enc.flops / 1e9 - 721.2) < 1.0, enc.flops / 1e9
    assert abs(dec.flops / 1e9 - 785.1) < 1.0, dec.flops / 1e9


def test_pipeline_matches_serial_calls():
    """RoundTripPipeline (several compress + decompress jobs in flight on separate streams and launch plans)
    returns, job by job, exactly what serial compress() / decompress() calls return."""
    from gpu_utils import make_model
    from rgbd_b200.pipeline import RoundTripPipeline
    net, _ = make_model(rgbd_b200.ELIC_united, "mid", 0, precision="bf16")
    jobs = []
    for j in range(5):
        rgb, depth = synthetic_pairs(2, 128, 128, seed=100 + j)
        jobs.append((rgb.to(DEV), depth.to(DEV)))
    want = []
    for rgb, depth in jobs:
        c = net.compress(rgb, depth)
        r = net.decompress(c["r_strings"], c["d_strings"], c["shape"])
        want.append((c, r["x_hat"]["r"].clone(), r["x_hat"]["d"].clone()))
    seen = {}

    def sink(j, slot, stream, x_r, x_d):
        seen[j] = (x_r.clone(), x_d.clone())

    res = RoundTripPipeline(net, 2).run(jobs, sink=sink, keep_last=5)
    torch.cuda.synchronize()
    assert [j for j, _, _ in res] == [0, 1, 2, 3, 4] and sorted(seen) == [0, 1, 2, 3, 4]
    for j, c, _ in res:
        assert c["r_strings"] == want[j][0]["r_strings"] and c["d_strings"] == want[j][0]["d_strings"], j
        assert torch.equal(seen[j][0], want[j][1]) and torch.equal(seen[j][1], want[j][2]), j


@pytest.mark.parametrize("cls_name,H,W", [("ELIC_united_R2D", 530, 730), ("ELIC_united", 1080, 1920)])
def test_other_baseline_shapes_roundtrip(cls_name, H, W):
    """BASELINE configs[3] / [4]: SUN RGB-D-shaped 530x730 through ELIC_united_R2D (padded to 576x768) and a
    1080x1920 pair (padded to 1088x1920) in the bf16 tensor-core mode.  Size-independent properties: the
    decoder reproduces every encoder symbol, every stream is byte-identical to the oracle coder on the
    GPU's symbols, compress is idempotent, the reconstruction is finite and in [0, 1]."""
    from gpu_utils import make_model
    cls = getattr(rgbd_b200, cls_name)
    net, sd = make_model(cls, "realistic", 0, precision="bf16")
    orc = OracleCodec(sd, cross=cls_name == "ELIC_united")
    rgb, depth = synthetic_pairs(1, H, W, seed=9)
    rgb, depth = pad_to_multiple(rgb), pad_to_multiple(depth)
    Hp, Wp = rgb.shape[-2:]
    assert Hp % 64 == 0 and Wp % 64 == 0
    out = net.compress(rgb.to(DEV), depth.to(DEV))
    assert tuple(out["shape"]) == (Hp // 64, Wp // 64)
    _check_bytes_against_oracle_coder(net, orc, out, 1, Hp, Wp)
    enc = net._program("encoder", 1, Hp, Wp)
    assert enc.io["ny"] == 320 * (Hp // 16) * (Wp // 16)
    sym = {k: enc.io["st"][k]["ysym"].clone() for k in ("r", "d")}
    rec = net.decompress(out["r_strings"], out["d_strings"], out["shape"])
    dec = net._program("decoder", 1, Hp // 64, Wp // 64)
    for k in ("r", "d"):
        assert torch.equal(dec.io["st"][k]["ysym"], sym[k]), k
    for m, c in (("r", 3), ("d", 1)):
        x = rec["x_hat"][m]
        assert x.shape == (1, c, Hp, Wp) and torch.isfinite(x).all() and x.min() >= 0 and x.max() <= 1
    out2 = net.compress(rgb.to(DEV), depth.to(DEV))
    assert out2["r_strings"] == out["r_strings"] and out2["d_strings"] == out["d_strings"]


def test_cuda_graph_replay_gives_the_same_bytes():
    """The static launch list replayed as a CUDA graph (bench default) == the eager launch list."""
    from gpu_utils import make_model
    from rgbd_b200 import lib as L
    net, _ = make_model(rgbd_b200.ELIC_united, "mid", 0, precision="bf16")
    rgb, depth = synthetic_pairs(2, 128, 192, seed=31)
    rgb, depth = rgb.to(DEV), depth.to(DEV)
    eager = net.compress(rgb, depth)
    rec_e = net.decompress(eager["r_strings"], eager["d_strings"], eager["shape"])
    net2, _ = make_model(rgbd_b200.ELIC_united, "mid", 0, precision="bf16")
    net2.use_cuda_graph = True
    for _ in range(2):   # first call captures, second replays
        L.load().rgbd_launch_count(1)
        g = net2.compress(rgb, depth)
        rec_g = net2.decompress(g["r_strings"], g["d_strings"], g["shape"])
        assert L.load().rgbd_launch_count(0) > 500     # replays report their kernel nodes
        assert g["r_strings"] == eager["r_strings"] and g["d_strings"] == eager["d_strings"]
        assert torch.equal(rec_g["x_hat"]["r"], rec_e["x_hat"]["r"]) and torch.equal(rec_g["x_hat"]["d"], rec_e["x_hat"]["d"])


def test_container_files_roundtrip_through_the_codec(tmp_path):
    """compress -> the reference's on-disk container (bitstream_io) -> decompress == decompress of the in-memory dict."""
    from gpu_utils import make_model
    from rgbd_b200 import bitstream_io as bio
    net, _ = make_model(rgbd_b200.ELIC_united, "mid", 0, precision="bf16")
    rgb, depth = synthetic_pairs(1, 120, 150, seed=3)
    rgb_p, depth_p = pad_to_multiple(rgb), pad_to_multiple(depth)
    out = net.compress(rgb_p.to(DEV), depth_p.to(DEV))
    rp, dp = str(tmp_path / "rgb" / "x.bin"), str(tmp_path / "depth" / "x.bin")
    bpp_r, bpp_d = bio.save_compressed(out, (120, 150), rp, dp)
    assert bpp_r > 0 and bpp_d > 0
    rs, ds, shape, hw = bio.load_compressed(rp, dp)
    assert hw == (120, 150) and tuple(shape) == tuple(out["shape"])
    a = net.decompress(out["r_strings"], out["d_strings"], out["shape"])
    b = net.decompress(rs, ds, shape)
    assert torch.equal(a["x_hat"]["r"], b["x_hat"]["r"]) and torch.equal(a["x_hat"]["d"], b["x_hat"]["d"])
