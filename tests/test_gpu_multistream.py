"""Opt-in multi-stream bitstream layout: the y symbols of an image cut into equal sub-streams, each a
complete rANS string that the reference coder would produce from (and decode back to) exactly those symbols, carried as
further strings of the container's y entry.  The default layout stays the reference's single stream."""
import numpy as np
import pytest
import torch

import rgbd_b200
from oracle import coder
from oracle.model_oracle import OracleCodec
from rgbd_b200.synthetic import synthetic_pairs

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("cls_name,precision,sub", [("ELIC_united", "fp32", 4), ("ELIC_united", "bf16", 16),
                                                    ("ELIC_united_R2D", "bf16", 4)])
def test_multistream_substreams_equal_reference_coder_and_roundtrip(cls_name, precision, sub, tmp_path):
    from gpu_utils import make_model
    from rgbd_b200 import bitstream_io as bio
    cls = getattr(rgbd_b200, cls_name)
    net, sd = make_model(cls, "mid", 0, precision=precision, stream_layout="multi", sub_channels=sub)
    ref_net, _ = make_model(cls, "mid", 0, precision=precision)          # same weights, reference layout
    orc = OracleCodec(sd, cross=cls_name == "ELIC_united")
    B, H, W = 2, 128, 192
    rgb, depth = synthetic_pairs(B, H, W, seed=41)
    out = net.compress(rgb.to(DEV), depth.to(DEV))
    single = ref_net.compress(rgb.to(DEV), depth.to(DEV))
    h, w = H // 16, W // 16
    nsub, sublen = 2 * 320 // sub, sub * h * (w // 2)
    prog = net._program("encoder", B, H, W)
    assert prog.io["nsub"] == nsub and prog.io["sublen"] == sublen
    for which, key, name in (("r", "r_strings", "rgb"), ("d", "d_strings", "depth")):
        assert len(out[key][0]) == B * nsub and len(out[key][1]) == B
        assert out[key][1] == single[key][1]                            # z strings do not change
        st = prog.io["st"][which]
        ysym = st["ysym"].cpu().numpy()
        yidx = st["yidx"].cpu().numpy().astype(np.int32)
        # same symbols as the single-stream layout of the same network
        assert np.array_equal(ysym, ref_net._program("encoder", B, H, W).io["st"][which]["ysym"].cpu().numpy())
        t = orc.gc_tables(name)
        for i in range(B):
            for k in range(nsub):
                s_, i_ = ysym[i, k * sublen:(k + 1) * sublen], yidx[i, k * sublen:(k + 1) * sublen]
                sub_stream = out[key][0][i * nsub + k]
                # KAT: byte-identical to RansEncoder.encode_with_indexes on this sub-stream's symbols ...
                assert sub_stream == coder.encode_with_indexes(s_, i_, t), (which, i, k)
                if k % 37 == 0:   # ... and the reference decoder reads them back
                    assert np.array_equal(coder.decode_with_indexes(sub_stream, i_, t), s_), (which, i, k)
    rec = net.decompress(out["r_strings"], out["d_strings"], out["shape"])
    ref = ref_net.decompress(single["r_strings"], single["d_strings"], single["shape"])
    for m in ("r", "d"):
        assert torch.equal(rec["x_hat"][m], ref["x_hat"][m]), m           # same symbols -> same reconstruction, bit for bit
    dec = net._program("decoder", B, H // 64, W // 64, sub)
    for k in ("r", "d"):
        assert torch.equal(dec.io["st"][k]["ysym"], prog.io["st"][k]["ysym"]), k
    # a network configured for the reference layout decodes multi-stream input too (the layout is read off the strings)
    rec2 = ref_net.decompress(out["r_strings"], out["d_strings"], out["shape"])
    assert torch.equal(rec2["x_hat"]["r"], ref["x_hat"]["r"])
    # the existing container carries it (utils/IOutils.py:58-88 writes n strings per entry)
    one = net.compress(rgb[:1].to(DEV), depth[:1].to(DEV))
    rp, dp = str(tmp_path / "rgb" / "a"), str(tmp_path / "depth" / "a")
    bio.save_compressed(one, (H, W), rp, dp)
    rs, ds, shape, hw = bio.load_compressed(rp, dp)
    rec3 = net.decompress(rs, ds, shape)
    assert torch.equal(rec3["x_hat"]["d"][0], ref["x_hat"]["d"][0])
    # corrupt sub-stream: flip one byte in the middle of a long one -> the consumed-exactly check (or the symbols) notice
    bad = [list(g) for g in out["r_strings"]]
    k_long = max(range(len(bad[0])), key=lambda j: len(bad[0][j]))
    b_ = bytearray(bad[0][k_long])
    b_[len(b_) // 2] ^= 0x5A
    bad[0][k_long] = bytes(b_)
    try:
        r_bad = net.decompress(bad, out["d_strings"], out["shape"])
        assert not torch.equal(r_bad["x_hat"]["r"], rec["x_hat"]["r"])
    except ValueError as e:
        assert "corrupt stream" in str(e)


def test_truncated_stream_is_reported():
    from gpu_utils import make_model
    net, _ = make_model(rgbd_b200.ELIC_united, "mid", 0, precision="bf16")
    rgb, depth = synthetic_pairs(1, 128, 128, seed=2)
    out = net.compress(rgb.to(DEV), depth.to(DEV))
    y = out["r_strings"][0][0]
    cut = [[y[:len(y) // 2 // 4 * 4]], out["r_strings"][1]]
    with pytest.raises(ValueError, match="corrupt stream"):
        net.decompress(cut, out["d_strings"], out["shape"])
    # the intact strings still decode afterwards (the failed call left no state behind)
    net.decompress(out["r_strings"], out["d_strings"], out["shape"])
