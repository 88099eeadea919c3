"""oracle/stf_oracle.py (SymmetricalTransFormerUnited restated over a state_dict) against the golden produced by the
unmodified reference (oracle/make_golden_stf.py): latents, bytes, reconstructions and likelihoods."""
import json

import numpy as np
import torch

import rgbd_b200
from oracle.stf_oracle import StfOracle
from oracle.tables import updated_state_dict
from rgbd_b200.synthetic import synthetic_pairs, synthetic_state_dict


def _setup(golden_dir):
    g = np.load(f"{golden_dir}/model_stf_united.npz")
    meta = json.loads(str(g["meta"]))
    net = rgbd_b200.STF_united(config=rgbd_b200.model_config(), channel=4).eval()
    sd = updated_state_dict(synthetic_state_dict(net, meta["seed"], meta["preset"]))
    rgb, depth = synthetic_pairs(1, meta["H"], meta["W"], seed=meta["input_seed"])
    return g, net, sd, rgb, depth


def test_state_dict_keys_equal_reference(golden_dir):
    net = rgbd_b200.STF_united(config=rgbd_b200.model_config(), channel=4)
    net.update(force=True)                                   # the golden shapes were taken after update()
    want = json.load(open(f"{golden_dir}/state_dict_keys_stf_united.json"))
    got = {k: list(v.shape) for k, v in net.state_dict().items()}
    assert list(got) == list(want) and got == want          # same keys, same order, same shapes (1244)


def test_stf_oracle_reproduces_reference_golden(golden_dir):
    g, net, sd, rgb, depth = _setup(golden_dir)
    orc = StfOracle(sd)
    with torch.no_grad():
        yr, yd = orc.g_a(rgb, depth)
        assert np.abs(yr.numpy() - g["y_r"]).max() < 2e-5 and np.abs(yd.numpy() - g["y_d"]).max() < 2e-5
        c = orc.compress(rgb, depth)
        assert tuple(c["shape"]) == tuple(g["shape"])
        for key, name in (("r_strings", "r"), ("d_strings", "d")):
            assert c[key][1][0] == g[name + "z_bytes"].tobytes()
            want = len(g[name + "y_bytes"])
            assert abs(len(c[key][0][0]) - want) <= 0.002 * want          # (bit-equal where the latents are)
        d = orc.decompress(c["r_strings"], c["d_strings"], c["shape"])
        for m in ("r", "d"):
            assert float(((d["x_hat"][m] - torch.from_numpy(g["xhat_" + m])) ** 2).mean()) < 1e-6
        f = orc.forward(rgb, depth)
        for m in ("r", "d"):
            assert float(((f["x_hat"][m] - torch.from_numpy(g["fwd_xhat_" + m])) ** 2).mean()) < 1e-6
            got, want = f[m + "_likelihoods"]["y"], g["lik_y_" + m]
            bits_g, bits_w = float(-torch.log2(got).sum()), float(-np.log2(want).sum())
            assert abs(bits_g - bits_w) / bits_w < 1e-4
