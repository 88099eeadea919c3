"""Pins the torch-CPU model restatement (oracle/model_oracle.py) against tensors and bytes the
UNMODIFIED reference produced for the same seeded weights and inputs (tests/golden/model_*.npz)."""
import json

import numpy as np
import pytest
import torch

import rgbd_b200
from oracle.model_oracle import OracleCodec
from rgbd_b200.synthetic import synthetic_pairs, synthetic_state_dict


def _setup(golden_dir, name):
    g = np.load(f"{golden_dir}/model_{name}.npz")
    meta = json.loads(str(g["meta"]))
    cls = rgbd_b200.ELIC_united if meta["cross"] else rgbd_b200.ELIC_united_R2D
    net = cls(config=rgbd_b200.model_config(), channel=4).eval()
    net.load_state_dict(synthetic_state_dict(net, meta["seed"], meta["preset"]))
    net.update(force=True)   # host-side table build (no GPU involved)
    rgb, depth = synthetic_pairs(1, meta["H"], meta["W"], seed=meta["input_seed"])
    return g, meta, OracleCodec(net.state_dict(), cross=meta["cross"]), rgb, depth


@pytest.mark.parametrize("name", ["united", "r2d"])
def test_oracle_reproduces_reference(golden_dir, name):
    g, meta, orc, rgb, depth = _setup(golden_dir, name)
    c = orc.compress(rgb, depth, trace=True)
    tr = c["_trace"]
    # float tensors: same torch CPU kernels, same op order -> bit-equal
    for k in ("y_r", "y_d", "z_r", "z_d"):
        assert np.array_equal(tr[k].numpy(), g[k]), k
    # bitstreams: byte-identical to the reference's
    assert c["r_strings"][0][0] == g["r_y"].tobytes() and c["r_strings"][1][0] == g["r_z"].tobytes()
    assert c["d_strings"][0][0] == g["d_y"].tobytes() and c["d_strings"][1][0] == g["d_z"].tobytes()
    assert tuple(c["shape"]) == tuple(g["shape"])
    d = orc.decompress(c["r_strings"], c["d_strings"], c["shape"])
    assert np.array_equal(d["x_hat"]["r"].numpy(), g["xhat_r"])
    assert np.array_equal(d["x_hat"]["d"].numpy(), g["xhat_d"])
    # every symbol round-trips: decoder's y_hat == encoder's y_hat
    assert torch.equal(d["_trace"]["yhat_r"], tr["yhat_r"]) and torch.equal(d["_trace"]["yhat_d"], tr["yhat_d"])


@pytest.mark.parametrize("name", ["united", "r2d"])
def test_oracle_forward_reproduces_reference(golden_dir, name):
    g, meta, orc, rgb, depth = _setup(golden_dir, name)
    f = orc.forward(rgb, depth)
    assert np.array_equal(f["x_hat"]["r"].numpy(), g["fwd_xhat_r"])
    assert np.array_equal(f["x_hat"]["d"].numpy(), g["fwd_xhat_d"])
    for k, t in (("lik_y_r", f["r_likelihoods"]["y"]), ("lik_y_d", f["d_likelihoods"]["y"]),
                 ("lik_z_r", f["r_likelihoods"]["z"]), ("lik_z_d", f["d_likelihoods"]["z"])):
        assert np.array_equal(t.numpy(), g[k]), k
        assert t.min() >= torch.tensor(1e-9)  # LowerBound(1e-9) in fp32


def test_state_dict_keys_match_reference(golden_dir):
    for name, cls in (("united", rgbd_b200.ELIC_united), ("r2d", rgbd_b200.ELIC_united_R2D)):
        want = json.load(open(f"{golden_dir}/state_dict_keys_{name}.json"))
        net = cls(config=rgbd_b200.model_config(), channel=4)
        net.update(force=True)                           # the golden shapes were taken after update()
        got = {k: list(v.shape) for k, v in net.state_dict().items()}
        assert list(got) == list(want), name           # same keys, same order
        assert got == want, name                         # same shapes, CDF tables included
