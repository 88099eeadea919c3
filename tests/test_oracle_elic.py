"""Pins oracle/elic_oracle.py (single-modality ELIC) against tensors and bytes the UNMODIFIED reference
produced (tests/golden/model_elic_{rgb,depth}.npz, oracle/make_golden_elic.py), and the module tree of
rgbd_b200.ELIC against the reference's state_dict keys."""
import json

import numpy as np
import pytest
import torch

import rgbd_b200
from oracle.elic_oracle import ElicOracle
from rgbd_b200.synthetic import synthetic_pairs, synthetic_state_dict


def _setup(golden_dir, name):
    g = np.load(f"{golden_dir}/model_elic_{name}.npz")
    meta = json.loads(str(g["meta"]))
    net = rgbd_b200.ELIC(config=rgbd_b200.model_config(), channel=meta["channel"]).eval()
    net.load_state_dict(synthetic_state_dict(net, meta["seed"], meta["preset"]))
    net.update(force=True)
    rgb, depth = synthetic_pairs(1, meta["H"], meta["W"], seed=meta["input_seed"])
    return g, meta, ElicOracle(net.state_dict()), (rgb if meta["channel"] == 3 else depth)


@pytest.mark.parametrize("name", ["rgb", "depth"])
def test_elic_oracle_reproduces_reference(golden_dir, name):
    g, meta, orc, x = _setup(golden_dir, name)
    c = orc.compress(x, trace=True)
    tr = c["_trace"]
    assert np.array_equal(tr["y"].numpy(), g["y"]) and np.array_equal(tr["z"].numpy(), g["z"])
    assert c["strings"][0][0] == g["y_bytes"].tobytes() and c["strings"][1][0] == g["z_bytes"].tobytes()
    assert tuple(c["shape"]) == tuple(g["shape"])
    d = orc.decompress(c["strings"], c["shape"])
    assert np.array_equal(d["x_hat"].numpy(), g["xhat"])
    assert torch.equal(d["_trace"]["yhat"], tr["yhat"])
    f = orc.forward(x)
    assert np.array_equal(f["x_hat"].numpy(), g["fwd_xhat"])
    assert np.array_equal(f["likelihoods"]["y_likelihoods"].numpy(), g["lik_y"])
    assert np.array_equal(f["likelihoods"]["z_likelihoods"].numpy(), g["lik_z"])


def test_elic_state_dict_keys_match_reference(golden_dir):
    for name, ch in (("rgb", 3), ("depth", 1)):
        want = json.load(open(f"{golden_dir}/state_dict_keys_elic_{name}.json"))
        net = rgbd_b200.ELIC(config=rgbd_b200.model_config(), channel=ch)
        net.update(force=True)
        got = {k: list(v.shape) for k, v in net.state_dict().items()}
        assert list(got) == list(want), name
        assert got == want, name
    assert list(rgbd_b200.modelZoo) == ["ELIC_united_R2D", "ELIC_united", "ELIC", "STF_united"]   # substring lookup order (models/__init__.py:11-20)
