"""ORACLE (test infrastructure): golden vectors of SymmetricalTransFormerUnited (models/stf_united.py) from the UNMODIFIED
reference.  Build container only:  python -m oracle.make_golden_stf"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")
H, W = 256, 256          # the ESA at 1/16 scale needs >= 15 pixels a side (conv 3x3 s2 + max_pool2d(7, 3))


def main():
    from oracle.ref_loader import import_reference
    import_reference()                       # sys.path + compiled extensions + shims
    from config.config import model_config
    from models.stf_united import SymmetricalTransFormerUnited
    from rgbd_b200.synthetic import synthetic_pairs, synthetic_state_dict
    torch.manual_seed(0)
    net = SymmetricalTransFormerUnited(config=model_config(), channel=4).eval()
    net.load_state_dict(synthetic_state_dict(net, 0, "mid"))
    net.update(force=True)
    rgb, depth = synthetic_pairs(1, H, W, seed=4321)
    with torch.no_grad():
        yr, yd = net.g_a(rgb, depth)
        c = net.compress(rgb, depth)
        d = net.decompress(c["r_strings"], c["d_strings"], c["shape"])
        f = net(rgb, depth)
    out = {"y_r": yr.numpy(), "y_d": yd.numpy(), "xhat_r": d["x_hat"]["r"].numpy(), "xhat_d": d["x_hat"]["d"].numpy(),
           "fwd_xhat_r": f["x_hat"]["r"].numpy(), "fwd_xhat_d": f["x_hat"]["d"].numpy(),
           "lik_y_r": f["r_likelihoods"]["y"].numpy(), "lik_y_d": f["d_likelihoods"]["y"].numpy(),
           "shape": np.array(list(c["shape"])),
           "ry_bytes": np.frombuffer(c["r_strings"][0][0], dtype=np.uint8), "rz_bytes": np.frombuffer(c["r_strings"][1][0], dtype=np.uint8),
           "dy_bytes": np.frombuffer(c["d_strings"][0][0], dtype=np.uint8), "dz_bytes": np.frombuffer(c["d_strings"][1][0], dtype=np.uint8),
           "meta": np.array(json.dumps(dict(H=H, W=W, preset="mid", seed=0, input_seed=4321)))}
    np.savez_compressed(os.path.join(GOLD, "model_stf_united.npz"), **out)
    keys = {k: list(v.shape) for k, v in net.state_dict().items()}
    with open(os.path.join(GOLD, "state_dict_keys_stf_united.json"), "w") as fh:
        json.dump(keys, fh, indent=0)
    print("y std", float(yr.std()), float(yd.std()), "bytes", {k: len(out[k]) for k in out if k.endswith("bytes")}, "keys", len(keys),
          "xhat range", float(out["xhat_r"].min()), float(out["xhat_r"].max()))


if __name__ == "__main__":
    main()
