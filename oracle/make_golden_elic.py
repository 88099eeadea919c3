"""ORACLE (test infrastructure): golden vectors of the single-modality ELIC (models/elic.py) from the UNMODIFIED
reference, for channel = 3 (rgb) and channel = 1 (depth).  Build container only:  python -m oracle.make_golden_elic"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    from oracle.ref_loader import import_reference
    import_reference()                       # sys.path + compiled extensions + shims
    from config.config import model_config
    from models.elic import ELIC
    from rgbd_b200.synthetic import synthetic_pairs, synthetic_state_dict
    for channel, name in ((3, "rgb"), (1, "depth")):
        torch.manual_seed(0)
        net = ELIC(config=model_config(), channel=channel).eval()
        net.load_state_dict(synthetic_state_dict(net, 0, "mid"))
        net.update(force=True)
        rgb, depth = synthetic_pairs(1, 128, 192, seed=4321)
        x = rgb if channel == 3 else depth
        with torch.no_grad():
            y = net.g_a(x)
            z = net.h_a(y)
            c = net.compress(x)
            d = net.decompress(c["strings"], c["shape"])
            f = net(x)
        out = {"y": y.numpy(), "z": z.numpy(), "xhat": d["x_hat"].numpy(), "fwd_xhat": f["x_hat"].numpy(),
               "lik_y": f["likelihoods"]["y_likelihoods"].numpy(), "lik_z": f["likelihoods"]["z_likelihoods"].numpy(),
               "shape": np.array(list(c["shape"])),
               "y_bytes": np.frombuffer(c["strings"][0][0], dtype=np.uint8),
               "z_bytes": np.frombuffer(c["strings"][1][0], dtype=np.uint8),
               "meta": np.array(json.dumps(dict(H=128, W=192, preset="mid", seed=0, input_seed=4321, channel=channel)))}
        np.savez_compressed(os.path.join(GOLD, f"model_elic_{name}.npz"), **out)
        keys = {k: list(v.shape) for k, v in net.state_dict().items()}
        with open(os.path.join(GOLD, f"state_dict_keys_elic_{name}.json"), "w") as fh:
            json.dump(keys, fh, indent=0)
        print(name, "y bytes", len(out["y_bytes"]), "z bytes", len(out["z_bytes"]), "keys", len(keys))


if __name__ == "__main__":
    main()
