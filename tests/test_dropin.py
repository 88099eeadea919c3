"""The drop-in launcher against the REAL reference, where it is mounted (the build container; skipped on the GPU box):
`dropin.install` re-assigns the two entries of `models.modelZoo` in place, the tester's substring lookup
(testing/tester.py:55-59) then constructs OUR classes from the reference's own `model_config()`, and they accept a
state_dict with the reference's keys."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("RGBD_REFERENCE_ROOT", "/root/reference")

SCRIPT = r'''
import json, os, sys
root, ref = sys.argv[1], sys.argv[2]
sys.path.insert(0, root)
from oracle import ref_loader
ref_loader.load_ref_ext("ans"); ref_loader.load_ref_ext("_CXX")          # the reference's compiled coder modules
sys.path.insert(0, os.path.join(root, "oracle", "shims"))               # timm / pytorch_msssim stand-ins
import rgbd_b200
from rgbd_b200 import dropin
zoo = dropin.install(ref)
import models                                                            # the reference's package
from testing import tester as ref_tester                                 # binds `from models import modelZoo`
assert ref_tester.modelZoo is models.modelZoo is zoo
keys = list(zoo)
from config.config import model_config                                   # the reference's config
picked = {}
for model_name in ("ELIC_united", "ELIC_united_R2D", "my_ELIC_united_R2D_run"):
    for name, cls in zoo.items():                                        # testing/tester.py:55-59
        if model_name.find(name) != -1:
            net = cls(config=model_config(), channel=4).eval()
            picked[model_name] = [type(net).__module__, type(net).__name__, len(net.state_dict())]
            break
# the reference's own Tester.get_net + restore (testing/tester.py:55-66,100-108) on our class: construct by substring lookup,
# torch.load, load_state_dict, update(force=True), .to(device).  __init__ is bypassed only because it hard-codes
# device = "cuda" and opens a dataset; get_net / restore run unmodified (this container has no GPU: device = "cpu").
import tempfile, torch
from testing.tester_united import TesterUnited
from rgbd_b200.synthetic import synthetic_state_dict
harness = {}
with tempfile.TemporaryDirectory() as tmp:
    seed = rgbd_b200.ELIC_united(config=model_config(), channel=4).eval()
    seed.load_state_dict(synthetic_state_dict(seed, 0, "mid"))
    seed.update(force=True)
    ckpt = os.path.join(tmp, "checkpoint_best_loss.pth.tar")
    torch.save({"state_dict": seed.state_dict(), "epoch": 7}, ckpt)
    t = object.__new__(TesterUnited)
    t.device, t.channel, t.ckpt_dir_path = "cpu", 4, tmp
    epoch = t.get_net(model_config=model_config(), model_name="ELIC_united", ckpt_path=None)   # finds the file in ckpt_dir_path
    same = all(torch.equal(a, b) for a, b in zip(t.net.state_dict().values(), seed.state_dict().values()))
    harness = {"epoch": int(epoch), "cls": [type(t.net).__module__, type(t.net).__name__], "state_equal": bool(same),
               "cdf_rows": int(t.net.rgb_gaussian_conditional.quantized_cdf.shape[0]), "training": bool(t.net.training)}
print(json.dumps({"keys": keys, "picked": picked, "harness": harness,
                  "ours": [rgbd_b200.ELIC_united.__module__, rgbd_b200.ELIC_united_R2D.__module__]}))
'''


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")), reason="reference checkout not mounted here")
def test_dropin_patches_the_reference_model_zoo(golden_dir):
    from oracle import ref_loader
    if not ref_loader.reference_available():
        pytest.skip("oracle/_ref not built")
    out = subprocess.run([sys.executable, "-c", SCRIPT, ROOT, REF], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    info = json.loads(out.stdout.strip().splitlines()[-1])
    keys = info["keys"]
    assert keys.index("ELIC_united_R2D") < keys.index("ELIC_united")      # substring lookup order survives the patch
    want_keys = {"ELIC_united": len(json.load(open(os.path.join(golden_dir, "state_dict_keys_united.json")))),
                 "ELIC_united_R2D": len(json.load(open(os.path.join(golden_dir, "state_dict_keys_r2d.json"))))}
    for model_name, cls_name in (("ELIC_united", "ELIC_united"), ("ELIC_united_R2D", "ELIC_united_R2D"),
                                 ("my_ELIC_united_R2D_run", "ELIC_united_R2D")):
        mod, name, nkeys = info["picked"][model_name]
        assert name == cls_name and mod in info["ours"], (model_name, mod, name)   # OUR class, not the reference's
        assert nkeys == want_keys[cls_name]
    h = info["harness"]      # the reference's Tester.get_net / restore ran on our class
    assert h["epoch"] == 7 and h["cls"][1] == "ELIC_united" and h["cls"][0] in info["ours"]
    assert h["state_equal"] and h["cdf_rows"] == 64 and not h["training"]
