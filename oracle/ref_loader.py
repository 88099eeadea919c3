"""ORACLE (test infrastructure): import the UNMODIFIED reference in the build container.

Only works where /root/reference is mounted (never on the GPU box).  Used by
oracle/make_golden.py to generate tests/golden/* and by the `not gpu` tests that
cross-check the oracle restatement live when the reference is available.

The reference's two pybind11 modules are compiled by oracle/Makefile into oracle/_ref/
from the sources where they lie; they are pre-registered in sys.modules under their
dotted names so that the reference's pure-python `compressai` package (imported straight
from the read-only mount) finds them.
"""
import glob
import importlib.machinery
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("RGBD_REFERENCE_ROOT", "/root/reference")


def ref_ext_available():
    return bool(glob.glob(os.path.join(HERE, "_ref", "ans*.so")))


def load_ref_ext(name):
    """Load oracle/_ref/<name>*.so as module compressai.<name> (no reference python needed)."""
    full = "compressai." + name
    if full in sys.modules:
        return sys.modules[full]
    hits = glob.glob(os.path.join(HERE, "_ref", name + "*.so"))
    if not hits:
        raise ImportError(f"oracle/_ref/{name}*.so not built (run `make -C oracle ref` in the build container)")
    loader = importlib.machinery.ExtensionFileLoader(full, hits[0])
    spec = importlib.util.spec_from_file_location(full, hits[0], loader=loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    sys.modules[full] = mod
    return mod


def reference_available():
    return os.path.isdir(os.path.join(REF_ROOT, "CompressAI", "compressai")) and ref_ext_available()


def import_reference():
    """Returns (ELIC_united, ELIC_united_R2D, model_config) from the read-only reference."""
    if not reference_available():
        raise ImportError("reference not available here")
    import torch

    load_ref_ext("ans")
    load_ref_ext("_CXX")
    for p in (os.path.join(HERE, "shims"), os.path.join(REF_ROOT, "CompressAI"), REF_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    if not torch.cuda.is_available():
        torch.cuda.synchronize = lambda *a, **k: None  # elic_united.py:431,449 on a CUDA-less host
    from config.config import model_config
    from models.elic_united import ELIC_united
    from models.elic_united_R2D import ELIC_united_R2D

    return ELIC_united, ELIC_united_R2D, model_config
