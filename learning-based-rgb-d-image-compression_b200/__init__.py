"""B200-native ELIC_united compress/decompress path.

The directory name follows the project naming rule and is not a valid Python identifier;
import it through the `rgbd_b200` alias module at the repository root
(`import rgbd_b200`), or `importlib.import_module("learning-based-rgb-d-image-compression_b200")`.
"""
from .elic_united import ELIC_united  # noqa: F401
from .elic_united_r2d import ELIC_united_R2D  # noqa: F401
from .elic import ELIC  # noqa: F401
from .stf_united import STF_united, SymmetricalTransFormerUnited  # noqa: F401
from .config import Config, model_config  # noqa: F401
from . import bitstream_io, lib, metrics, synthetic  # noqa: F401
from .metrics import AverageMeter, compute_metrics  # noqa: F401

# lookup is by substring in dict order, so the R2D key must come first (models/__init__.py:11-20)
modelZoo = {"ELIC_united_R2D": ELIC_united_R2D, "ELIC_united": ELIC_united, "ELIC": ELIC, "STF_united": SymmetricalTransFormerUnited}
