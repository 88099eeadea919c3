"""TEST INFRASTRUCTURE — torch-CPU emulation of bf16 tensor-core arithmetic on top of the fp32 oracle.

Every Conv2d / ConvTranspose2d of the oracle (model_oracle.OracleCodec._c / _d, which restate the reference's layers:
modules/transform/*.py, modules/layers/*.py) runs with its input and weights rounded to bf16, fp32 accumulation, and its
output rounded to bf16 — what a tensor core with bf16 operands and an fp32 accumulator computes, with nothing else
changed.  It answers one question for the parity tests: how far from the fp32 oracle does bf16 arithmetic ITSELF put a
reconstruction?  The CUDA path in bf16 mode is then held to that figure (within a margin) instead of to an arbitrary
bar: on the synthetic stand-in weights the emulation sits 40 dB (`mid`) to 45 dB (`realistic`) from the fp32 oracle
(PSNR with peak = 4 sigma, see tests/gpu_utils.recon_fidelity_db).  Never imported by the product package.
"""
import torch.nn.functional as F

from .elic_oracle import ElicOracle
from .model_oracle import OracleCodec
from .stf_oracle import StfOracle


def _bf(t):
    return t.bfloat16().float()


class _Bf16Convs:
    def _c(self, name, x, stride=1, pad=0):
        return _bf(F.conv2d(_bf(x), _bf(self.sd[name + ".weight"]), self.sd[name + ".bias"], stride=stride, padding=pad))

    def _d(self, name, x, k=5, s=2):
        return _bf(F.conv_transpose2d(_bf(x), _bf(self.sd[name + ".weight"]), self.sd[name + ".bias"], stride=s,
                                      padding=k // 2, output_padding=s - 1))


class Bf16OracleCodec(_Bf16Convs, OracleCodec):
    pass


class Bf16ElicOracle(_Bf16Convs, ElicOracle):
    pass


class Bf16StfOracle(_Bf16Convs, StfOracle):
    """+ every nn.Linear of the Swin blocks as a bf16 GEMM, and the tensors a bf16 path stores between launches
    (LayerNorm outputs, the residual stream) rounded to bf16."""

    def _lin(self, p, x):
        w, bias = self.sd[p + ".weight"], self.sd.get(p + ".bias")
        return _bf(F.linear(_bf(x), _bf(w), bias))

    def _ln(self, p, x):
        return _bf(super()._ln(p, x))

    @staticmethod
    def _add(a, b):
        return _bf(a + b)          # the residual stream lives in bf16 between launches
