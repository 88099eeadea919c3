"""Parameter-holder module tree with the reference's exact state_dict key names.

These nn.Modules only OWN weights (so `load_state_dict`, `.to()`, `.parameters()` behave like
the reference model); none of them has a torch forward.  The compute graph they
describe is compiled by engine.py into launches of the CUDA kernels behind the C-ABI.

Reference layer definitions:
  conv / deconv factories            modules/layers/conv.py:12-24
  ResidualBottleneck                 modules/layers/res_blk.py:7-27
  AttentionBlock (+ ResidualUnit)    CompressAI/compressai/layers/layers.py:162-213
  ESA, SE_Block, bi_spf[_single]     modules/transform/attention.py:14-97
  transforms g_a / g_s / h_a / h_s   modules/transform/analysis.py:63-181,238-249,
                                     modules/transform/synthesis.py:126-242,305-380
  ChannelContextEX                   modules/transform/context.py:10-30
  EntropyParametersEX                modules/transform/entropy.py:56-78
"""
import torch.nn as nn


def conv(cin, cout, k=5, s=2):
    return nn.Conv2d(cin, cout, kernel_size=k, stride=s, padding=k // 2)


def deconv(cin, cout, k=5, s=2):
    return nn.ConvTranspose2d(cin, cout, kernel_size=k, stride=s, output_padding=s - 1, padding=k // 2)


def _chain(*convs):
    """nn.Sequential with placeholders at the activation slots so conv indices are 0, 2, 4."""
    mods = []
    for i, c in enumerate(convs):
        mods.append(c)
        if i + 1 < len(convs):
            mods.append(nn.Identity())
    return nn.Sequential(*mods)


class ResidualBottleneck(nn.Module):
    def __init__(self, N=192, out=None):
        super().__init__()
        out = N if out is None else out
        self.branch = _chain(nn.Conv2d(N, N // 2, 1), nn.Conv2d(N // 2, N // 2, 3, padding=1), nn.Conv2d(N // 2, out, 1))
        self.skip = nn.Conv2d(N, out, 1) if N != out else None


class ResidualUnit(nn.Module):
    def __init__(self, N):
        super().__init__()
        self.conv = _chain(nn.Conv2d(N, N // 2, 1), nn.Conv2d(N // 2, N // 2, 3, padding=1), nn.Conv2d(N // 2, N, 1))


class AttentionBlock(nn.Module):
    def __init__(self, N):
        super().__init__()
        self.conv_a = nn.Sequential(ResidualUnit(N), ResidualUnit(N), ResidualUnit(N))
        self.conv_b = nn.Sequential(ResidualUnit(N), ResidualUnit(N), ResidualUnit(N), nn.Conv2d(N, N, 1))


class ESA(nn.Module):
    def __init__(self, n_feats):
        super().__init__()
        f = n_feats // 4
        self.conv1 = nn.Conv2d(n_feats, f, 1)
        self.conv_f = nn.Conv2d(f, f, 1)
        self.conv_max = nn.Conv2d(f, f, 3, padding=1)
        self.conv2 = nn.Conv2d(f, f, 3, stride=2, padding=0)
        self.conv3 = nn.Conv2d(f, f, 3, padding=1)
        self.conv3_ = nn.Conv2d(f, f, 3, padding=1)
        self.conv4 = nn.Conv2d(f, n_feats, 1)


class BiSpfSingle(nn.Module):
    """bi_spf_single: only the depth branch receives cross-modal features."""

    def __init__(self, N):
        super().__init__()
        self.r_ext = nn.Conv2d(N, N // 2, 3, padding=1)
        self.d_ext = nn.Conv2d(N, N // 2, 3, padding=1)
        self.d_esa = ESA(N)


class BiSpf(BiSpfSingle):
    """bi_spf (Bi-CPT): both branches exchange features."""

    def __init__(self, N):
        super().__init__(N)
        self.r_esa = ESA(N)


class SEBlock(nn.Module):
    def __init__(self, ch, reduction=16):
        super().__init__()
        self.fc = nn.Sequential(nn.Linear(ch, ch // reduction, bias=False), nn.Identity(),
                                nn.Linear(ch // reduction, ch, bias=False), nn.Identity())


class HyperTransformBlock(nn.Module):
    def __init__(self, cin, cout, is_last=False):
        super().__init__()
        self.se = SEBlock(cin)
        self.is_last = is_last
        self.deconv = deconv(cin, cout, k=3, s=1) if is_last else deconv(cin, cout, k=5, s=2)


class EntropyParametersEX(nn.Module):
    def __init__(self, in_dim, out_dim):
        super().__init__()
        self.fusion = _chain(nn.Conv2d(in_dim, in_dim // 6, 1),
                             nn.Conv2d(in_dim // 6, out_dim * 4 // 3, 3, padding=1),
                             nn.Conv2d(out_dim * 4 // 3, out_dim, 5, padding=2))
        self.se = SEBlock(in_dim)


class ChannelContextEX(nn.Module):
    def __init__(self, in_dim, out_dim):
        super().__init__()
        # "fushion" [sic] is the reference's attribute name and therefore the checkpoint key
        self.fushion = _chain(nn.Conv2d(in_dim, 224, 5, padding=2), nn.Conv2d(224, 128, 5, padding=2),
                              nn.Conv2d(128, out_dim, 5, padding=2))


def _rb3(N):
    return [ResidualBottleneck(N), ResidualBottleneck(N), ResidualBottleneck(N)]


class AnalysisTransform(nn.Module):
    """g_a: AnalysisTransformEXcross (cross=True) / AnalysisTransformEXSingle (cross=False)."""

    def __init__(self, N, M, cross=True):
        super().__init__()
        self.cross = cross
        spf = BiSpf if cross else BiSpfSingle
        rin = 2 * N if cross else N  # rgb branch sees the concat only in the bidirectional model
        self.rgb_analysis_transform = nn.Sequential(
            conv(3, N), *_rb3(N), spf(N), conv(rin, N), *_rb3(N), AttentionBlock(N), spf(N),
            conv(rin, N), *_rb3(N), spf(N), conv(rin, M), AttentionBlock(M))
        self.depth_analysis_transform = nn.Sequential(
            conv(1, N), *_rb3(N), nn.Identity(), conv(2 * N, N), *_rb3(N), AttentionBlock(N), nn.Identity(),
            conv(2 * N, N), *_rb3(N), nn.Identity(), conv(2 * N, M), AttentionBlock(M))


class SynthesisTransform(nn.Module):
    """g_s: SynthesisTransformEXcross / SynthesisTransformEXSingle."""

    def __init__(self, N, M, cross=True):
        super().__init__()
        self.cross = cross
        spf = BiSpf if cross else BiSpfSingle
        rin = 2 * N if cross else N

        def stage(cin):
            return [ResidualBottleneck(cin, N), ResidualBottleneck(N), ResidualBottleneck(N)]

        self.rgb_synthesis_transform = nn.Sequential(
            AttentionBlock(M), deconv(M, N), spf(N), *stage(rin), deconv(N, N), AttentionBlock(N), spf(N),
            *stage(rin), deconv(N, N), spf(N), *stage(rin), deconv(N, 3))
        self.depth_synthesis_transform = nn.Sequential(
            AttentionBlock(M), deconv(M, N), nn.Identity(), *stage(2 * N), deconv(N, N), AttentionBlock(N),
            nn.Identity(), *stage(2 * N), deconv(N, N), nn.Identity(), *stage(2 * N), deconv(N, 1))


class HyperAnalysis(nn.Module):
    """h_a: HyperAnalysisEXcross (no cross-modal link)."""

    def __init__(self, N, M):
        super().__init__()
        self.rgb_reduction = _chain(nn.Conv2d(M, N, 3, padding=1), conv(N, N), conv(N, N))
        self.depth_reduction = _chain(nn.Conv2d(M, N, 3, padding=1), conv(N, N), conv(N, N))


class HyperSynthesis(nn.Module):
    """h_s: HyperSynthesisEXcross / HyperSynthesisEXSingle."""

    def __init__(self, N, M, cross=True):
        super().__init__()
        self.cross = cross
        k = 2 if cross else 1
        self.r_h_s1 = HyperTransformBlock(k * N, M)
        self.r_h_s2 = HyperTransformBlock(k * M, M * 3 // 2)
        self.r_h_s3 = HyperTransformBlock(k * M * 3 // 2, 2 * M, True)
        self.d_h_s1 = HyperTransformBlock(2 * N, M)
        self.d_h_s2 = HyperTransformBlock(2 * M, M * 3 // 2)
        self.d_h_s3 = HyperTransformBlock(M * 3, 2 * M, True)
