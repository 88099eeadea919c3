"""Parameter holders of SymmetricalTransFormerUnited's transforms (models/stf_united.py:15-614).

Only constructor shapes and attribute names matter here — they must equal the reference's so that its checkpoints load
(`state_dict` keys in the same order); the forward passes live in stf_united.py as launch plans over our kernels.
"""
import torch
import torch.nn as nn

from .modules import BiSpf


class Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class WindowAttention(nn.Module):
    def __init__(self, dim, window_size, num_heads):
        super().__init__()
        self.dim, self.window_size, self.num_heads = dim, window_size, num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * window_size - 1) ** 2, num_heads))
        # pair-wise relative position index of the tokens of a window (stf_united.py:62-73)
        coords = torch.stack(torch.meshgrid([torch.arange(window_size), torch.arange(window_size)], indexing="ij")).flatten(1)
        rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0).contiguous()
        rel[:, :, 0] += window_size - 1
        rel[:, :, 1] += window_size - 1
        rel[:, :, 0] *= 2 * window_size - 1
        self.register_buffer("relative_position_index", rel.sum(-1))
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)


class SwinTransformerBlock(nn.Module):
    def __init__(self, dim, num_heads, window_size, shift_size, mlp_ratio=4.0):
        super().__init__()
        self.dim, self.num_heads, self.window_size, self.shift_size = dim, num_heads, window_size, shift_size
        self.norm1 = nn.LayerNorm(dim)
        self.attn = WindowAttention(dim, window_size, num_heads)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))


class PatchMerging(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.reduction = nn.Linear(4 * dim, 2 * dim, bias=False)
        self.norm = nn.LayerNorm(4 * dim)


class PatchSplit(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.reduction = nn.Linear(dim, dim * 2, bias=False)
        self.norm = nn.LayerNorm(dim)


class BasicLayer(nn.Module):
    def __init__(self, dim, depth, num_heads, window_size, downsample=None):
        super().__init__()
        self.window_size = window_size
        self.blocks = nn.ModuleList(SwinTransformerBlock(dim, num_heads, window_size, 0 if i % 2 == 0 else window_size // 2)
                                    for i in range(depth))
        self.downsample = downsample(dim) if downsample is not None else None


class PatchEmbed(nn.Module):
    def __init__(self, patch_size, in_chans, embed_dim):
        super().__init__()
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)
        self.norm = nn.LayerNorm(embed_dim)


class AnalysisTransformSTF(nn.Module):
    """AnalysisTransformSTFunited (stf_united.py:408-511)."""

    def __init__(self, patch_size=2, embed_dim=48, depths=(2, 2, 6, 2), num_heads=(3, 6, 12, 24), window_size=4):
        super().__init__()
        self.embed_dim = embed_dim
        self.rgb_patch_embed = PatchEmbed(patch_size, 3, embed_dim)
        self.depth_patch_embed = PatchEmbed(patch_size, 1, embed_dim)
        self.rgb_ana_layers, self.depth_ana_layers = nn.ModuleList(), nn.ModuleList()
        dim = embed_dim
        n = len(depths)
        for i in range(n):
            for lst in (self.rgb_ana_layers, self.depth_ana_layers):
                lst.append(BasicLayer(dim, depths[i], num_heads[i], window_size, PatchMerging if i < n - 1 else None))
            dim *= 2
            if i < n - 1:
                self.rgb_ana_layers.append(BiSpf(dim))
                self.depth_ana_layers.append(nn.Identity())


class SynthesisTransformSTF(nn.Module):
    """SynthesisTransformSTFunited (stf_united.py:514-613)."""

    def __init__(self, patch_size=2, embed_dim=48, depths=(2, 2, 6, 2), num_heads=(3, 6, 12, 24), window_size=4):
        super().__init__()
        self.embed_dim = embed_dim
        depths, num_heads = depths[::-1], num_heads[::-1]
        self.rgb_syn_layers, self.depth_syn_layers = nn.ModuleList(), nn.ModuleList()
        dim = embed_dim * 8
        n = len(depths)
        for i in range(n):
            for lst in (self.rgb_syn_layers, self.depth_syn_layers):
                lst.append(BasicLayer(dim, depths[i], num_heads[i], window_size, PatchSplit if i < n - 1 else None))
            dim //= 2
            if i < n - 1:
                self.rgb_syn_layers.append(BiSpf(dim))
                self.depth_syn_layers.append(nn.Identity())

        def end(cout):
            return nn.Sequential(nn.Conv2d(embed_dim, embed_dim * patch_size ** 2, 5, 1, 2), nn.PixelShuffle(patch_size),
                                 nn.Conv2d(embed_dim, cout, 3, 1, 1))

        self.rgb_end_conv = end(3)
        self.depth_end_conv = end(1)
