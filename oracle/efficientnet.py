"""CPU restatement of the 2-D patch encoder (scope row N4): ``timm.create_model("efficientnet_b0",
features_only=True)`` as the reference instantiates it (``puzzle_diff/model/backbones/efficient_gat.py:40-42``)
and consumes it (``visual_features``, ``:149-189``: feature maps 2 and 3 -- 40 channels at stride 8, 112 channels
at stride 16 -- flattened channel-major and concatenated: 40*4*4 + 112*2*2 = 1088 values per 32x32 patch).

TEST INFRASTRUCTURE.  ``timm`` is an un-vendored, unpinned third-party dependency (``singularity/build/
conda_env.yaml``) and is absent from this image, so the architecture is restated from its published definition
(EfficientNet-B0, Tan & Le 2019: stem 3x3/2 -> 32; MBConv1 k3 16; MBConv6 k3 24 x2 /2; MBConv6 k5 40 x2 /2;
MBConv6 k3 80 x3 /2; MBConv6 k5 112 x3; MBConv6 k5 192 x4 /2; MBConv6 k3 320; squeeze-excite with a quarter of
the block's INPUT channels, SiLU, BatchNorm eps 1e-5, symmetric "same" padding k // 2) with timm's parameter
names (``conv_stem``, ``bn1``, ``blocks.S.B.{conv_pw, bn1, conv_dw, bn2, se.conv_reduce, se.conv_expand,
conv_pwl, bn3}``; the depthwise-separable first stage has ``conv_dw, bn1, se, conv_pw, bn2``).  **Parity of this
file is pinned against torchvision's ``efficientnet_b0``** -- an independent implementation of the same published
architecture that IS installed here -- by ``tests/test_oracle_efficientnet.py`` (weights copied across by a key
map, feature taps equal to 1e-5); against timm itself it stays unpinned.
"""
import torch
from torch import nn

# (expand ratio, kernel, stride, out channels, repeats) per stage -- EfficientNet-B0
STAGES = [(1, 3, 1, 16, 1), (6, 3, 2, 24, 2), (6, 5, 2, 40, 2), (6, 3, 2, 80, 3), (6, 5, 1, 112, 3), (6, 5, 2, 192, 4), (6, 3, 1, 320, 1)]
FEATURE_STAGES = (0, 1, 2, 4, 6)   # timm features_only taps: the last block of every stride level (reductions 2 .. 32)


class SqueezeExcite(nn.Module):
    def __init__(self, channels, reduced):
        super().__init__()
        self.conv_reduce = nn.Conv2d(channels, reduced, 1)
        self.conv_expand = nn.Conv2d(reduced, channels, 1)

    def forward(self, x):
        s = x.mean((2, 3), keepdim=True)
        s = self.conv_expand(nn.functional.silu(self.conv_reduce(s)))
        return x * torch.sigmoid(s)


class DepthwiseSeparableConv(nn.Module):   # expand ratio 1
    def __init__(self, cin, cout, k, stride):
        super().__init__()
        self.conv_dw = nn.Conv2d(cin, cin, k, stride, k // 2, groups=cin, bias=False)
        self.bn1 = nn.BatchNorm2d(cin)
        self.se = SqueezeExcite(cin, max(1, cin // 4))
        self.conv_pw = nn.Conv2d(cin, cout, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        self.skip = stride == 1 and cin == cout

    def forward(self, x):
        y = nn.functional.silu(self.bn1(self.conv_dw(x)))
        y = self.bn2(self.conv_pw(self.se(y)))
        return x + y if self.skip else y


class InvertedResidual(nn.Module):
    def __init__(self, cin, cout, k, stride, expand):
        super().__init__()
        mid = cin * expand
        self.conv_pw = nn.Conv2d(cin, mid, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(mid)
        self.conv_dw = nn.Conv2d(mid, mid, k, stride, k // 2, groups=mid, bias=False)
        self.bn2 = nn.BatchNorm2d(mid)
        self.se = SqueezeExcite(mid, max(1, cin // 4))
        self.conv_pwl = nn.Conv2d(mid, cout, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(cout)
        self.skip = stride == 1 and cin == cout

    def forward(self, x):
        y = nn.functional.silu(self.bn1(self.conv_pw(x)))
        y = nn.functional.silu(self.bn2(self.conv_dw(y)))
        y = self.bn3(self.conv_pwl(self.se(y)))
        return x + y if self.skip else y


def build_blocks():
    blocks, cin = [], 32
    for expand, k, stride, cout, repeats in STAGES:
        stage = []
        for b in range(repeats):
            s = stride if b == 0 else 1
            stage.append(DepthwiseSeparableConv(cin, cout, k, s) if expand == 1 else InvertedResidual(cin, cout, k, s, expand))
            cin = cout
        blocks.append(nn.Sequential(*stage))
    return nn.Sequential(*blocks)


class EfficientNetB0FeaturesRef(nn.Module):
    """``forward(x[N, 3, H, W]) -> [f0 .. f4]`` (16 / 24 / 40 / 112 / 320 channels at strides 2 .. 32), eval-mode BatchNorm."""

    def __init__(self):
        super().__init__()
        self.conv_stem = nn.Conv2d(3, 32, 3, 2, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(32)
        self.blocks = build_blocks()

    def forward(self, x):
        x = nn.functional.silu(self.bn1(self.conv_stem(x)))
        feats = []
        for s, stage in enumerate(self.blocks):
            x = stage(x)
            if s in FEATURE_STAGES:
                feats.append(x)
        return feats


def torchvision_key_map():
    """timm-style key of this module -> torchvision ``efficientnet_b0().features`` key (same tensors, other names)."""
    m = {"conv_stem.weight": "0.0.weight"}
    bn = ("weight", "bias", "running_mean", "running_var", "num_batches_tracked")
    for f in bn:
        m[f"bn1.{f}"] = f"0.1.{f}"
    for s, (expand, k, stride, cout, repeats) in enumerate(STAGES):
        for b in range(repeats):
            t, o = f"{s + 1}.{b}.block", f"blocks.{s}.{b}"
            if expand == 1:
                parts = [("conv_dw", "bn1", 0), ("conv_pw", "bn2", 2)]
                se = 1
            else:
                parts = [("conv_pw", "bn1", 0), ("conv_dw", "bn2", 1), ("conv_pwl", "bn3", 3)]
                se = 2
            for conv, bnn, idx in parts:
                m[f"{o}.{conv}.weight"] = f"{t}.{idx}.0.weight"
                for f in bn:
                    m[f"{o}.{bnn}.{f}"] = f"{t}.{idx}.1.{f}"
            for mine, theirs in (("conv_reduce", "fc1"), ("conv_expand", "fc2")):
                for f in ("weight", "bias"):
                    m[f"{o}.se.{mine}.{f}"] = f"{t}.{se}.{theirs}.{f}"
    return m


def visual_features_ref(backbone, patch_rgb, mean, std):
    """``Eff_GAT.visual_features`` (efficient_gat.py:149-189) for ``model == "efficientnet_b0"``."""
    x = (patch_rgb - mean) / std
    feats = backbone(x)
    return torch.cat([feats[2].reshape(x.shape[0], -1), feats[3].reshape(x.shape[0], -1)], -1)
