"""EfficientNet-B0 patch encoder on the CUDA operators (scope row N4, 2-D side).

Mirror of what the reference builds with ``timm.create_model("efficientnet_b0", features_only=True)``
(``puzzle_diff/model/backbones/efficient_gat.py:40-42``) and reads in ``Eff_GAT.visual_features`` (``:149-189``):
feature maps 2 and 3 of the pyramid (40 channels at stride 8 and 112 channels at stride 16; 4x4 and 2x2 pixels for the
32x32 puzzle patches), flattened channel-major and concatenated to 1088 values per patch.  Same parameter / buffer names
as timm's model (``conv_stem``, ``bn1``, ``blocks.S.B.conv_pw / bn1 / conv_dw / bn2 / se.conv_reduce / se.conv_expand /
conv_pwl / bn3``), so the ``model.visual_backbone.*`` entries of a reference checkpoint load unchanged; stages 5 and 6
(192 / 320 channels, stride 32) are parameter storage only -- the reference never reads their output.

Execution (eval mode = frozen backbone, the shipped setting ``freeze_backbone=True``): NHWC fp32, eval-mode BatchNorm
folded into the preceding convolution once per weight version, every point-wise convolution and squeeze-excite FC is ONE
``da_op_linear`` launch (bias + SiLU / sigmoid fused), the stem / depth-wise convolutions and the squeeze / excite
reductions are the ``da_op_conv2d_nhwc`` / ``da_op_dwconv2d_nhwc`` / ``da_op_spatial_mean`` / ``da_op_channel_scale``
kernels (``csrc/conv.cu``).  Patches are processed in chunks so the widest activation (96 channels at 16x16) stays
below ~1 GB.  No torch arithmetic on the device path apart from layout glue (NCHW <-> NHWC views); CPU tensors raise.
"""
import ctypes as C
from typing import List

import torch
from torch import Tensor, nn

from . import _cabi
from .engine import _ptr, _stream, op_linear

ACT_NONE, ACT_SILU, ACT_SIGMOID = 0, 4, 5
# (expand ratio, kernel, stride, out channels, repeats) per stage -- EfficientNet-B0
STAGES = [(1, 3, 1, 16, 1), (6, 3, 2, 24, 2), (6, 5, 2, 40, 2), (6, 3, 2, 80, 3), (6, 5, 1, 112, 3), (6, 5, 2, 192, 4), (6, 3, 1, 320, 1)]
LAST_COMPUTED_STAGE = 4   # feature maps 2 (after stage 2) and 3 (after stage 4) are the ones visual_features reads


class _SE(nn.Module):
    def __init__(self, channels, reduced):
        super().__init__()
        self.conv_reduce = nn.Conv2d(channels, reduced, 1)
        self.conv_expand = nn.Conv2d(reduced, channels, 1)


class _DSBlock(nn.Module):   # timm DepthwiseSeparableConv (expand ratio 1)
    def __init__(self, cin, cout, k, stride):
        super().__init__()
        self.conv_dw = nn.Conv2d(cin, cin, k, stride, k // 2, groups=cin, bias=False)
        self.bn1 = nn.BatchNorm2d(cin)
        self.se = _SE(cin, max(1, cin // 4))
        self.conv_pw = nn.Conv2d(cin, cout, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(cout)
        self.cfg = (cin, cin, cout, k, stride, stride == 1 and cin == cout)


class _MBBlock(nn.Module):   # timm InvertedResidual
    def __init__(self, cin, cout, k, stride, expand):
        super().__init__()
        mid = cin * expand
        self.conv_pw = nn.Conv2d(cin, mid, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(mid)
        self.conv_dw = nn.Conv2d(mid, mid, k, stride, k // 2, groups=mid, bias=False)
        self.bn2 = nn.BatchNorm2d(mid)
        self.se = _SE(mid, max(1, cin // 4))
        self.conv_pwl = nn.Conv2d(mid, cout, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(cout)
        self.cfg = (cin, mid, cout, k, stride, stride == 1 and cin == cout)


def _fold(conv_w: Tensor, bn: nn.BatchNorm2d):
    """Eval-mode BatchNorm folded into the convolution in front of it: per-output-channel scale and shift."""
    scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    return conv_w * scale.reshape(-1, *([1] * (conv_w.dim() - 1))), bn.bias - bn.running_mean * scale


def _check(st, what):
    if st != _cabi.DA_OK:
        raise _cabi.DiffAssembleError(st, f"{what} failed")


class EfficientNetB0Features(nn.Module):
    def __init__(self, chunk: int = 4096):
        super().__init__()
        self.conv_stem = nn.Conv2d(3, 32, 3, 2, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(32)
        blocks, cin = [], 32
        for expand, k, stride, cout, repeats in STAGES:
            stage = []
            for b in range(repeats):
                s = stride if b == 0 else 1
                stage.append(_DSBlock(cin, cout, k, s) if expand == 1 else _MBBlock(cin, cout, k, s, expand))
                cin = cout
            blocks.append(nn.Sequential(*stage))
        self.blocks = nn.Sequential(*blocks)
        self.chunk = chunk
        self._packed, self._packed_key = None, None

    # -- weights: folded + laid out for the kernels, once per weight version ---------------------------------
    def _pack(self):
        key = tuple((k, v.data_ptr(), v._version) for k, v in self.state_dict().items())
        if self._packed is not None and key == self._packed_key:
            return self._packed
        with torch.no_grad():
            def pw(conv, bn):   # 1x1 convolution -> linear [Cout, Cin]
                w, b = _fold(conv.weight[:, :, 0, 0], bn)
                return w.contiguous().float(), b.contiguous().float()

            def dw(conv, bn):   # depth-wise: [C, 1, k, k] -> [k, k, C]
                w, b = _fold(conv.weight[:, 0], bn)
                return w.permute(1, 2, 0).contiguous().float(), b.contiguous().float()

            def se(m):          # the two FCs; the reduced width is padded to a multiple of 4 (16-byte rows)
                cr, c = m.conv_reduce.weight.shape[0], m.conv_reduce.weight.shape[1]
                crp = (cr + 3) // 4 * 4
                w1 = torch.zeros((crp, c), device=m.conv_reduce.weight.device); w1[:cr] = m.conv_reduce.weight[:, :, 0, 0]
                b1 = torch.zeros((crp,), device=w1.device); b1[:cr] = m.conv_reduce.bias
                w2 = torch.zeros((c, crp), device=w1.device); w2[:, :cr] = m.conv_expand.weight[:, :, 0, 0]
                return w1.float(), b1.float(), w2.contiguous().float(), m.conv_expand.bias.contiguous().float()

            ws, bs = _fold(self.conv_stem.weight, self.bn1)
            packed = {"stem": (ws.permute(0, 2, 3, 1).contiguous().float(), bs.contiguous().float()), "blocks": []}
            for s in range(LAST_COMPUTED_STAGE + 1):
                for blk in self.blocks[s]:
                    e = {"cfg": blk.cfg, "dw": dw(blk.conv_dw, blk.bn2 if isinstance(blk, _MBBlock) else blk.bn1), "se": se(blk.se)}
                    if isinstance(blk, _MBBlock):
                        e["expand"] = pw(blk.conv_pw, blk.bn1)
                        e["project"] = pw(blk.conv_pwl, blk.bn3)
                    else:
                        e["expand"] = None
                        e["project"] = pw(blk.conv_pw, blk.bn2)
                    packed["blocks"].append((s, e))
        self._packed, self._packed_key = packed, key
        return packed

    # -- kernels ------------------------------------------------------------------------------------------------
    @staticmethod
    def _conv_stem(lib, x, w, b, N, H, W):
        Ho, Wo = (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1
        y = torch.empty((N, Ho, Wo, 32), dtype=torch.float32, device=x.device)
        _check(lib.da_op_conv2d_nhwc(_ptr(x), _ptr(w), _ptr(b), _ptr(y), N, H, W, 3, 32, 3, 2, 1, ACT_SILU, _stream(x.device)), "da_op_conv2d_nhwc")
        return y

    @staticmethod
    def _dwconv(lib, x, w, b, k, stride):
        N, H, W, Cc = x.shape
        Ho, Wo = (H + 2 * (k // 2) - k) // stride + 1, (W + 2 * (k // 2) - k) // stride + 1
        y = torch.empty((N, Ho, Wo, Cc), dtype=torch.float32, device=x.device)
        _check(lib.da_op_dwconv2d_nhwc(_ptr(x), _ptr(w), _ptr(b), _ptr(y), N, H, W, Cc, k, stride, k // 2, ACT_SILU, _stream(x.device)),
               "da_op_dwconv2d_nhwc")
        return y

    @staticmethod
    def _squeeze_excite(lib, x, se):
        w1, b1, w2, b2 = se
        N, H, W, Cc = x.shape
        pooled = torch.empty((N, Cc), dtype=torch.float32, device=x.device)
        _check(lib.da_op_spatial_mean(_ptr(x), _ptr(pooled), Cc, N, H * W, Cc, _stream(x.device)), "da_op_spatial_mean")
        gate = op_linear(op_linear(pooled, w1, b1, act=ACT_SILU, mode="fp32"), w2, b2, act=ACT_SIGMOID, mode="fp32")
        _check(lib.da_op_channel_scale(_ptr(x), _ptr(gate), Cc, N, H * W, Cc, _stream(x.device)), "da_op_channel_scale")
        return x

    def _forward_chunk(self, lib, packed, x_nhwc) -> List[Tensor]:
        N, H, W, _ = x_nhwc.shape
        with torch.cuda.device(x_nhwc.device):
            h = self._conv_stem(lib, x_nhwc, *packed["stem"], N, H, W)
            feats = {}
            for s, e in packed["blocks"]:
                cin, mid, cout, k, stride, skip = e["cfg"]
                inp = h
                n_, hh, ww, _ = h.shape
                if e["expand"] is not None:
                    h = op_linear(h.reshape(n_ * hh * ww, cin), *e["expand"], act=ACT_SILU, mode="fp32").reshape(n_, hh, ww, mid)
                h = self._dwconv(lib, h, *e["dw"], k, stride)
                h = self._squeeze_excite(lib, h, e["se"])
                n_, hh, ww, _ = h.shape
                h = op_linear(h.reshape(n_ * hh * ww, mid), *e["project"], act=ACT_NONE, mode="fp32").reshape(n_, hh, ww, cout)
                if skip:   # residual of the repeated blocks
                    _check(lib.da_op_add_inplace(_ptr(h), _ptr(inp), h.numel(), _stream(h.device)), "da_op_add_inplace")
                feats[s] = h
        return [feats[0], feats[1], feats[2], feats[4]]

    @torch.no_grad()
    def forward(self, x: Tensor, mean: Tensor = None, std: Tensor = None) -> List[Tensor]:
        """x: [N, 3, H, W] patches -> the first four maps of the feature pyramid, NCHW like timm's (16 / 24 / 40 / 112
        channels at strides 2 / 4 / 8 / 16); the stride-32 map the reference never reads is not computed.  With ``mean`` /
        ``std`` ([3] or broadcastable) the reference's input normalisation ``(x - mean) / std`` (efficient_gat.py:150) is
        applied inside the layout kernel; without them ``x`` is taken as already normalised."""
        if self.training:
            raise NotImplementedError("the CUDA EfficientNet encoder implements eval-mode BatchNorm (frozen backbone) only")
        if not x.is_cuda:
            raise RuntimeError("EfficientNetB0Features was called with CPU tensors: the B200 encoder has no CPU path")
        lib = _cabi.load_library()
        packed = self._pack()
        mean = (torch.zeros(3, device=x.device) if mean is None else mean.to(x.device)).reshape(-1).float().contiguous()
        std = (torch.ones(3, device=x.device) if std is None else std.to(x.device)).reshape(-1).float().contiguous()
        outs = None
        for i in range(0, x.shape[0], self.chunk):
            xs = x[i:i + self.chunk].float().contiguous()
            n, c, hh, ww = xs.shape
            xc = torch.empty((n, hh, ww, c), dtype=torch.float32, device=x.device)   # NHWC, normalised
            with torch.cuda.device(x.device):
                _check(lib.da_op_normalize_to_nhwc(_ptr(xs), _ptr(mean), _ptr(std), _ptr(xc), n, c, hh * ww, _stream(x.device)),
                       "da_op_normalize_to_nhwc")
            fs = [f.permute(0, 3, 1, 2) for f in self._forward_chunk(lib, packed, xc)]
            if outs is None:
                outs = [[f] for f in fs]
            else:
                for o, f in zip(outs, fs):
                    o.append(f)
        return [torch.cat(o) if len(o) > 1 else o[0] for o in outs]
