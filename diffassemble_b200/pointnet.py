"""PointNet fragment encoder on the CUDA operators (scope row N4, 3-D side).

Mirror of ``puzzle_diff/model/backbones/pointnet.py:8-43`` (``PointNet(feat_dim, global_feat=True)``): five
point-wise ``Conv1d(kernel_size=1, bias=False)`` + ``BatchNorm1d`` layers (ReLU after the first four) and a
global max over the points of a fragment.  Same parameter / buffer names, so the ``model.pcd_backbone.*``
entries of a reference checkpoint load unchanged.  In eval mode -- the only mode the sampling path uses
(``efficient_gat_3d.py:231-236`` with ``freeze_backbone``) -- BatchNorm is an affine map of its running
statistics and is folded into the layer: every layer is ONE ``da_op_linear`` launch (exact-fp32 GEMM with fused
bias + ReLU) over all points of all fragments, followed by ONE ``da_op_segment_max`` launch.  No torch
arithmetic on the device path; CPU tensors raise.
"""
import torch
from torch import Tensor, nn

from .engine import op_linear, op_segment_max

ACT_NONE, ACT_RELU = 0, 3


class PointNet(nn.Module):
    def __init__(self, feat_dim, global_feat=True):
        super().__init__()
        dims = [3, 64, 64, 64, 128, feat_dim]
        for i in range(5):
            setattr(self, f"conv{i + 1}", nn.Conv1d(dims[i], dims[i + 1], kernel_size=1, bias=False))
            setattr(self, f"bn{i + 1}", nn.BatchNorm1d(dims[i + 1]))
        self.global_feat = global_feat
        self._folded_cache, self._folded_key = None, None

    def _folded(self):
        """Eval-mode BatchNorm folded into the 1x1 convolutions: five (W', b') pairs with y = x W'^T + b', computed once
        per weight version (not on every forward); the first layer's K = 3 is padded to 4 (16-byte rows)."""
        key = tuple((k, v.data_ptr(), v._version) for k, v in self.state_dict().items())
        if self._folded_cache is None or key != self._folded_key:
            out = []
            for i in range(1, 6):
                conv, bn = getattr(self, f"conv{i}"), getattr(self, f"bn{i}")
                scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
                w = conv.weight[:, :, 0] * scale[:, None]
                b = bn.bias - bn.running_mean * scale
                if i == 1:
                    w = torch.cat([w, w.new_zeros(w.shape[0], 1)], 1)
                out.append((w.float().contiguous(), b.float().contiguous()))
            self._folded_cache, self._folded_key = out, key
        return self._folded_cache

    @torch.no_grad()
    def forward(self, x: Tensor) -> Tensor:
        """x: [B, N, 3] point clouds (one per fragment) -> [B, feat_dim] (or [B, N, feat_dim])."""
        if self.training:
            raise NotImplementedError("the CUDA PointNet encoder implements eval-mode BatchNorm (frozen backbone) only")
        if not x.is_cuda:
            raise RuntimeError("PointNet was called with CPU tensors: the B200 encoder has no CPU path")
        B, N, _ = x.shape
        h = torch.zeros((B * N, 4), dtype=torch.float32, device=x.device)   # K padded to 4 (16-byte rows)
        h[:, :3] = x.reshape(B * N, 3)
        for i, (w, b) in enumerate(self._folded(), start=1):
            h = op_linear(h, w, b, act=ACT_RELU if i < 5 else ACT_NONE, mode="fp32")
        if not self.global_feat:
            return h.reshape(B, N, -1)
        seg = torch.arange(0, (B + 1) * N, N, dtype=torch.int32, device=x.device)
        return op_segment_max(h, seg)
