"""Oracle restatement of the PointNet fragment encoder (test infrastructure).

Semantics of ``puzzle_diff/model/backbones/pointnet.py:8-43`` (``PointNet(feat_dim, global_feat=True)``) in plain
torch: five bias-free point-wise convolutions 3 -> 64 -> 64 -> 64 -> 128 -> feat_dim, each followed by a
BatchNorm1d, ReLU after the first four, then the maximum over the points of a fragment.  Parameter names
(``conv{k}`` / ``bn{k}``) are the reference's so that ``state_dict``s interchange.  The reference file is
importable in the build container, so ``tests/golden/ref_pointnet.pt`` comes from the reference class itself and
``tests/test_oracle_pinned.py`` pins this restatement to it.
"""
import torch
from torch import nn

WIDTHS = (3, 64, 64, 64, 128)


class PointNetRef(nn.Module):
    def __init__(self, feat_dim, global_feat=True):
        super().__init__()
        widths = WIDTHS + (feat_dim,)
        for k in range(1, len(widths)):
            self.add_module(f"conv{k}", nn.Conv1d(widths[k - 1], widths[k], kernel_size=1, bias=False))
        for k in range(1, len(widths)):
            self.add_module(f"bn{k}", nn.BatchNorm1d(widths[k]))
        self.n_layers = len(widths) - 1
        self.global_feat = global_feat

    def forward(self, points):
        """points: [fragments, N, 3] -> [fragments, feat_dim] (global) or [fragments, N, feat_dim]."""
        h = points.permute(0, 2, 1).contiguous()            # channels first for Conv1d
        for k in range(1, self.n_layers + 1):
            h = getattr(self, f"bn{k}")(getattr(self, f"conv{k}")(h))
            if k < self.n_layers:
                h = torch.relu(h)
        if self.global_feat:
            return h.amax(dim=-1)
        return h.permute(0, 2, 1).contiguous()
