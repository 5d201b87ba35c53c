"""Oracle restatement of the PointNet fragment encoder (test infrastructure).

Follows ``puzzle_diff/model/backbones/pointnet.py:8-43`` line by line in plain torch (the reference file is
importable, so ``tests/golden/ref_pointnet.pt`` is produced by the reference class itself and pins this one)."""
import torch.nn.functional as F
from torch import nn


class PointNetRef(nn.Module):
    def __init__(self, feat_dim, global_feat=True):
        super().__init__()
        self.conv1 = nn.Conv1d(3, 64, kernel_size=1, bias=False)
        self.conv2 = nn.Conv1d(64, 64, kernel_size=1, bias=False)
        self.conv3 = nn.Conv1d(64, 64, kernel_size=1, bias=False)
        self.conv4 = nn.Conv1d(64, 128, kernel_size=1, bias=False)
        self.conv5 = nn.Conv1d(128, feat_dim, kernel_size=1, bias=False)
        self.bn1, self.bn2, self.bn3 = nn.BatchNorm1d(64), nn.BatchNorm1d(64), nn.BatchNorm1d(64)
        self.bn4, self.bn5 = nn.BatchNorm1d(128), nn.BatchNorm1d(feat_dim)
        self.global_feat = global_feat

    def forward(self, x):  # x: [B, N, 3]
        x = x.transpose(2, 1).contiguous()
        x = F.relu(self.bn1(self.conv1(x)))
        x = F.relu(self.bn2(self.conv2(x)))
        x = F.relu(self.bn3(self.conv3(x)))
        x = F.relu(self.bn4(self.conv4(x)))
        x = self.bn5(self.conv5(x))
        if self.global_feat:
            return x.max(dim=-1)[0]
        return x.transpose(2, 1).contiguous()
