"""Oracle restatement of the denoiser trunks (test infrastructure).

* ``EffGATRef``   follows ``puzzle_diff/model/backbones/efficient_gat.py:23-146``.
* ``EffGAT3dRef`` follows ``puzzle_diff/model/backbones/efficient_gat_3d.py:57-220``.

The visual / point-cloud encoders (timm, GrouPy, PointNet) are upstream of the
step loop and out of scope; both classes take pre-computed node features.
Parameter names equal the reference's so that a reference ``state_dict``
(minus ``visual_backbone.*`` / ``pcd_backbone.*``) loads with ``strict=False``.
"""
import torch
from torch import nn
from torch.nn import functional as F

from .gnn import ExophormerGNNRef, TransformerGNNRef
from .so3 import matrix_to_quaternion, skew_to_rmat

# efficient_gat.py:45-51
COMBINED_FEATURES_DIM = {
    "resnet18": 3136,
    "resnet50": 12352,
    "efficientnet_b0": 1088 + 32 + 32,
    "resnet18equiv": 1088 + 32 + 32,
}

# efficient_gat_3d.py:73-99
FEAT_DIM_3D = {
    "pointnet_inv": 1024,
    "pointnet": 128,
    "pointnet_plus": 256,
    "vn_dgcnn": 768,
    "vn_dgcnn_inv": 256,
    "vnn": 2104,
}


def _gnn(architecture, dim, n_layers, virt_nodes):
    if architecture == "transformer":
        return TransformerGNNRef(dim, hidden_dim=32 * 8, heads=8, output_size=dim, n_layers=n_layers)
    if architecture == "exophormer":
        return ExophormerGNNRef(dim, hidden_dim=32 * 8, heads=8, output_size=dim, n_layers=n_layers, virt_nodes=virt_nodes)
    raise NotImplementedError(f"architecture {architecture!r} is out of scope (SURVEY.md section 2.1)")


class EffGATRef(nn.Module):
    def __init__(
        self,
        steps,
        input_channels=2,
        output_channels=2,
        n_layers=4,
        model="efficientnet_b0",
        architecture="transformer",
        virt_nodes=4,
    ):
        super().__init__()
        self.model = model
        self.combined_features_dim = COMBINED_FEATURES_DIM[model]
        self.input_channels = input_channels
        self.output_channels = output_channels
        D = self.combined_features_dim
        self.gnn_backbone = _gnn(architecture, D, n_layers, virt_nodes)
        self.time_emb = nn.Embedding(steps, 32)
        self.pos_mlp = nn.Sequential(nn.Linear(input_channels, 16), nn.GELU(), nn.Linear(16, 32))
        self.final_mlp = nn.Sequential(nn.Linear(D, 32), nn.GELU(), nn.Linear(32, output_channels))
        self.mlp = nn.Sequential(nn.Linear(D, 128), nn.GELU(), nn.Linear(128, D))
        # present-but-unused parameters of the reference (efficient_gat.py:105-112)
        self.linear1 = nn.Linear(8192, 544)
        self.linear2 = nn.Linear(4096, 544)
        self.register_buffer("mean", torch.tensor([0.4850, 0.4560, 0.4060])[None, :, None, None])
        self.register_buffer("std", torch.tensor([0.2290, 0.2240, 0.2250])[None, :, None, None])
        # efficient_gat.py:40-42: timm.create_model(model, features_only=True); restated for efficientnet_b0
        # (oracle/efficientnet.py).  Created LAST so that the denoiser's default-initialised weights under a given
        # torch.manual_seed stay what they were when the self-generated fixtures of tests/golden/make_golden.py were made.
        self.visual_backbone = None
        if model == "efficientnet_b0":
            from .efficientnet import EfficientNetB0FeaturesRef

            self.visual_backbone = EfficientNetB0FeaturesRef()

    def visual_features(self, patch_rgb):  # :149-189 (frozen backbone, eval-mode BatchNorm)
        from .efficientnet import visual_features_ref

        return visual_features_ref(self.visual_backbone.eval(), patch_rgb, self.mean, self.std)

    def forward_with_feats(self, xy_pos, time, patch_rgb, edge_index, patch_feats, batch):
        time_feats = self.time_emb(time)  # :131
        pos_feats = self.pos_mlp(xy_pos)  # :132
        combined_feats = torch.cat([patch_feats, pos_feats, time_feats], -1)  # :134
        combined_feats = self.mlp(combined_feats)  # :135
        feats, attentions = self.gnn_backbone(x=combined_feats, edge_index=edge_index, batch=batch)  # :138
        final_feats = self.final_mlp(feats + combined_feats)  # :144
        return final_feats, attentions


class EffGAT3dRef(nn.Module):
    def __init__(
        self,
        steps,
        input_channels=7,
        t_channels=3,
        r_channels=3,
        n_layers=4,
        architecture="transformer",
        virt_nodes=8,
        backbone="pointnet",
    ):
        super().__init__()
        feat_dim = FEAT_DIM_3D[backbone]
        self.combined_features_dim = feat_dim + 32 + 32
        self.gnn_feat_dim = self.combined_features_dim
        self.input_channels = input_channels
        D = self.gnn_feat_dim
        self.gnn_backbone = _gnn(architecture, D, n_layers, virt_nodes)
        self.time_emb = nn.Embedding(steps, 32)
        self.pos_mlp = nn.Sequential(nn.Linear(input_channels, 16), nn.GELU(), nn.Linear(16, 32))
        self.mlp = nn.Sequential(
            nn.Linear(self.combined_features_dim, 256), nn.LeakyReLU(0.2), nn.Linear(256, D), nn.LeakyReLU(0.2)
        )
        self.mlp_t = nn.Sequential(nn.Linear(D, 256), nn.GELU(), nn.Linear(256, t_channels))
        self.mlp_r = nn.Sequential(nn.Linear(D, 256), nn.GELU(), nn.Linear(256, r_channels))

    def forward_with_feats(self, xy_pos, time, edge_index, pcd_feats, batch):
        time_feats = self.time_emb(time)  # :181
        pos_feats = self.pos_mlp(xy_pos)  # :182
        combined_feats = torch.cat([pcd_feats, pos_feats, time_feats], -1)
        combined_feats = self.mlp(combined_feats)
        feats, attentions = self.gnn_backbone(x=combined_feats, edge_index=edge_index, batch=batch)
        t_pred = self.mlp_t(feats + combined_feats)  # :211
        r_pred = self.mlp_r(feats + combined_feats)  # :213
        r_pred = matrix_to_quaternion(skew_to_rmat(r_pred))  # :217
        r_pred = F.normalize(r_pred, p=2, dim=-1)  # :218
        return torch.hstack((r_pred, t_pred)), attentions  # :220
