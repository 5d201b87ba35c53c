"""diffassemble_b200: B200-native (sm_100a) DiffAssemble denoiser + DDPM/DDIM sampling loop.

The public names mirror the reference's (``puzzle_diff/model``): ``GNN_Diffusion``,
``Eff_GAT``, ``Eff_GAT_3d``, ``Transformer_GNN``, ``Exophormer_GNN``, ``ModelMeanType``,
``ModelScheduler``.  All arithmetic runs in ``lib/libdiffassemble_b200.so`` (built by
``python -m diffassemble_b200.build``) behind the C ABI in ``include/diffassemble_b200.h``.
"""
from .backbones import Eff_GAT, Eff_GAT_3d, Exophormer_GNN, Transformer_GNN, TransformerConv  # noqa: F401
from .engine import DenoiserEngine, op_graph_attention, op_graph_attention_dense, op_linear, op_segment_max  # noqa: F401
from .efficientnet import EfficientNetB0Features  # noqa: F401
from .pointnet import PointNet  # noqa: F401
from .spatial_diffusion import (  # noqa: F401
    GNN_Diffusion,
    ModelMeanType,
    ModelScheduler,
    cosine_beta_schedule,
    cosine_discrete_beta_schedule,
    extract,
    linear_beta_schedule,
)
from .spatial_diffusion_3d import GNN_Diffusion_3d  # noqa: F401
from . import metrics, sharding, topology  # noqa: F401
from .metrics import greedy_cost_assignment, greedy_cost_assignment_batched  # noqa: F401

__all__ = [
    "GNN_Diffusion", "GNN_Diffusion_3d", "Eff_GAT", "Eff_GAT_3d", "Transformer_GNN", "Exophormer_GNN",
    "TransformerConv", "ModelMeanType", "ModelScheduler", "DenoiserEngine", "sharding", "topology",
]
