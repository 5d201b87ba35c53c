"""Oracle restatement of ``torch_geometric.nn.TransformerConv`` (test infrastructure).

PyG is not vendored by the reference and not installable here; this restates its
documented behaviour for the constructor arguments the reference uses
(``Transformer_GNN.py:10-24``, ``exophormer_gnn.py:139-153``): ``concat=True,
beta=False, dropout=0, edge_dim=None, bias=True, root_weight=True``, aggregation
``add``, flow ``source_to_target`` (``edge_index[0]`` = source j, ``edge_index[1]``
= target i).  The edge-list formulation (gather -> segment softmax -> scatter-add)
mirrors what PyG executes, so this module is also the timed "reference
formulation restated" CPU baseline.
"""
import math

import torch
from torch import nn


def segment_softmax(src: torch.Tensor, index: torch.Tensor, num_nodes: int) -> torch.Tensor:
    """``torch_geometric.utils.softmax(src, index, num_nodes=N)``.

    out = exp(src - max_per_segment) / (sum_per_segment + 1e-16); segments are the
    edges that share a target node.  ``src`` is ``[E, H]``.
    """
    E = src.shape[0]
    if E == 0:
        return src.clone()
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    src_max = torch.full((num_nodes,) + tuple(src.shape[1:]), float("-inf"), dtype=src.dtype, device=src.device)
    src_max.scatter_reduce_(0, idx, src.detach(), reduce="amax", include_self=True)
    out = (src - src_max.index_select(0, index)).exp()
    out_sum = torch.zeros((num_nodes,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    out_sum.index_add_(0, index, out)
    out_sum = out_sum + 1e-16
    return out / out_sum.index_select(0, index)


class TransformerConvRef(nn.Module):
    """``TransformerConv(in_channels, out_channels, heads, concat=True)``.

    Parameter names follow PyG (``lin_key``, ``lin_query``, ``lin_value``,
    ``lin_skip``; weights ``[H*C, in]``) so a reference checkpoint's
    ``state_dict`` loads unchanged.
    """

    def __init__(self, in_channels: int, out_channels: int, heads: int = 1):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.heads = heads
        hc = heads * out_channels
        self.lin_key = nn.Linear(in_channels, hc)
        self.lin_query = nn.Linear(in_channels, hc)
        self.lin_value = nn.Linear(in_channels, hc)
        self.lin_skip = nn.Linear(in_channels, hc, bias=True)

    def forward(self, x, edge_index, return_attention_weights=None):
        H, C = self.heads, self.out_channels
        n = x.shape[0]
        q = self.lin_query(x).view(n, H, C)
        k = self.lin_key(x).view(n, H, C)
        v = self.lin_value(x).view(n, H, C)
        src, dst = edge_index[0], edge_index[1]
        q_i = q.index_select(0, dst)
        k_j = k.index_select(0, src)
        alpha = (q_i * k_j).sum(dim=-1) / math.sqrt(C)  # [E, H]
        alpha = segment_softmax(alpha, dst, n)
        msg = v.index_select(0, src) * alpha.view(-1, H, 1)
        out = torch.zeros((n, H, C), dtype=x.dtype, device=x.device)
        out.index_add_(0, dst, msg)
        out = out.reshape(n, H * C)
        out = out + self.lin_skip(x)
        if return_attention_weights:
            return out, (edge_index, alpha)
        return out
