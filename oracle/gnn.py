"""Oracle restatement of the two graph-transformer backbones (test infrastructure).

* ``TransformerGNNRef``  follows ``puzzle_diff/model/backbones/Transformer_GNN.py:5-46``.
* ``ExophormerGNNRef``   follows ``puzzle_diff/model/backbones/exophormer_gnn.py:132-215``
  including its virtual-node wiring exactly as shipped (mis-aligned src/dst
  concatenation, duplicate virtual->virtual edges, cross-graph edges when B>1).
"""
import torch
from torch import nn

from .transformer_conv import TransformerConvRef


def _layers(input_size, hidden_dim, heads, output_size, n_layers):
    # Transformer_GNN.py:9-25 / exophormer_gnn.py:138-154
    return nn.ModuleList(
        [TransformerConvRef(input_size, hidden_dim // heads, heads)]
        + [TransformerConvRef(hidden_dim, hidden_dim // heads, heads) for _ in range(n_layers - 2)]
        + [TransformerConvRef(hidden_dim, output_size // heads, heads)]
    )


class TransformerGNNRef(nn.Module):
    def __init__(self, input_size, hidden_dim, heads, output_size, n_layers=4):
        super().__init__()
        self.module_list = _layers(input_size, hidden_dim, heads, output_size, n_layers)
        self.n_layers = n_layers

    def forward(self, x, edge_index, move_to_cpu=False, batch=None, *args):
        # Transformer_GNN.py:29-46: GELU (exact erf) after every layer but the last.
        attentions = []
        for i in range(self.n_layers - 1):
            x, atts = self.module_list[i](x, edge_index, return_attention_weights=True)
            x = nn.functional.gelu(x)
            attentions.append(atts)
        x, atts = self.module_list[-1](x, edge_index, return_attention_weights=True)
        attentions.append(atts)
        return x, attentions


def exophormer_wiring(edge_index: torch.Tensor, batch: torch.Tensor, virt_nodes: int):
    """Index-only replay of ``exophormer_gnn.py:164-200``.

    Returns ``(virtual_ids[V*B], batch_ext, edge_index_ext)`` where ``virtual_ids``
    are the embedding rows appended after the real nodes.
    """
    n_graphs = int(batch.max()) + 1
    num_real = len(batch)
    virtual_ids = torch.arange(virt_nodes).repeat(n_graphs)  # :169
    batch_ext = torch.cat((batch, torch.arange(n_graphs).repeat(virt_nodes)))  # :180-182
    virt_edges = []
    for i in batch_ext.unique():  # :185
        num_nodes = int((batch_ext == i).sum())  # counts real + virtual rows of graph i (:186)
        i = int(i)
        virt_edge = torch.arange(num_real + i * virt_nodes, num_real + (i + 1) * virt_nodes).repeat(num_nodes)
        virt_edges.append(virt_edge)
    virt_edges = torch.cat(virt_edges)
    src = torch.cat([torch.arange(num_real), virt_edges])  # :198
    dst = torch.cat([virt_edges, torch.arange(num_real)])  # :199
    edge_index_ext = torch.hstack((edge_index, torch.stack((src, dst))))  # :200
    return virtual_ids, batch_ext, edge_index_ext


class ExophormerGNNRef(nn.Module):
    def __init__(self, input_size, hidden_dim, heads, output_size, n_layers=4, virt_nodes=4):
        super().__init__()
        self.module_list = _layers(input_size, hidden_dim, heads, output_size, n_layers)
        self.virt_nodes = virt_nodes
        if self.virt_nodes > 0:
            self.virt_node_embedding = nn.Embedding(virt_nodes, input_size)
        self.n_layers = n_layers

    def forward(self, x, edge_index, move_to_cpu=False, batch=None, mean_value=False, *args):
        attentions = []
        num_real_nodes = len(batch)
        if self.virt_nodes > 0:
            virtual_ids, batch, edge_index = exophormer_wiring(edge_index, batch, self.virt_nodes)
            if mean_value:  # :172-173
                virt_h = x.mean(dim=0).unsqueeze(0).repeat(len(virtual_ids), 1)
            else:
                virt_h = self.virt_node_embedding(virtual_ids)
            x = torch.cat((x, virt_h))
        # :202-207 -- no activation between layers, attention returned for the last only
        for i in range(self.n_layers - 1):
            x = self.module_list[i](x, edge_index)
        x, atts = self.module_list[-1](x, edge_index, return_attention_weights=True)
        if self.virt_nodes > 0:
            x = x[:num_real_nodes]
        attentions.append(atts)
        return x, attentions
