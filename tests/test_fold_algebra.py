"""The weight folding of the inference path (csrc/fold.cu) is an ALGEBRAIC rewrite of
``Eff_GAT.forward_with_feats`` (efficient_gat.py:121-146).  This CPU test restates the folded evaluation order with the
oracle's own modules in float64 and holds it to the oracle's literal forward -- independent of any CUDA code:

  * ``mlp[2]`` (Linear(128, D), no activation) composed into the four projections of the first TransformerConv, the
    virtual rows of the exophormer wiring entering through their embedding (the one-hot columns of ``fold0``);
  * ``final_mlp[0]`` (Linear(D, 32)) composed into the last TransformerConv's value / skip projections and into the trunk
    residual: the attention aggregates 32-channel values per head, and neither ``combined`` nor the last layer's D-wide
    output nor the head GEMM exist.

The CUDA side of the same statement is ``tests/test_gpu_benchmark_config.py`` (folded vs literal pipeline vs live oracle).
"""
import math

import pytest
import torch

import oracle
from common import reseed_parameters, synth_graph_batch
from oracle.gnn import exophormer_wiring
from oracle.transformer_conv import segment_softmax


def _attention(q, k, v, edge_index, H):
    """sum_j alpha_ij v_j per head for arbitrary value width (TransformerConv's message / softmax / aggregate stage)."""
    n = q.shape[0]
    C, Cv = q.shape[1] // H, v.shape[1] // H
    src, dst = edge_index[0], edge_index[1]
    a = (q.view(n, H, C).index_select(0, dst) * k.view(n, H, C).index_select(0, src)).sum(-1) / math.sqrt(C)
    a = segment_softmax(a, dst, n)
    out = torch.zeros((n, H, Cv), dtype=q.dtype)
    out.index_add_(0, dst, v.view(n, H, Cv).index_select(0, src) * a.view(-1, H, 1))
    return out   # [n, H, Cv]


def _folded_forward(m, xy_pos, time, edge_index, patch_feats, batch):
    gnn, H = m.gnn_backbone, 8
    exo = isinstance(gnn, oracle.ExophormerGNNRef) and gnn.virt_nodes > 0
    W1, b1 = m.mlp[0].weight, m.mlp[0].bias
    W2, b2 = m.mlp[2].weight, m.mlp[2].bias
    Wa, ba = m.final_mlp[0].weight, m.final_mlp[0].bias
    M = xy_pos.shape[0]
    h = torch.nn.functional.gelu(torch.cat([patch_feats, m.pos_mlp(xy_pos), m.time_emb(time)], -1) @ W1.T + b1)   # [M, 128]
    ei = edge_index
    if exo:
        vids, _, ei = exophormer_wiring(edge_index, batch, gnn.virt_nodes)
        virt = gnn.virt_node_embedding(vids)                                                                       # [V * B, D]
    convs = list(gnn.module_list)
    gelu_between = isinstance(gnn, oracle.TransformerGNNRef)
    # ---- first layer: [Q | K | V | skip]_0 = h (W_0 W_2)^T + (W_0 b_2 + b_0); virtual rows: W_0 emb + b_0 ----
    c0 = convs[0]
    parts = []
    for lin in (c0.lin_query, c0.lin_key, c0.lin_value, c0.lin_skip):
        real = h @ (lin.weight @ W2).T + (lin.weight @ b2 + lin.bias)
        parts.append(torch.cat([real, virt @ lin.weight.T + lin.bias]) if exo else real)
    q, k, v, s = parts
    x = _attention(q, k, v, ei, H).reshape(q.shape[0], -1) + s
    if gelu_between:
        x = torch.nn.functional.gelu(x)
    for conv in convs[1:-1]:
        x = conv(x, ei)
        if gelu_between:
            x = torch.nn.functional.gelu(x)
    # ---- last layer + final_mlp[0]: per-head aggregates of V' = x (W_a^h W_v^h)^T + W_a^h b_v^h (32 channels per head) ----
    cl = convs[-1]
    D = Wa.shape[1]
    C = D // H
    q, k = cl.lin_query(x), cl.lin_key(x)
    vprime = []
    for hh in range(H):
        Wah = Wa[:, hh * C:(hh + 1) * C]                                  # [32, C]
        Wvh, bvh = cl.lin_value.weight[hh * C:(hh + 1) * C], cl.lin_value.bias[hh * C:(hh + 1) * C]
        vprime.append(x @ (Wah @ Wvh).T + Wah @ bvh)                      # [n, 32]
    vprime = torch.stack(vprime, 1).reshape(x.shape[0], H * 32)
    partial = _attention(q, k, vprime, ei, H)[:M]                         # [M, H, 32]
    u = partial.sum(1) + x[:M] @ (Wa @ cl.lin_skip.weight).T + h @ (Wa @ W2).T + (ba + Wa @ (cl.lin_skip.bias + b2))
    return m.final_mlp[2](torch.nn.functional.gelu(u))


@pytest.mark.parametrize("arch,V,sizes", [("exophormer", 4, [40, 25]), ("exophormer", 0, [30]), ("transformer", 0, [36, 20, 7])])
def test_folded_evaluation_order_equals_the_literal_forward(arch, V, sizes):
    torch.manual_seed(0)
    m = oracle.EffGATRef(steps=300, input_channels=4, output_channels=4, architecture=arch, virt_nodes=V).double().eval()
    reseed_parameters(m, 31)
    m = m.double()
    ei, batch = synth_graph_batch(sizes, kind="expander", degree="60%", seed=5)
    M = sum(sizes)
    g = torch.Generator().manual_seed(1)
    feats = torch.randn(M, 1088, generator=g, dtype=torch.float64)
    x = torch.randn(M, 4, generator=g, dtype=torch.float64)
    t = torch.randint(0, 300, (M,), generator=g)
    with torch.no_grad():
        want, _ = m.forward_with_feats(x, t, None, ei, feats, batch)
        got = _folded_forward(m, x, t, ei, feats, batch)
    err = float((got - want).abs().max() / want.abs().max())
    assert err < 1e-11, (arch, V, err)


def test_table_form_of_the_prologue_equals_the_literal_first_linear():
    """`prologue_table_kernel` (csrc/pointwise.cu): h = act(P + tt[t] + (W_1[:, pos] W_p2) GELU(pos_mlp[0](x))) with
    P = feats W_1[:, :Dv]^T + b_1 (hoisted, `da_set_features`), tt[t] = time_emb[t] W_1[:, time]^T + W_1[:, pos] b_p2 and the
    composed [Hm, 16] matrix built once per `da_load_weights` (csrc/api.cu) -- against efficient_gat.py:131-135."""
    torch.manual_seed(0)
    m = oracle.EffGATRef(steps=300, input_channels=4, output_channels=4, architecture="transformer", virt_nodes=0).double().eval()
    reseed_parameters(m, 5)
    m = m.double()
    M, Dv = 50, 1088
    g = torch.Generator().manual_seed(2)
    feats = torch.randn(M, Dv, generator=g, dtype=torch.float64)
    x = torch.randn(M, 4, generator=g, dtype=torch.float64)
    t = torch.randint(0, 300, (M,), generator=g)
    W1, b1 = m.mlp[0].weight, m.mlp[0].bias
    with torch.no_grad():
        want = torch.nn.functional.gelu(m.mlp[0](torch.cat([feats, m.pos_mlp(x), m.time_emb(t)], -1)))
        P = feats @ W1[:, :Dv].T + b1
        Wpos, Wtime = W1[:, Dv:Dv + 32], W1[:, Dv + 32:Dv + 64]
        wc = Wpos @ m.pos_mlp[2].weight                                   # [Hm, 16]
        tt = m.time_emb.weight @ Wtime.T + Wpos @ m.pos_mlp[2].bias       # [T, Hm]
        hid = torch.nn.functional.gelu(m.pos_mlp[0](x))                   # [M, 16]
        got = torch.nn.functional.gelu(P + tt[t] + hid @ wc.T)
    assert float((got - want).abs().max() / want.abs().max()) < 1e-12
