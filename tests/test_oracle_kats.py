"""Known-answer tests that anchor the CPU oracle -- in particular the two un-vendored third-party pieces
(``TransformerConv``, pytorch3d's quaternion conversions) that ``test_oracle_pinned.py`` cannot pin.

Each test states a property of the reference formulation that must
hold independently of weights."""
import math

import pytest
import torch
import torch.nn.functional as F

import oracle
from oracle import so3
from oracle.transformer_conv import TransformerConvRef, segment_softmax


def _conv(in_c=16, c=4, h=2, seed=0):
    torch.manual_seed(seed)
    return TransformerConvRef(in_c, c, h).double()


def test_dense_graph_equals_sdpa():
    conv = _conv(32, 8, 4)
    n = 11
    x = torch.randn(n, 32, dtype=torch.float64)
    ei = oracle.dense_edge_index(n)
    y = conv(x, ei)
    q = conv.lin_query(x).view(n, 4, 8).transpose(0, 1)
    k = conv.lin_key(x).view(n, 4, 8).transpose(0, 1)
    v = conv.lin_value(x).view(n, 4, 8).transpose(0, 1)
    ref = F.scaled_dot_product_attention(q, k, v).transpose(0, 1).reshape(n, 32) + conv.lin_skip(x)
    assert torch.allclose(y, ref, atol=1e-12)


def test_dense_edge_index_is_row_major_with_self_loops():
    ei = oracle.dense_edge_index(3)
    assert ei.tolist() == [[0, 0, 0, 1, 1, 1, 2, 2, 2], [0, 1, 2, 0, 1, 2, 0, 1, 2]]


def test_uniform_attention_when_query_is_zero():
    conv = _conv()
    with torch.no_grad():
        conv.lin_query.weight.zero_()
        conv.lin_query.bias.zero_()
    x = torch.randn(6, 16, dtype=torch.float64)
    ei = torch.tensor([[0, 1, 2, 3, 4], [5, 5, 5, 0, 0]])
    y, (_, alpha) = conv(x, ei, return_attention_weights=True)
    v = conv.lin_value(x)
    assert torch.allclose(y[5], v[[0, 1, 2]].mean(0) + conv.lin_skip(x)[5], atol=1e-12)
    assert torch.allclose(alpha[:3], torch.full((3, 2), 1 / 3, dtype=torch.float64))


def test_single_in_edge_has_alpha_one_and_isolated_node_is_skip_only():
    conv = _conv()
    x = torch.randn(4, 16, dtype=torch.float64)
    ei = torch.tensor([[2], [1]])
    y, (_, alpha) = conv(x, ei, return_attention_weights=True)
    assert torch.allclose(alpha, torch.ones(1, 2, dtype=torch.float64))
    assert torch.allclose(y[1], conv.lin_value(x)[2] + conv.lin_skip(x)[1], atol=1e-12)
    for i in (0, 2, 3):  # no in-edges -> aggregate is exactly zero
        assert torch.equal(y[i], conv.lin_skip(x)[i])


def test_duplicate_edge_counts_twice():
    conv = _conv()
    x = torch.randn(3, 16, dtype=torch.float64)
    ei = torch.tensor([[0, 0, 1], [2, 2, 2]])
    _, (_, alpha) = conv(x, ei, return_attention_weights=True)
    assert torch.allclose(alpha[0], alpha[1])
    q = conv.lin_query(x).view(3, 2, 4)
    k = conv.lin_key(x).view(3, 2, 4)
    s0 = (q[2] * k[0]).sum(-1) / 2.0
    s1 = (q[2] * k[1]).sum(-1) / 2.0
    expect = torch.exp(s0) / (2 * torch.exp(s0) + torch.exp(s1))
    assert torch.allclose(alpha[0], expect, atol=1e-12)


def test_segment_softmax_denominator_eps():
    src = torch.tensor([[0.0], [0.0]])
    out = segment_softmax(src, torch.tensor([0, 0]), 1)
    assert torch.allclose(out, torch.full((2, 1), 0.5))


def test_exophormer_wiring_counts_match_survey():
    # n=900, V=4, B=1 -> 900 + (900+4)*4 extra edges, 2716 duplicate virtual self loops
    from oracle.gnn import exophormer_wiring

    n, V = 900, 4
    ei = torch.zeros((2, 0), dtype=torch.long)
    _, batch_ext, ext = exophormer_wiring(ei, torch.zeros(n, dtype=torch.long), V)
    assert ext.shape[1] == n + (n + V) * V
    assert len(batch_ext) == n + V
    virt_self = ((ext[0] >= n) & (ext[1] >= n)).sum().item()
    assert virt_self == (n + V) * V - n
    # every real node: exactly one edge to a virtual node and one from a virtual node
    assert torch.equal(torch.bincount(ext[0][ext[0] < n], minlength=n), torch.ones(n, dtype=torch.long))
    assert torch.equal(torch.bincount(ext[1][ext[1] < n], minlength=n), torch.ones(n, dtype=torch.long))


def test_exophormer_wiring_crosses_graphs_when_batched():
    from oracle.gnn import exophormer_wiring

    batch = torch.arange(4).repeat_interleave(50)
    _, _, ext = exophormer_wiring(torch.zeros((2, 0), dtype=torch.long), batch, 4)
    src, dst = ext
    real_to_virt = src < 200
    graph_of_virt = (dst[real_to_virt] - 200) // 4
    assert (graph_of_virt != batch[src[real_to_virt]]).any()


def _diffusion(sampling, mean_type, ratio=1, T=50):
    torch.manual_seed(0)
    return oracle.GNNDiffusionRef(steps=T, sampling=sampling, rotation=True, inference_ratio=ratio,
                                  model_mean_type=oracle.ModelMeanType[mean_type]).eval()


def test_schedule_buffers():
    m = _diffusion("DDPM", "EPSILON", T=300)
    assert m.betas[0].item() == pytest.approx(1e-4) and m.betas[-1].item() == pytest.approx(0.02)
    assert torch.allclose(m.alphas_cumprod, torch.cumprod(1 - m.betas, 0))
    assert m.alphas_cumprod_prev[0] == 1.0
    assert m.posterior_variance[0] == 0.0


def test_ddpm_step_with_zero_model_output():
    m = _diffusion("DDPM", "EPSILON")
    with torch.no_grad():
        for p in m.model.final_mlp[2].parameters():
            p.zero_()
    n = 9
    ei = oracle.dense_edge_index(n)
    x = torch.randn(n, 4)
    feats = torch.randn(n, 1088)
    t = torch.full((n,), 0, dtype=torch.long)
    y, _ = m.p_sample(x, t, 0, edge_index=ei, patch_feats=feats, batch=torch.zeros(n, dtype=torch.long))
    assert torch.allclose(y, m.sqrt_recip_alphas[0] * x)
    t = torch.full((n,), 7, dtype=torch.long)
    noise = torch.randn(n, 4)
    y, _ = m.p_sample(x, t, 7, edge_index=ei, patch_feats=feats, batch=torch.zeros(n, dtype=torch.long), noise=noise)
    assert torch.allclose(y, m.sqrt_recip_alphas[7] * x + m.posterior_variance[7].sqrt() * noise)


def test_ddim_last_step_returns_x0():
    m = _diffusion("DDIM", "START_X", ratio=10)
    n = 9
    ei = oracle.dense_edge_index(n)
    x, feats = torch.randn(n, 4), torch.randn(n, 1088)
    t = torch.zeros(n, dtype=torch.long)
    b = torch.zeros(n, dtype=torch.long)
    y, _ = m.p_sample(x, t, 0, edge_index=ei, patch_feats=feats, batch=b)
    x0 = m.forward_with_feats(x, t, None, ei, feats, b)
    assert torch.allclose(y, x0, atol=1e-6)


def test_so3_kats():
    q = so3.matrix_to_quaternion(so3.skew_to_rmat(torch.zeros(2, 3)))
    assert torch.allclose(q, torch.tensor([[1.0, 0, 0, 0]]).repeat(2, 1))
    r = torch.tensor([[0.3, -0.2, 0.5], [1.0, 2.0, -0.5]], dtype=torch.float64)
    R = so3.skew_to_rmat(r)
    assert torch.allclose(R @ R.transpose(-1, -2), torch.eye(3, dtype=torch.float64).expand(2, 3, 3), atol=1e-12)
    assert torch.allclose(so3.skew2vec(so3.log_rmat(R)), r, atol=1e-10)
    assert torch.allclose(so3.quaternion_to_matrix(so3.matrix_to_quaternion(R)), R, atol=1e-12)
    half = so3.so3_scale(R, torch.tensor([0.5, 0.5], dtype=torch.float64))
    assert torch.allclose(half @ half, R, atol=1e-10)
    # rotation about z by 90 degrees -> q = (cos 45, 0, 0, sin 45)
    qz = so3.matrix_to_quaternion(so3.skew_to_rmat(torch.tensor([[0.0, 0.0, math.pi / 2]], dtype=torch.float64)))
    assert torch.allclose(qz, torch.tensor([[math.sqrt(0.5), 0, 0, math.sqrt(0.5)]], dtype=torch.float64), atol=1e-12)


def test_expander_is_regular_symmetric_and_seeded():
    import numpy as np

    e = oracle.generate_random_expander(30, "60%", rng=np.random.default_rng(3), check_spectral_gap=True)
    d = round(60 * 29 / 100)
    assert e.shape == (30 * d, 2)
    deg = torch.bincount(e[:, 1], minlength=30)
    assert (deg == d).all()
    pairs = set(map(tuple, e.tolist()))
    assert all((b, a) in pairs for a, b in pairs) and all(a != b for a, b in pairs)
    e2 = oracle.generate_random_expander(30, "60%", rng=np.random.default_rng(3), check_spectral_gap=True)
    assert torch.equal(e, e2)
    small = oracle.generate_random_expander(6, 4)
    assert small.shape == (30, 2)


def test_greedy_assignment_kats():
    # pieces exactly on the grid: the assignment is the inverse permutation, all distances 0
    side = 5
    y = torch.linspace(-1, 1, side); x = torch.linspace(-1, 1, side)
    grid = torch.stack(torch.meshgrid(x, y, indexing="xy"), -1).reshape(-1, 2)
    perm = torch.randperm(25, generator=torch.Generator().manual_seed(0))
    for flag in (False, True):
        a = oracle.greedy_cost_assignment_ref(grid[perm], grid, separately_rounded=flag)
        assert a.shape == (25, 3) and (a[:, 2] == 0).all()
        order = a[torch.sort(a[:, 0])[1]]
        assert torch.equal(order[:, 1], perm)
    # greedy, not optimal: the globally closest pair is taken first
    p1 = torch.tensor([[0.0, 0.0], [1.0, 0.0]]); p2 = torch.tensor([[0.9, 0.0], [5.0, 0.0]])
    a = oracle.greedy_cost_assignment_ref(p1, p2)
    assert a[:, :2].tolist() == [[1, 0], [0, 1]]
