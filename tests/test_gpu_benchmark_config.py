"""Parity at the BENCHMARKED sizes (BASELINE.json configs[2] / configs[4]) against the LIVE oracle.

`bench.py` measures 900-node Exphander-60 % graphs with 8 virtual nodes per graph through the tensor-core path
(`gemm_mode="bf16x3"`, `attn_mode="auto"`): dense 128 x 64 tiles with promoted virtual-node sources on extra bitmap
columns, hub rows with ~900 in-edges on the lane-per-edge CSR kernel, and -- with more than one graph in the batch --
the reference's cross-graph virtual wiring (exophormer_gnn.py:185-200).  These tests run exactly that configuration
(one and two graphs, so that the cross-graph wiring exists) and compare one fused DDIM step with the CPU oracle's
`p_sample` on the same inputs at the 1e-4 bar, at the first (t = 290) and last (t = 0) timestep of the schedule the
launch script ships (`singularity/gianscarpe/train_celeba_rot.sh:12,15`).  Weights come from `reseed_parameters`
(attention weights spanning 1e-4 .. 0.9), not from default initialisers whose attention is uniform.
"""
import numpy as np
import pytest
import torch

import oracle
from common import TOL, reseed_parameters, rel_err, synth_graph_batch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _pair(arch, V, gemm, attn, seed=11):
    import diffassemble_b200 as dab

    ref = oracle.GNNDiffusionRef(steps=300, sampling="DDIM", rotation=True, architecture=arch, virt_nodes=V,
                                 model_mean_type=oracle.ModelMeanType.START_X, inference_ratio=10).eval()
    reseed_parameters(ref, seed)
    mod = dab.GNN_Diffusion(steps=300, sampling="DDIM", rotation=True, architecture=arch, virt_nodes=V,
                            model_mean_type=dab.ModelMeanType.START_X, inference_ratio=10, gemm_mode=gemm, attn_mode=attn)
    mod.load_state_dict(ref.state_dict(), strict=True)
    return ref, mod.to(DEV)


_ORACLE_CACHE = {}


def _oracle_steps(key, ref, x, ei, feats, batch, ts):
    """Oracle outputs are shared between the (gemm, attn) parametrisations: one CPU evaluation per case."""
    if key not in _ORACLE_CACHE:
        torch.set_num_threads(max(1, torch.get_num_threads()))
        out = {}
        with torch.no_grad():
            for i in ts:
                t = torch.full((x.shape[0],), i, dtype=torch.long)
                out[i] = ref.p_sample(x, t, i, edge_index=ei, patch_feats=feats, batch=batch)[0]
        _ORACLE_CACHE[key] = out
    return _ORACLE_CACHE[key]


@pytest.mark.parametrize("gemm,attn", [("bf16x3", "auto"), ("fp32", "csr")])
@pytest.mark.parametrize("B", [1, 2])
def test_c3_exphander60_v8_900_nodes_vs_live_oracle(B, gemm, attn):
    ref, mod = _pair("exophormer", 8, gemm, attn)
    sizes = [900] * B
    ei, batch = synth_graph_batch(sizes, kind="expander", degree="60%", seed=40)
    M = sum(sizes)
    g = torch.Generator().manual_seed(5 + B)
    feats, x = torch.randn(M, 1088, generator=g), torch.randn(M, 4, generator=g)
    want = _oracle_steps(("exph", B), ref, x, ei, feats, batch, (290, 0))
    ei_d, b_d, f_d = ei.to(DEV), batch.to(DEV), feats.to(DEV)
    for i in (290, 0):
        t = torch.full((M,), i, dtype=torch.long, device=DEV)
        got, _ = mod.p_sample(x.to(DEV), t, i, cond=None, edge_index=ei_d, patch_feats=f_d, batch=b_d)
        assert rel_err(got, want[i]) < TOL, (B, gemm, attn, i, rel_err(got, want[i]))
    stats = mod.model._engine.graph_stats()
    if attn == "auto":
        # every in-graph edge AND the promoted virtual -> real edges ran on the tensor cores
        assert stats["dense_edges"] >= ei.shape[1] + M, stats
        assert stats["dense_graphs"] == B


@pytest.mark.parametrize("gemm,attn", [("bf16x3", "auto"), ("fp32", "csr")])
def test_c3_dense_900_nodes_vs_live_oracle(gemm, attn):
    """The fully connected 900-node puzzle (810 000 edges incl. self loops), transformer architecture (GELU between
    layers): the end of the density sweep of SURVEY.md section 8(d)."""
    ref, mod = _pair("transformer", 0, gemm, attn, seed=12)
    n = 900
    ei, batch = oracle.dense_edge_index(n), torch.zeros(n, dtype=torch.long)
    g = torch.Generator().manual_seed(9)
    feats, x = torch.randn(n, 1088, generator=g), torch.randn(n, 4, generator=g)
    want = _oracle_steps(("dense",), ref, x, ei, feats, batch, (290, 0))
    ei_d, b_d, f_d = ei.to(DEV), batch.to(DEV), feats.to(DEV)
    for i in (290, 0):
        t = torch.full((n,), i, dtype=torch.long, device=DEV)
        got, _ = mod.p_sample(x.to(DEV), t, i, cond=None, edge_index=ei_d, patch_feats=f_d, batch=b_d)
        assert rel_err(got, want[i]) < TOL, (gemm, attn, i, rel_err(got, want[i]))


@pytest.mark.parametrize("degree", ["20%", "40%"])
def test_c3_sparser_exphander_900_nodes_vs_live_oracle(degree):
    """d = 20 % / 40 % (the rest of the roofline sweep), V = 4, tensor-core path."""
    ref, mod = _pair("exophormer", 4, "bf16x3", "auto", seed=13)
    ei, batch = synth_graph_batch([900], kind="expander", degree=degree, seed=77)
    g = torch.Generator().manual_seed(3)
    feats, x = torch.randn(900, 1088, generator=g), torch.randn(900, 4, generator=g)
    want = _oracle_steps(("sparse", degree), ref, x, ei, feats, batch, (290,))
    t = torch.full((900,), 290, dtype=torch.long, device=DEV)
    got, _ = mod.p_sample(x.to(DEV), t, 290, cond=None, edge_index=ei.to(DEV), patch_feats=feats.to(DEV), batch=batch.to(DEV))
    assert rel_err(got, want[290]) < TOL, rel_err(got, want[290])


def test_c3_sampling_loop_900_nodes_end_of_trajectory():
    """Whole 30-step DDIM loop on one 900-node Exphander graph (V = 8): free-running CUDA trajectory against the
    free-running oracle trajectory from the same x_T (end-of-trajectory parity)."""
    ref, mod = _pair("exophormer", 8, "bf16x3", "auto", seed=14)
    ei, batch = synth_graph_batch([900], kind="expander", degree="60%", seed=41)
    g = torch.Generator().manual_seed(21)
    feats, x = torch.randn(900, 1088, generator=g), torch.randn(900, 4, generator=g)
    xo = x.clone()
    ei_d, b_d, f_d = ei.to(DEV), batch.to(DEV), feats.to(DEV)
    xg = x.to(DEV)
    worst = 0.0
    with torch.no_grad():
        for i in reversed(range(0, 300, 10)):
            t = torch.full((900,), i, dtype=torch.long)
            xo, _ = ref.p_sample(xo, t, i, edge_index=ei, patch_feats=feats, batch=batch)
            xg, _ = mod.p_sample(xg, t.to(DEV), i, cond=None, edge_index=ei_d, patch_feats=f_d, batch=b_d)
            worst = max(worst, rel_err(xg, xo))
    assert worst < TOL, worst


@pytest.mark.parametrize("gemm,attn", [("bf16x3", "auto"), ("fp32", "csr")])
def test_c5_training_gradients_at_144_node_graphs(gemm, attn):
    """configs[4] shape: 12 x 12 dense puzzles (2 x 144 nodes), loss + every parameter gradient against the oracle's
    autograd (Huber loss, per-graph timesteps, spatial_diffusion.py:432-483,707-722)."""
    import diffassemble_b200 as dab

    ref = oracle.GNNDiffusionRef(steps=300, sampling="DDIM", rotation=True, architecture="transformer", virt_nodes=0,
                                 model_mean_type=oracle.ModelMeanType.START_X, inference_ratio=10)
    reseed_parameters(ref, 15)
    mod = dab.GNN_Diffusion(steps=300, sampling="DDIM", rotation=True, architecture="transformer", virt_nodes=0,
                            model_mean_type=dab.ModelMeanType.START_X, inference_ratio=10, gemm_mode=gemm, attn_mode=attn)
    mod.load_state_dict(ref.state_dict(), strict=True)
    mod = mod.to(DEV)
    sizes = [144, 144]
    ei, batch = synth_graph_batch(sizes)
    M = sum(sizes)
    g = torch.Generator().manual_seed(31)
    feats = torch.randn(M, 1088, generator=g)
    x0 = torch.rand(M, 4, generator=g) * 2 - 1
    noise = torch.randn(M, 4, generator=g)
    t = torch.tensor([37, 251])[batch]
    ref.train()
    loss_ref = ref.p_losses(x0, t, noise=noise, loss_type="huber", edge_index=ei, patch_feats=feats, batch=batch)
    loss_ref.backward()
    loss = mod.p_losses(x0.to(DEV), t.to(DEV), noise=noise.to(DEV), loss_type="huber", cond=feats.to(DEV),
                        edge_index=ei.to(DEV), batch=batch.to(DEV))
    loss.backward()
    assert abs(loss.item() - loss_ref.item()) < TOL * max(1.0, abs(loss_ref.item()))
    ref_grads = {k: p.grad for k, p in ref.named_parameters() if p.grad is not None}
    n_checked = 0
    worst = (0.0, "")
    for k, p in mod.named_parameters():
        if k not in ref_grads or p.grad is None:
            continue
        gr = ref_grads[k]
        if gr.abs().max() < 1e-9:   # lin_key.bias: softmax is shift-invariant, the true gradient is exactly 0
            assert p.grad.abs().max() < 1e-9, k
            continue
        e = rel_err(p.grad, gr)
        worst = max(worst, (e, k))
        n_checked += 1
        # gradients are long cancelling sums (over 288 nodes x 20 736 edges): 1e-3 relative of the tensor's max
        assert e < 1e-3, (k, e)
    assert n_checked >= 38, n_checked
    print(f"gradient parity at 2 x 144 nodes [{gemm}/{attn}]: {n_checked} tensors, worst {worst[0]:.2e} ({worst[1]})")


def test_folded_path_runs_at_c3_and_matches_the_unfolded_pipeline(monkeypatch):
    """csrc/fold.cu: at the benchmarked configuration the steps take the weight-folded path (first projection from the
    128-wide trunk hidden incl. the one-hot columns of the virtual rows, last layer aggregated on 32-channel values);
    `DA_NO_FOLD=1` keeps the unfolded pipeline.  Both must agree with the live oracle, and with each other far inside
    the parity bar."""
    ei, batch = synth_graph_batch([900, 900], kind="expander", degree="60%", seed=40)
    M = 1800
    g = torch.Generator().manual_seed(7)
    feats, x = torch.randn(M, 1088, generator=g), torch.randn(M, 4, generator=g)
    outs = {}
    for no_fold, persist in (("0", "1"), ("0", "0"), ("1", "1")):
        monkeypatch.setenv("DA_NO_FOLD", no_fold)
        monkeypatch.setenv("DA_FOLD_PERSIST", persist)   # "0": one (tile, head) item per CTA instead of the persistent kernel
        ref, mod = _pair("exophormer", 8, "bf16x3", "auto")
        t = torch.full((M,), 290, dtype=torch.long, device=DEV)
        got, _ = mod.p_sample(x.to(DEV), t, 290, cond=None, edge_index=ei.to(DEV), patch_feats=feats.to(DEV), batch=batch.to(DEV))
        info = mod.model._engine.plan_info()
        assert info["real_rows_clean"] == 1, info
        assert info["folded"] == (1 if no_fold == "0" else 0), info
        if no_fold == "0" and persist == "0":
            assert rel_err(got.cpu(), outs["0"]) < 1e-6, rel_err(got.cpu(), outs["0"])   # same arithmetic, different scheduling
            continue
        outs[no_fold] = got.cpu()
    want = _oracle_steps(("exph", 2, "fold"), ref, x, ei, feats, batch, (290,))[290]
    assert rel_err(outs["0"], want) < TOL, rel_err(outs["0"], want)
    assert rel_err(outs["1"], want) < TOL, rel_err(outs["1"], want)
    assert rel_err(outs["0"], outs["1"]) < 2e-5, rel_err(outs["0"], outs["1"])


def test_folded_path_small_mixed_batches():
    """The folded path is per batch: dense-tile graphs whose rows are all clean take it (ragged sizes, non-multiple-of-128
    tails, transformer architecture without virtual rows); a batch with a graph too small for the tiles does not."""
    for arch, V, sizes, expect in (("transformer", 0, [144, 100, 64], 1), ("exophormer", 4, [200, 130], 1),
                                   ("exophormer", 4, [200, 20], 0)):
        ref, mod = _pair(arch, V, "bf16x3", "auto", seed=19)
        ei, batch = synth_graph_batch(sizes, kind="expander", degree="60%", seed=3)
        M = sum(sizes)
        g = torch.Generator().manual_seed(2)
        feats, x = torch.randn(M, 1088, generator=g), torch.randn(M, 4, generator=g)
        with torch.no_grad():
            t = torch.full((M,), 150, dtype=torch.long)
            want = ref.p_sample(x, t, 150, edge_index=ei, patch_feats=feats, batch=batch)[0]
        got, _ = mod.p_sample(x.to(DEV), t.to(DEV), 150, cond=None, edge_index=ei.to(DEV), patch_feats=feats.to(DEV), batch=batch.to(DEV))
        info = mod.model._engine.plan_info()
        assert info["folded"] == expect, (arch, sizes, info)
        assert rel_err(got, want) < TOL, (arch, sizes, rel_err(got, want))


@pytest.mark.parametrize("arch,V,sizes", [("exophormer", 8, [900, 900, 900]), ("transformer", 0, [900, 300, 144, 64]),
                                          ("exophormer", 4, [130] * 40)])
def test_persistent_hidden_layer_kernel_equals_the_per_tile_kernel(monkeypatch, arch, V, sizes):
    """csrc/attn_hidden.cu: the 32-channel hidden layers run on ONE persistent CTA per SM with two independent streams of
    (tile, head) items whenever every real row is finalised on the tensor cores.  Same scores, same masked online softmax,
    same split-bf16 products as `attn_dense_kernel<32, 4>`; only the P V accumulation is grouped differently (P_hi [V_hi |
    V_lo] as one N = 64 instruction, two accumulators summed at the end), so a free-running 3-step DDIM trajectory must
    agree with `DA_HIDDEN_PERSIST=0` (per-(tile, head) CTAs) to fp32 rounding -- with full and partial tiles, more items
    than streams (3 x 8 x 8 and 40 x 2 x 8 items), fewer items than streams, GELU between the layers (transformer) and
    none (exophormer) -- and both with the live oracle."""
    ei, batch = synth_graph_batch(sizes, kind="expander", degree="60%", seed=21)
    M = sum(sizes)
    g = torch.Generator().manual_seed(3)
    feats, x = torch.randn(M, 1088, generator=g), torch.randn(M, 4, generator=g)
    outs = {}
    for persist in ("1", "0"):
        monkeypatch.setenv("DA_HIDDEN_PERSIST", persist)
        ref, mod = _pair(arch, V, "bf16x3", "auto", seed=23)
        xt = x.to(DEV)
        traj = []
        for i in (290, 280, 270):
            t = torch.full((M,), i, dtype=torch.long, device=DEV)
            xt, _ = mod.p_sample(xt, t, i, cond=None, edge_index=ei.to(DEV), patch_feats=feats.to(DEV), batch=batch.to(DEV))
            traj.append(xt.cpu())
        info = mod.model._engine.plan_info()
        assert info["real_rows_clean"] == 1, info
        assert info["persistent_hidden_launches"] == (9 if persist == "1" else 0), info
        outs[persist] = traj
    for a, b in zip(outs["1"], outs["0"]):
        assert rel_err(a, b) < 2e-6, (arch, sizes, rel_err(a, b))
    with torch.no_grad():
        t = torch.full((M,), 290, dtype=torch.long)
        want = ref.p_sample(x, t, 290, edge_index=ei, patch_feats=feats, batch=batch)[0]
    assert rel_err(outs["1"][0], want) < TOL, rel_err(outs["1"][0], want)


@pytest.mark.parametrize("switch", ["DA_NO_PDL=1", "DA_HIDDEN_NOMAX=0", "DA_NO_PRO_TABLE=1", "DA_NO_GATHER_RIDE=1", "DA_FOLD_PERSIST=0"])
def test_development_switches_change_scheduling_not_results(monkeypatch, switch):
    """Every A/B switch of DESIGN.md section 4 selects another schedule or another grouping of the same arithmetic:
    plain stream order instead of programmatic dependent launch (bit-identical), the running maximum in the persistent
    kernel's score loop (another reference point of the online softmax), the literal prologue instead of its table form
    (fp32 re-association), separate gather launches (bit-identical), per-(tile, head) CTAs for the folded last layer.  Two
    free-running DDIM steps on the benchmarked configuration (2 x 900 nodes, Exphander 60 %, 8 virtual nodes) must agree
    with the default path to fp32 rounding."""
    ei, batch = synth_graph_batch([900, 900], kind="expander", degree="60%", seed=40)
    M = 1800
    g = torch.Generator().manual_seed(7)
    feats, x = torch.randn(M, 1088, generator=g), torch.randn(M, 4, generator=g)
    outs = []
    for env in (None, switch):
        if env is not None:
            k, v = env.split("=")
            monkeypatch.setenv(k, v)
        ref, mod = _pair("exophormer", 8, "bf16x3", "auto")
        xt = x.to(DEV)
        for i in (290, 280):
            t = torch.full((M,), i, dtype=torch.long, device=DEV)
            xt, _ = mod.p_sample(xt, t, i, cond=None, edge_index=ei.to(DEV), patch_feats=feats.to(DEV), batch=batch.to(DEV))
        outs.append(xt.cpu())
    exact = switch in ("DA_NO_PDL=1", "DA_NO_GATHER_RIDE=1")
    if exact:
        assert torch.equal(outs[0], outs[1]), (switch, rel_err(outs[0], outs[1]))
    else:
        assert rel_err(outs[0], outs[1]) < 3e-6, (switch, rel_err(outs[0], outs[1]))
