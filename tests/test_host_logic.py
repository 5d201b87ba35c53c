"""CPU-only tests of the host mirror: API surface, wiring, topology, schedule scalars, the
C-ABI library's exported symbols, and loud failure without a GPU."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest
import torch

import diffassemble_b200 as dab
import oracle
from diffassemble_b200 import _cabi, sharding, topology

ROOT = Path(__file__).resolve().parents[1]


def test_library_exports_every_declared_symbol(lib_built):
    header = (ROOT / "include" / "diffassemble_b200.h").read_text()
    declared = set(re.findall(r"\b(da_[a-z_0-9]+)\s*\(", header))
    declared -= {"da_handle", "da_config", "da_status"}
    assert declared == set(_cabi.EXPORTED_SYMBOLS)
    lib = ctypes.CDLL(str(lib_built))
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert _cabi.load_library().da_abi_version() == _cabi.DA_ABI_VERSION


def test_struct_layouts_match_header():
    assert ctypes.sizeof(_cabi.da_config) == 4 * (15 + 8)
    assert ctypes.sizeof(_cabi.da_step_coef) == 4 * 14
    assert ctypes.sizeof(_cabi.da_weight_desc) == 32


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_fails_loudly_without_gpu(lib_built):
    lib = _cabi.load_library()
    cfg = _cabi.da_config()
    cfg.abi_version = _cabi.DA_ABI_VERSION
    h = ctypes.c_void_p(0)
    st = lib.da_create(ctypes.byref(h), ctypes.byref(cfg))
    assert st == _cabi.DA_ERR_CUDA and not h.value
    assert b"no CPU fallback" in lib.da_last_error(None)
    m = dab.GNN_Diffusion(steps=10, rotation=True)
    n = 4
    with pytest.raises(RuntimeError, match="no CPU path"):
        m.forward_with_feats(torch.zeros(n, 4), torch.zeros(n, dtype=torch.long), None, oracle.dense_edge_index(n),
                             torch.zeros(n, 1088), torch.zeros(n, dtype=torch.long))


def test_product_never_imports_oracle():
    for py in (ROOT / "diffassemble_b200").rglob("*.py"):
        src = py.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), py
    for src in (ROOT / "diffassemble_b200" / "csrc").iterdir():
        assert "oracle" not in src.read_text(), src


@pytest.mark.parametrize("arch,V", [("transformer", 0), ("exophormer", 4), ("exophormer", 8)])
def test_state_dict_keys_match_reference_layout(arch, V):
    ref = oracle.GNNDiffusionRef(steps=20, rotation=True, architecture=arch, virt_nodes=V)
    mod = dab.GNN_Diffusion(steps=20, rotation=True, architecture=arch, virt_nodes=V)
    assert set(ref.state_dict()) == set(mod.state_dict())
    for k, v in ref.state_dict().items():
        assert mod.state_dict()[k].shape == v.shape, k
    # the key names the reference checkpoints use
    keys = set(mod.state_dict())
    for k in ["model.time_emb.weight", "model.pos_mlp.0.weight", "model.mlp.2.bias", "model.final_mlp.0.weight",
              "model.gnn_backbone.module_list.3.lin_skip.weight", "model.linear1.weight", "model.mean", "betas",
              "sqrt_recipm1_alphas_cumprod", "posterior_variance"]:
        assert k in keys, k
    if V:
        assert mod.state_dict()["model.gnn_backbone.virt_node_embedding.weight"].shape == (V, 1152)


def test_state_dict_3d():
    mod = dab.GNN_Diffusion_3d(steps=20, sampling="DDIM", backbone="pointnet")
    sd = mod.state_dict()
    assert sd["model.mlp_t.2.weight"].shape == (3, 256) and sd["model.mlp.0.weight"].shape == (256, 192)
    assert sd["model.gnn_backbone.module_list.3.lin_key.weight"].shape == (192, 256)


@pytest.mark.parametrize("sizes,V", [([5, 7, 3], 4), ([12], 8), ([4, 4, 4, 4], 2)])
def test_extend_graph_equals_reference_wiring(sizes, V):
    from oracle.gnn import exophormer_wiring

    ei, batch = oracle.batch_graphs([oracle.dense_edge_index(n) for n in sizes], sizes)
    vid, _, ext = exophormer_wiring(ei, batch, V)
    gnn = dab.Exophormer_GNN(64, 256, 8, 64, 4, virt_nodes=V)
    ext2, num_total, vid2 = gnn.extend_graph(ei, batch)
    assert torch.equal(ext, ext2)
    assert num_total == sum(sizes) + V * len(sizes)
    assert torch.equal(vid.to(torch.int32), vid2)


def test_topology_matches_oracle():
    assert torch.equal(topology.dense_edge_index(7), oracle.dense_edge_index(7))
    for n, deg, seed in [(30, "60%", 1), (64, 20, 2), (25, 6, 3), (900, "60%", 4), (6, 3, 5)]:
        a = topology.expander_edge_index(n, deg, rng=np.random.default_rng(seed))
        b = oracle.generate_random_expander(n, deg, rng=np.random.default_rng(seed), check_spectral_gap=False).t()
        assert torch.equal(a, b), (n, deg)
    # With the spectral-gap retry on, every candidate is a relabelled circulant graph with the SAME
    # spectrum, so the reference's "keep the best lambda_2 of 5 draws" is decided by float32 ARPACK
    # noise.  The faithful statement is: the result is one of the 5 draws of the seeded generator.
    a = topology.expander_edge_index(40, "40%", rng=np.random.default_rng(9), check_spectral_gap=True)
    rng = np.random.default_rng(9)
    draws = [np.stack(oracle.generate_random_regular_graph(40, 16, rng)) for _ in range(5)]
    assert any(np.array_equal(a.numpy(), d) for d in draws)
    assert topology.resolve_degree(900, "60%") == 539


def test_step_coefficients_follow_reference_buffers():
    mod = dab.GNN_Diffusion(steps=300, sampling="DDIM", inference_ratio=10, rotation=True,
                            model_mean_type=dab.ModelMeanType.START_X)
    c = mod._step_coef(290, mod._pred_code())
    assert c.has_prev == 1 and c.pred == _cabi.DA_PRED_START_X and c.eta == 0.0
    assert c.acp == pytest.approx(mod.alphas_cumprod[290].item(), rel=0, abs=0)
    assert c.acp_prev == mod.alphas_cumprod[280].item()
    c0 = mod._step_coef(0, mod._pred_code())
    assert c0.has_prev == 0 and c0.acp_prev == 1.0
    ref = oracle.GNNDiffusionRef(steps=300)
    for name in ["betas", "alphas_cumprod", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod",
                 "posterior_variance", "sqrt_one_minus_alphas_cumprod", "sqrt_recip_alphas"]:
        assert torch.equal(getattr(mod, name), getattr(ref, name)), name


def test_constructor_surface():
    import inspect

    sig = inspect.signature(dab.GNN_Diffusion.__init__).parameters
    for kw in ["steps", "inference_ratio", "sampling", "learning_rate", "save_and_sample_every", "bb",
               "classifier_free_prob", "classifier_free_w", "noise_weight", "rotation", "model_mean_type",
               "input_channels", "output_channels", "scheduler", "visual_pretrained", "freeze_backbone", "backbone",
               "n_layers", "architecture", "virt_nodes", "all_equivariant"]:
        assert kw in sig, kw
    m = dab.GNN_Diffusion(steps=30, sampling="DDIM")
    for meth in ["forward", "forward_with_feats", "visual_features", "q_sample", "p_losses", "p_sample",
                 "p_sample_ddpm", "p_sample_ddim", "p_sample_loop", "sample", "prediction_step", "predict_step",
                 "validation_step", "test_step", "configure_optimizers", "initialize_torchmetrics"]:
        assert callable(getattr(m, meth)), meth
    with pytest.raises(NotImplementedError):
        dab.GNN_Diffusion(steps=30, architecture="gcn")


def test_shard_bounds_and_shard_batch():
    assert [sharding.shard_bounds(32, 8, r) for r in range(8)] == [(4 * r, 4 * r + 4) for r in range(8)]
    assert [sharding.shard_bounds(5, 2, r) for r in range(2)] == [(0, 3), (3, 5)]
    sizes = [3, 5, 2, 4]
    ei, batch = oracle.batch_graphs([oracle.dense_edge_index(n) for n in sizes], sizes)
    x = torch.arange(14.0)[:, None]
    got = []
    for r in range(2):
        ei_r, b_r, (x_r,), (n0, n1) = sharding.shard_batch(ei, batch, [x], 2, r)
        assert ei_r.min() == 0 and ei_r.max() == n1 - n0 - 1 and b_r.min() == 0
        got.append(x_r)
    assert torch.equal(torch.cat(got), x)
    bad = torch.tensor([[0], [13]])
    with pytest.raises(ValueError):
        sharding.shard_batch(torch.cat([ei, bad], 1), batch, [x], 2, 1)


def test_adafactor_descriptor_layout_matches_header():
    """``da_adafactor_param`` is filled as a numpy record (training.FusedAdafactor): 6 pointers + 2 int32 + 2 float."""
    import numpy as np

    from diffassemble_b200.training import FusedAdafactor

    opt = FusedAdafactor([torch.nn.Parameter(torch.zeros(3, 4))])
    assert opt._DESC.itemsize == 64
    assert [opt._DESC.fields[k][1] for k in ("p", "g", "sq_row", "sq_col", "sq", "rms", "rows", "cols", "beta2t", "rel_step")] == \
        [0, 8, 16, 24, 32, 40, 48, 52, 56, 60]
    header = (ROOT / "include" / "diffassemble_b200.h").read_text()
    body = header[header.index("typedef struct da_adafactor_param {"):header.index("} da_adafactor_param;")]
    order = re.findall(r"\b(p|g|sq_row|sq_col|sq|rms_out|rows, cols|beta2t, rel_step);", body)
    assert order == ["p", "g", "sq_row", "sq_col", "sq", "rms_out", "rows, cols", "beta2t, rel_step"]
    assert np.dtype(opt._DESC).alignment <= 8


def test_fused_adafactor_falls_back_to_stock_path_on_cpu_parameters():
    """Parameters the kernel does not cover (here: CPU tensors) take transformers' own code inside the same step,
    so the optimizer is a drop-in even for modules that keep some state on the host."""
    from transformers.optimization import Adafactor

    from diffassemble_b200.training import FusedAdafactor

    g = torch.Generator().manual_seed(0)
    base = [torch.randn(6, 5, generator=g), torch.randn(7, generator=g)]
    a = [torch.nn.Parameter(t.clone()) for t in base]
    b = [torch.nn.Parameter(t.clone()) for t in base]
    oa, ob = Adafactor(a), FusedAdafactor(b)
    for _ in range(3):
        for pa, pb in zip(a, b):
            gr = torch.randn(pa.shape, generator=g)
            pa.grad, pb.grad = gr.clone(), gr.clone()
        oa.step(); ob.step()
    for pa, pb in zip(a, b):
        assert torch.equal(pa, pb)
        assert pb.grad is not None   # gradients are left in place


def test_pointnet_mirror_has_reference_state_dict_and_no_cpu_path():
    enc = dab.PointNet(128)
    ref = oracle.PointNetRef(128)
    assert {k: tuple(v.shape) for k, v in enc.state_dict().items()} == {k: tuple(v.shape) for k, v in ref.state_dict().items()}
    with pytest.raises(RuntimeError, match="no CPU path"):
        enc.eval()(torch.zeros(1, 4, 3))
    with pytest.raises(NotImplementedError):
        enc.train()(torch.zeros(1, 4, 3))


def test_prefetch_requires_feature_matrix():
    m = dab.GNN_Diffusion(steps=10, rotation=True)
    with pytest.raises(ValueError, match="pre-computed node features"):
        m.prefetch(torch.zeros(4, 3, 32, 32), oracle.dense_edge_index(4), torch.zeros(4, dtype=torch.long))


def test_mirror_schedule_buffers_match_reference_modules():
    """a1: every registered length-T buffer of the mirror modules (2-D and 3-D, three schedulers, three T) equals the
    buffer of the REFERENCE module (tests/golden/ref_schedules.pt, written by executing the reference)."""
    from pathlib import Path

    d = torch.load(Path(__file__).resolve().parent / "golden" / "ref_schedules.pt")
    checked = 0
    for key, bufs in d.items():
        tag, sch, T = key.split("/")
        if tag == "2d":
            m = dab.GNN_Diffusion(steps=int(T), scheduler=dab.ModelScheduler[sch], rotation=True)
        else:
            m = dab.GNN_Diffusion_3d(steps=int(T), scheduler=dab.ModelScheduler[sch], backbone="pointnet", sampling="DDIM")
        mine = dict(m.named_buffers())
        for name, want in bufs.items():
            assert name in mine, (key, name)
            assert torch.allclose(mine[name], want, rtol=1e-6, atol=0), (key, name)
            checked += 1
    assert checked >= 100


@pytest.mark.parametrize("name,V", [("exph_2x64_v4_ddim", 4), ("exph_ragged_v8_ddim", 8), ("exph_v0_eps_ddim", 0)])
def test_mirror_virtual_wiring_equals_reference_run(name, V):
    """a10: ``Exophormer_GNN.extend_graph`` (host, once per batch) reproduces, edge for edge and in order, the edge list
    the REFERENCE's ``Exophormer_GNN.forward`` handed to its last TransformerConv (returned with the attention weights
    and stored in the fixture by tests/golden/make_reference_golden.py)."""
    from pathlib import Path

    d = torch.load(Path(__file__).resolve().parent / "golden" / f"ref_{name}.pt")
    gnn = dab.Exophormer_GNN(1152, hidden_dim=256, heads=8, output_size=1152, n_layers=4, virt_nodes=V)
    ext, num_total, virt_ids = gnn.extend_graph(d["edge_index"], d["batch"])
    assert torch.equal(ext, d["alpha_edge_index"])
    n_graphs = int(d["batch"].max()) + 1
    assert num_total == len(d["batch"]) + V * n_graphs
    if V > 0:
        assert torch.equal(virt_ids.long(), torch.arange(V).repeat(n_graphs))


def test_mirror_topology_equals_reference_generator_output():
    """a14: the host mirror of ``generate_random_expander`` against edge lists the REFERENCE function produced
    (tests/golden/ref_topology.pt): the single-attempt output exactly; the 5-attempt output (whose winner the reference
    itself picks by ARPACK noise) as a member of the candidate set."""
    from pathlib import Path

    d = torch.load(Path(__file__).resolve().parent / "golden" / "ref_topology.pt")
    for key, want in d.items():
        kind, n, deg, seed = key.split("/")
        n, seed = int(n), int(seed)
        deg = deg if deg.endswith("%") else int(deg)
        if kind == "expander1":
            got = topology.expander_edge_index(n, deg, rng=np.random.default_rng(seed))   # [2, E]
            assert torch.equal(got.t(), want), key
        else:
            dnum = topology.resolve_degree(n, deg)
            rng = np.random.default_rng(seed)
            cands = [torch.as_tensor(np.stack(topology.random_regular_edges(n, dnum, rng), 1)) for _ in range(5)]
            assert any(torch.equal(c, want) for c in cands), key


def test_mirror_state_dict_layout_equals_reference_checkpoint_layout():
    """b: a checkpoint written by the reference loads into the mirror: every key of the reference's ``state_dict``
    (tests/golden/ref_state_dict_keys.pt, from the reference modules themselves) exists in the mirror with the same
    shape -- for the 3-D module including the PointNet encoder and the ``identity`` buffer -- and the mirror adds
    nothing of its own.  (The 2-D visual encoder is an ``nn.Identity`` in the fixture: timm is absent.)"""
    from pathlib import Path

    d = torch.load(Path(__file__).resolve().parent / "golden" / "ref_state_dict_keys.pt")
    for key, want in d.items():
        parts = key.split("/")
        if parts[0] == "2d":
            if parts[1] == "norot":
                m = dab.GNN_Diffusion(steps=30, rotation=False, architecture="transformer")
            else:
                m = dab.GNN_Diffusion(steps=30, rotation=True, architecture=parts[1], virt_nodes=int(parts[2]))
        else:
            m = dab.GNN_Diffusion_3d(steps=30, sampling="DDIM", backbone="pointnet", architecture=parts[1])
        # (the fixture was read from the reference with timm replaced by a placeholder, so it has no visual_backbone
        # entries; the encoder's timm-style keys are held to the oracle's, which are pinned against torchvision's
        # implementation by a key map: tests/test_oracle_efficientnet.py)
        mine = {k: tuple(v.shape) for k, v in m.state_dict().items() if not k.startswith("model.visual_backbone.")}
        assert mine == want, (key, sorted(set(mine) ^ set(want))[:8])


def test_mirror_signatures_cover_reference_signatures():
    """b: every parameter of the reference's constructors / public methods on the path (tests/golden/ref_signatures.pt,
    read from the reference classes by ``inspect``) exists in the mirror with the same default; the mirror only adds
    keyword arguments of its own (``gemm_mode``, ``attn_mode``, ``noise=``, ``generator=``)."""
    import inspect
    from pathlib import Path

    d = torch.load(Path(__file__).resolve().parent / "golden" / "ref_signatures.pt")
    mine = {"GNN_Diffusion.__init__": dab.GNN_Diffusion.__init__, "GNN_Diffusion_3d.__init__": dab.GNN_Diffusion_3d.__init__,
            "Eff_GAT.__init__": dab.Eff_GAT.__init__, "Eff_GAT_3d.__init__": dab.Eff_GAT_3d.__init__,
            "Transformer_GNN.__init__": dab.Transformer_GNN.__init__, "Exophormer_GNN.__init__": dab.Exophormer_GNN.__init__,
            "PointNet.__init__": dab.PointNet.__init__, "GNN_Diffusion.forward_with_feats": dab.GNN_Diffusion.forward_with_feats,
            "GNN_Diffusion.p_sample_loop": dab.GNN_Diffusion.p_sample_loop, "GNN_Diffusion.p_losses": dab.GNN_Diffusion.p_losses,
            "GNN_Diffusion_3d.forward_with_feats": dab.GNN_Diffusion_3d.forward_with_feats,
            "Eff_GAT.forward_with_feats": dab.Eff_GAT.forward_with_feats,
            "Eff_GAT_3d.forward_with_feats": dab.Eff_GAT_3d.forward_with_feats}
    assert set(d) == set(mine)
    for key, want in d.items():
        params = inspect.signature(mine[key]).parameters
        order = [n for n in params if n not in ("self", "args", "kwargs")]
        ref_order = [name for name, _, _ in want]
        assert order[: len(ref_order)] == ref_order, (key, order, ref_order)      # positional calls keep working
        for name, has_default, dflt in want:
            p = params[name]
            if has_default:
                assert p.default is not inspect._empty, (key, name)
                got = p.default.name if hasattr(p.default, "name") and hasattr(p.default, "value") else repr(p.default)
                assert got == dflt, (key, name, got, dflt)


def test_dense_attention_kernel_keeps_two_ctas_per_sm(lib_built):
    """The 32-channel dense attention kernel is paced by its softmax warps and needs TWO resident CTAs per SM:
    2 x 192 threads x R registers <= 65536 -> R <= 168 (after rounding to the allocation granule).  ptxas' allocation of
    this kernel moves between 165 and 196 registers with unrelated edits (measured: 171 registers = one CTA per SM =
    0.26 ms instead of 0.19 ms per launch), so the build log is checked here, on the CPU, every round."""
    import re
    from pathlib import Path

    log = (Path(lib_built).parent.parent / "build" / "attn_dense.log").read_text()
    m = re.search(r"attn_dense_kernelILi32ELi4E.*?Used (\d+) registers", log, re.S)
    assert m, "attn_dense_kernel<32, 4> not found in the ptxas log"
    assert int(m.group(1)) <= 168, f"attn_dense_kernel<32, 4> uses {m.group(1)} registers: only one CTA per SM would fit"


def test_persistent_hidden_kernel_fits_one_cta_per_sm_without_spills(lib_built):
    """The persistent hidden-layer kernel (csrc/attn_hidden.cu) is ONE CTA of 384 threads per SM: 65536 / 384 -> at most
    168 registers per thread, and the production instantiation (no tracing, no running maximum: <false, true>) must reach
    that without local-memory spills -- the variant that kept the 64 scores live for the rare path spilled 60 bytes and
    lost 20 %.  Shared memory: 2 streams x 96 KB + barriers must stay under the 227 KB opt-in limit."""
    import re
    from pathlib import Path

    log = (Path(lib_built).parent.parent / "build" / "attn_hidden.log").read_text()
    m = re.search(r"attn_hidden_persist_kernelILb0ELb1EE.*?(\d+) bytes spill stores.*?Used (\d+) registers", log, re.S)
    assert m, "attn_hidden_persist_kernel<false, true> not found in the ptxas log"
    assert int(m.group(2)) <= 168, f"{m.group(2)} registers: the 384-thread CTA would not launch"
    assert int(m.group(1)) == 0, f"{m.group(1)} bytes of spill stores in the production score loop"
    src = (Path(lib_built).parent.parent / "csrc" / "attn_hidden.cu").read_text()
    assert "constexpr int NTH = 384;" in src and "2 * HST * KV_STAGE + SKIP_BYTES + 2 * OUT_PLANE" in src
    assert 2 * (2 * 4 * 8192 + 16384 + 2 * 8192) + 1024 <= 227 * 1024


def test_visual_backbone_state_dict_uses_timm_names():
    """N4: the mirror's EfficientNet-B0 encoder carries timm's parameter names (all 7 stages, so that a reference
    checkpoint's `model.visual_backbone.*` entries load strictly), identical to the oracle's."""
    from oracle.efficientnet import EfficientNetB0FeaturesRef

    mine = {k: tuple(v.shape) for k, v in dab.EfficientNetB0Features().state_dict().items()}
    want = {k: tuple(v.shape) for k, v in EfficientNetB0FeaturesRef().state_dict().items()}
    assert mine == want
    for k in ("conv_stem.weight", "bn1.running_var", "blocks.0.0.conv_dw.weight", "blocks.0.0.se.conv_reduce.bias",
              "blocks.2.1.conv_pwl.weight", "blocks.4.2.bn3.weight", "blocks.6.0.conv_pw.weight"):
        assert k in mine, k
    assert mine["blocks.2.0.conv_dw.weight"] == (144, 1, 5, 5) and mine["blocks.4.0.se.conv_reduce.weight"] == (20, 480, 1, 1)
