"""Shared helpers for the parity tests (test infrastructure; may import the oracle)."""
import torch

import oracle

TOL = 1e-4  # BASELINE.json north_star: per-step poses within 1e-4 relative of the fp32 oracle


def rel_err(y, ref):
    """max|y - ref| / max|ref|  (SURVEY.md section 8d parity metric)."""
    y, ref = y.detach().double().cpu(), ref.detach().double().cpu()
    return ((y - ref).abs().max() / ref.abs().max().clamp_min(1e-30)).item()


def quat_rel_err(y, ref):
    """Quaternion block compared up to sign (pytorch3d releases differ on standardisation)."""
    y, ref = y.detach().double().cpu(), ref.detach().double().cpu()
    sign = torch.where((y[:, :4] * ref[:, :4]).sum(-1, keepdim=True) < 0, -1.0, 1.0)
    y = torch.cat([y[:, :4] * sign, y[:, 4:]], 1)
    return ((y - ref).abs().max() / ref.abs().max()).item()


def make_pair_2d(seed=0, steps=300, sampling="DDPM", architecture="transformer", virt_nodes=4, rotation=True,
                 model_mean_type="EPSILON", inference_ratio=1, gemm_mode="fp32", attn_mode="csr", **extra):
    """Oracle module + product module sharing one seeded state_dict."""
    import diffassemble_b200 as dab

    torch.manual_seed(seed)
    ref = oracle.GNNDiffusionRef(
        steps=steps, sampling=sampling, rotation=rotation, architecture=architecture, virt_nodes=virt_nodes,
        model_mean_type=oracle.ModelMeanType[model_mean_type], inference_ratio=inference_ratio, **extra)
    ref.eval()
    mod = dab.GNN_Diffusion(
        steps=steps, sampling=sampling, rotation=rotation, architecture=architecture, virt_nodes=virt_nodes,
        model_mean_type=dab.ModelMeanType[model_mean_type], inference_ratio=inference_ratio, gemm_mode=gemm_mode,
        attn_mode=attn_mode, **extra)
    missing = mod.load_state_dict(ref.state_dict(), strict=True)
    return ref, mod


def make_pair_3d(seed=0, steps=300, backbone="pointnet", inference_ratio=10, model_mean_type="START_X",
                 gemm_mode="fp32", attn_mode="csr"):
    import diffassemble_b200 as dab

    torch.manual_seed(seed)
    ref = oracle.GNNDiffusion3dRef(steps=steps, backbone=backbone, inference_ratio=inference_ratio,
                                   model_mean_type=oracle.ModelMeanType[model_mean_type])
    ref.eval()
    mod = dab.GNN_Diffusion_3d(steps=steps, sampling="DDIM", backbone=backbone, inference_ratio=inference_ratio,
                               model_mean_type=dab.ModelMeanType[model_mean_type], gemm_mode=gemm_mode,
                               attn_mode=attn_mode)
    mod.load_state_dict(ref.state_dict(), strict=False)
    return ref, mod


def synth_graph_batch(sizes, kind="dense", degree="60%", seed=0):
    """edge_index / batch for a list of graph sizes (dense incl. self loops, or Exphander)."""
    import numpy as np

    eis = []
    for g, n in enumerate(sizes):
        if kind == "dense":
            eis.append(oracle.dense_edge_index(n))
        else:
            rng = np.random.default_rng(seed + g)
            eis.append(oracle.generate_random_expander(n, degree, rng=rng, check_spectral_gap=False).t().contiguous())
    return oracle.batch_graphs(eis, sizes)


def reseed_parameters(module, seed=0, gain=1.5, qk_gain=1.5):
    """Overwrite every parameter with values derived from (seed, parameter NAME, shape) only.

    The reference module and the oracle create their sub-modules in different orders, so
    ``torch.manual_seed`` + default initialisers give them different weights.  This makes the
    weights a function of the state_dict key alone: ``tests/golden/make_reference_golden.py``
    applies it to the real reference module, the tests apply it to the oracle / CUDA mirror, and
    both end up with identical tensors without 13 MB of weights per fixture in the repository.
    Magnitudes follow torch's defaults (uniform +-1/sqrt(fan_in), N(0,1) embeddings) times a gain
    (weights x1.5, query/key weights x1.5 again): with the plain defaults the attention of a
    random network is uniform to 1e-5 and the outputs barely differ between nodes, which would
    make the fixtures blind to attention errors; with the gain alpha spans 1e-4 .. 0.9.
    """
    import zlib

    with torch.no_grad():
        for name, p in sorted(module.named_parameters()):
            g = torch.Generator().manual_seed((zlib.crc32(name.encode()) + 7919 * seed) % (2**31))
            if "emb" in name:
                v = torch.randn(p.shape, generator=g)
            else:
                fan_in = p.shape[-1] if p.dim() > 1 else max(p.numel(), 16)
                v = (torch.rand(p.shape, generator=g) * 2 - 1) / fan_in ** 0.5
                if p.dim() > 1:
                    v = v * gain
                if "lin_query" in name or "lin_key" in name:
                    v = v * qk_gain
            p.copy_(v.to(p.dtype))
    return module


def pointnet_fixture_weights(module, seed):
    """Parameters AND BatchNorm running statistics of a PointNet exactly as tests/golden/make_reference_golden.py
    (case_pointnet) set them on the reference module."""
    from torch import nn

    reseed_parameters(module, seed, gain=1.0, qk_gain=1.0)
    with torch.no_grad():
        for k in range(1, 6):
            getattr(module, f"bn{k}").weight.add_(1.0)
        g = torch.Generator().manual_seed(4242 + seed)
        for name, m in sorted(module.named_modules()):
            if isinstance(m, nn.modules.batchnorm._BatchNorm):
                m.running_mean.copy_(0.2 * torch.randn(m.running_mean.shape, generator=g))
                m.running_var.copy_(0.5 + torch.rand(m.running_var.shape, generator=g))
    return module
