"""Scope row N4 on the GPU: the EfficientNet-B0 patch encoder (CUDA operators) against the oracle (pinned against
torchvision's implementation), stand-alone and through ``GNN_Diffusion.forward`` / ``p_sample_loop`` with image patches."""
import pytest
import torch

import oracle
from common import TOL, reseed_parameters, rel_err, synth_graph_batch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _randomize_bn(module, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, m in sorted(module.named_modules()):
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.copy_(0.2 * torch.randn(m.running_mean.shape, generator=g))
                m.running_var.copy_(0.5 + torch.rand(m.running_var.shape, generator=g))
                m.weight.copy_(0.5 + torch.rand(m.weight.shape, generator=g))
                m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g))


@pytest.mark.parametrize("n,size", [(7, 32), (300, 32), (5, 64)])
def test_efficientnet_b0_feature_pyramid(n, size):
    import diffassemble_b200 as dab
    from oracle.efficientnet import EfficientNetB0FeaturesRef

    torch.manual_seed(1)
    ref = EfficientNetB0FeaturesRef().eval()
    _randomize_bn(ref, 3)
    enc = dab.EfficientNetB0Features(chunk=128).eval()
    enc.load_state_dict(ref.state_dict(), strict=True)
    enc = enc.to(DEV)
    x = torch.randn(n, 3, size, size)
    with torch.no_grad():
        want = ref(x)
    got = enc(x.to(DEV))
    assert len(got) == 4
    for g_, w_ in zip(got, want[:4]):
        assert g_.shape == w_.shape
        assert rel_err(g_, w_) < 2e-5, rel_err(g_, w_)


def test_visual_features_and_forward_with_patches():
    """Eff_GAT.visual_features (normalisation + encoder + the 40 x 4 x 4 | 112 x 2 x 2 flatten) and the full
    `forward(xy_pos, time, patch_rgb, edge_index, batch)` of the reference API with IMAGE PATCHES as the condition."""
    import diffassemble_b200 as dab

    ref = oracle.GNNDiffusionRef(steps=100, sampling="DDIM", rotation=True, architecture="exophormer", virt_nodes=4,
                                 model_mean_type=oracle.ModelMeanType.START_X, inference_ratio=10).eval()
    enc_state = {k: v.clone() for k, v in ref.model.visual_backbone.state_dict().items()}
    reseed_parameters(ref, 5)
    # (the encoder keeps torch's default initialisation: the x1.5 gains of reseed_parameters compound over its 16 blocks
    # into activations of 1e16, where a comparison is meaningless)
    ref.model.visual_backbone.load_state_dict(enc_state)
    _randomize_bn(ref, 6)
    mod = dab.GNN_Diffusion(steps=100, sampling="DDIM", rotation=True, architecture="exophormer", virt_nodes=4,
                            model_mean_type=dab.ModelMeanType.START_X, inference_ratio=10)
    mod.load_state_dict(ref.state_dict(), strict=True)
    mod = mod.to(DEV).eval()
    sizes = [36, 64]
    ei, batch = synth_graph_batch(sizes, kind="expander", degree="60%")
    M = sum(sizes)
    g = torch.Generator().manual_seed(2)
    patches = torch.rand(M, 3, 32, 32, generator=g)
    x = torch.randn(M, 4, generator=g)
    t = torch.full((M,), 90, dtype=torch.long)
    with torch.no_grad():
        feats_ref = ref.model.visual_features(patches)
        want = ref.model.forward_with_feats(x, t, patches, ei, feats_ref, batch)[0]
    feats = mod.visual_features(patches.to(DEV))
    assert feats.shape == (M, 1088)
    assert rel_err(feats, feats_ref) < 2e-5
    got = mod.forward(x.to(DEV), t.to(DEV), patches.to(DEV), ei.to(DEV), batch.to(DEV))[0]   # (final_feats, attentions), efficient_gat.py:146
    assert rel_err(got, want) < TOL
    # the sampling loop takes patches as `cond` (spatial_diffusion.py:653 runs the encoder once per loop)
    imgs, _ = mod.p_sample_loop((M, 4), patches.to(DEV), ei.to(DEV), batch.to(DEV), generator=torch.Generator(device=DEV).manual_seed(0))
    assert len(imgs) == 10 and torch.isfinite(imgs[-1]).all()
