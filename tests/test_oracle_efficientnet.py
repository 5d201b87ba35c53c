"""The oracle's EfficientNet-B0 feature pyramid (scope row N4; timm is absent from the image) against torchvision's
independent implementation of the same published architecture: identical weights (copied across by a key map, with
randomised BatchNorm statistics), identical feature taps."""
import torch

from oracle.efficientnet import EfficientNetB0FeaturesRef, torchvision_key_map, visual_features_ref


def _randomize(module, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in module.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.copy_(0.2 * torch.randn(m.running_mean.shape, generator=g))
                m.running_var.copy_(0.5 + torch.rand(m.running_var.shape, generator=g))
                m.weight.copy_(0.5 + torch.rand(m.weight.shape, generator=g))
                m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g))


def test_oracle_matches_torchvision_efficientnet_b0():
    import torchvision

    torch.manual_seed(0)
    tv = torchvision.models.efficientnet_b0(weights=None).eval()
    _randomize(tv, 1)
    ref = EfficientNetB0FeaturesRef().eval()
    km = torchvision_key_map()
    tv_sd = tv.features.state_dict()
    assert set(km) == set(ref.state_dict()), sorted(set(km) ^ set(ref.state_dict()))[:5]
    ref.load_state_dict({k: tv_sd[v] for k, v in km.items()}, strict=True)
    x = torch.randn(5, 3, 32, 32)
    with torch.no_grad():
        feats = ref(x)
        h, taps = x, {}
        for i, f in enumerate(tv.features[:8]):
            h = f(h)
            taps[i] = h
    # timm's features_only taps = the last block of every stride level: torchvision features[1, 2, 3, 5, 7]
    for mine, theirs in zip(feats, (1, 2, 3, 5, 7)):
        assert mine.shape == taps[theirs].shape
        assert torch.allclose(mine, taps[theirs], rtol=1e-5, atol=1e-6), (theirs, (mine - taps[theirs]).abs().max())
    assert [tuple(f.shape[1:]) for f in feats] == [(16, 16, 16), (24, 8, 8), (40, 4, 4), (112, 2, 2), (320, 1, 1)]
    mean = torch.tensor([0.485, 0.456, 0.406])[None, :, None, None]
    std = torch.tensor([0.229, 0.224, 0.225])[None, :, None, None]
    with torch.no_grad():
        pf = visual_features_ref(ref, torch.rand(5, 3, 32, 32), mean, std)
    assert pf.shape == (5, 1088)   # efficient_gat.py:48: 1088 + 32 + 32
