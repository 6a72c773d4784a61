"""Pins the oracle's VGG-19 trunk restatement (oracle/restate.py:vgg19_features) against the named third-party
dependency: torchvision's ``vgg19(weights=None).features[:30]`` with forward hooks on the five ReLU taps
(relu1_1 .. relu5_1 = indices 1, 6, 11, 20, 29), on shared seeded weights.  CPU only.  (the Gram / VGG loss
is not in the reference, so torchvision is the anchor for the trunk; the Gram normalisation has no reference.)"""
import pytest
import torch

from oracle import restate as R

tv = pytest.importorskip("torchvision")


def _trunk(seed):
    torch.manual_seed(seed)
    return tv.models.vgg19(weights=None).features[:30].eval()


@pytest.mark.parametrize("seed,size", [(0, 64), (3, 48)])
def test_vgg19_taps_match_torchvision(seed, size):
    trunk = _trunk(seed)
    convs = [m for m in trunk if isinstance(m, torch.nn.Conv2d)]
    assert len(convs) == 13
    weights = [(m.weight.detach(), m.bias.detach()) for m in convs]
    g = torch.Generator().manual_seed(7)
    x = torch.rand(2, 3, size, size, generator=g) * 2 - 1
    taps = {}
    hooks = [trunk[i].register_forward_hook(lambda m, a, out, i=i: taps.__setitem__(i, out.clone())) for i in R.VGG19_TAPS]
    with torch.no_grad():
        trunk(x)
        ours = R.vgg19_features(weights, x)
    for h in hooks:
        h.remove()
    assert len(ours) == 5
    for i, o in zip(R.VGG19_TAPS, ours):
        assert o.shape == taps[i].shape
        torch.testing.assert_close(o, taps[i], rtol=0, atol=0)          # same ATen ops in the same order: bit-identical


def test_vgg19_tap_shapes():
    trunk = _trunk(0)
    weights = [(m.weight.detach(), m.bias.detach()) for m in trunk if isinstance(m, torch.nn.Conv2d)]
    with torch.no_grad():
        f = R.vgg19_features(weights, torch.zeros(1, 3, 32, 32))
    assert [tuple(t.shape[1:]) for t in f] == [(64, 32, 32), (128, 16, 16), (256, 8, 8), (512, 4, 4), (512, 2, 2)]


def test_gram_normalisation():
    f = torch.randn(2, 8, 5, 7)
    G = R.gram(f)
    ref = torch.einsum("bcp,bdp->bcd", f.flatten(2), f.flatten(2)) / (8 * 5 * 7)
    torch.testing.assert_close(G, ref, rtol=1e-5, atol=1e-6)
