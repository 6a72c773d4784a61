"""Gram / VGG style loss (north_star addition, not in the reference: parity is against the published
formulation restated in oracle/restate.py -- "parity unpinned" vs the reference)."""
import pytest
import torch

from tests.util import assert_parity

pytestmark = pytest.mark.gpu
DEV = "cuda"


def synth_images(B, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(B, 3, H, W, generator=g) * 2 - 1


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-4), ("bf16", 6e-2)])
def test_style_loss_fwd_bwd(precision, tol):
    from multi_style_transfer_gan_b200.style_loss import GramStyleLoss, VGG19Features
    from oracle import restate as R
    vgg = VGG19Features(DEV, seed=0)
    wcpu = [(w.cpu(), b.cpu()) for w, b in vgg.weights]
    y = synth_images(2, 32, 48, 1)
    s = synth_images(2, 32, 48, 2)
    yr = y.clone().requires_grad_(True)
    ref = R.style_loss(R.vgg19_features(wcpu, yr), R.vgg19_features(wcpu, s))
    ref.backward()
    loss_fn = GramStyleLoss(vgg, precision).set_style(s.to(DEV))
    yd = y.to(DEV).requires_grad_(True)
    loss = loss_fn(yd)
    loss.backward()
    assert abs(float(loss) - float(ref)) <= tol * abs(float(ref)), (float(loss), float(ref))
    assert_parity(yd.grad, yr.grad, 1e-3 if precision == "fp32" else 1e-1, "dL/dy")


def test_config3_shapes_run():
    """BASELINE.json configs[2] feature shapes: Gram + loss fwd/bwd on relu1_1..relu5_1 sized maps (B=4 here)."""
    from multi_style_transfer_gan_b200 import ops
    B = 4
    for C, HW in ((64, 256), (128, 128), (256, 64), (512, 32), (512, 16)):
        f = torch.relu(torch.randn(B, HW, HW, C, device=DEV)).bfloat16()
        tgt = ops.gram(torch.relu(torch.randn(B, HW, HW, C, device=DEV)).bfloat16())
        loss, g = ops.gram_loss_fwd(f, tgt)
        df = ops.gram_loss_bwd(f, g, tgt)
        ff = f.float().reshape(B, HW * HW, C)
        gref = torch.einsum("bpc,bpd->bcd", ff, ff) / (C * HW * HW)
        assert_parity(g, gref, 2e-3, f"gram C={C}")
        assert torch.isfinite(df.float()).all() and torch.isfinite(loss).all()
