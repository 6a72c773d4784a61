"""Parity metric used by every test (BASELINE.json: "<=1e-4 relative on fp32 outputs and losses",
"<=2e-2 max-abs on the [-1,1] image for bf16").

"relative" is taken against the tensor's scale, not element-wise (element-wise relative error is
meaningless at zero crossings): both  max|a-b| / max|b|  and  ||a-b||_2 / ||b||_2  must be <= rel.
For reference, the unmodified reference itself differs from an fp64 evaluation of the same network
by 4.7e-5 (normalised max) / 7e-6 (rel-L2) on the c=16 golden, so 1e-4 is ~2x the fp32 noise floor.
"""
import torch


def parity_errors(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    d = a - b
    scale = float(b.abs().max())
    nrm = float(b.norm())
    return (float(d.abs().max()) / (scale if scale > 0 else 1.0),
            float(d.norm()) / (nrm if nrm > 0 else 1.0))


def assert_parity(a, b, rel=1e-4, name="", floor=0.0, max_rel=None):
    """floor: absolute error allowed in addition (for quantities that are ~0 by construction).
    max_rel: separate bound for the normalised max error (default: rel)."""
    max_rel = rel if max_rel is None else max_rel
    assert a.shape == b.shape, f"{name}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    assert torch.isfinite(a).all(), f"{name}: non-finite values"
    d = (a - b).abs()
    scale = float(b.abs().max())
    emax = float(d.max()) if d.numel() else 0.0
    el2 = float((a - b).norm())
    assert emax <= max_rel * scale + floor, f"{name}: max|a-b|={emax:.3e} > {max_rel:g}*max|b|={max_rel*scale:.3e} (+{floor:g})"
    assert el2 <= rel * float(b.norm()) + floor * d.numel() ** 0.5, \
        f"{name}: rel-L2 {el2/float(b.norm()+1e-300):.3e} > {rel:g}"
