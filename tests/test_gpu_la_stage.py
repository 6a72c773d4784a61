"""Fused LocalAttention stage kernel (csrc/la_stage.cu: IN-apply -> qkv 1x1 -> window attention -> proj 1x1 in one
tcgen05 launch) against the oracle's LocalAttention (oracle/restate.py:local_attention = enhanced_generator.py:13-47),
through the C-ABI.  Tolerance: the bf16 tolerance of the unfused kernels' tests (rel 2e-2 of the output's max)."""
import pytest
import torch

from tests.util import assert_parity

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _case(C, N, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, C, H, W, generator=g)
    wq = torch.randn(3 * C, C, 1, 1, generator=g) * (2.0 / C) ** 0.5
    bq = torch.randn(3 * C, generator=g) * 0.1
    wp = torch.randn(C, C, 1, 1, generator=g) * (1.0 / C) ** 0.5
    bp = torch.randn(C, generator=g) * 0.1
    return x, wq, bq, wp, bp


def _bf(t):
    return t.bfloat16().float()


@pytest.mark.parametrize("C,N,H,W", [(64, 1, 4, 32), (64, 2, 8, 64), (128, 1, 4, 32), (128, 2, 16, 32),
                                       (64, 3, 12, 40), (128, 2, 8, 8), (64, 1, 64, 96), (128, 1, 32, 160),
                                       (64, 4, 128, 128), (128, 4, 64, 128)])
def test_la_stage_matches_oracle(C, N, H, W):
    from multi_style_transfer_gan_b200 import ops
    from oracle import restate as R
    x, wq, bq, wp, bp = _case(C, N, H, W, seed=C + H + W)
    xb, wqb, wpb = _bf(x), _bf(wq), _bf(wp)
    ref = R.local_attention(xb, wqb, bq, wpb, bp, ws=4)
    xd = xb.permute(0, 2, 3, 1).contiguous().to(DEV).bfloat16()
    wqd = ops.pack_weight(wq.to(DEV), ops.PACK_FWD, torch.bfloat16)
    wpd = ops.pack_weight(wp.to(DEV), ops.PACK_FWD, torch.bfloat16)
    assert ops.la_stage_supported(xd, wqd, wpd)
    out = ops.la_stage_fwd(xd, wqd, bq.to(DEV), wpd, bp.to(DEV))
    torch.cuda.synchronize()
    assert_parity(out.float().permute(0, 3, 1, 2).cpu(), ref, 2e-2, f"la_stage C={C} {H}x{W}")


@pytest.mark.parametrize("C,N,H,W", [(64, 2, 16, 64), (128, 2, 8, 64)])
def test_la_stage_fused_input_norm(C, N, H, W):
    """with in_stats: x is the RAW conv output, normalised (+ ReLU) on the fly exactly like msg_instnorm_apply"""
    from multi_style_transfer_gan_b200 import ops
    from oracle import restate as R
    x, wq, bq, wp, bp = _case(C, N, H, W, seed=7)
    x = x * 1.7 + 0.3
    xd = _bf(x).permute(0, 2, 3, 1).contiguous().to(DEV).bfloat16()
    st = ops.instnorm_stats(xd)
    a0 = ops.instnorm_apply(xd, st, ops.ACT_RELU)                       # the stand-alone apply: what the fused path must reproduce
    wqd = ops.pack_weight(wq.to(DEV), ops.PACK_FWD, torch.bfloat16)
    wpd = ops.pack_weight(wp.to(DEV), ops.PACK_FWD, torch.bfloat16)
    fused = ops.la_stage_fwd(xd, wqd, bq.to(DEV), wpd, bp.to(DEV), in_stats=st, in_act=ops.ACT_RELU)
    plain = ops.la_stage_fwd(a0, wqd, bq.to(DEV), wpd, bp.to(DEV))
    torch.cuda.synchronize()
    assert torch.equal(fused, plain)                                    # same arithmetic -> bit-identical
    ref = R.local_attention(a0.float().permute(0, 3, 1, 2).cpu(), _bf(wq), bq, _bf(wp), bp, ws=4)
    assert_parity(fused.float().permute(0, 3, 1, 2).cpu(), ref, 2e-2, "la_stage fused norm")


def test_la_stage_unsupported_shapes_are_refused():
    from multi_style_transfer_gan_b200 import _lib, ops
    x = torch.zeros(1, 8, 8, 32, device=DEV, dtype=torch.bfloat16)
    w = torch.zeros(32 * 96, device=DEV, dtype=torch.bfloat16)
    assert not ops.la_stage_supported(x, w, w)
    with pytest.raises(_lib.MsgError):
        ops.la_stage_fwd(x, w, torch.zeros(96, device=DEV), w, torch.zeros(32, device=DEV))
