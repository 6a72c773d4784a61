"""The row-ring schedule of the transposed-conv kernel (slab.py: convt_ring_row_mmas / convt_ring_weights, restated in
csrc/convt_ring.cu) executed on tensors: every (ky, kx) contribution reaches its output row exactly once, a ring slot is never
shared by two live rows even with the drain running CONVT_RING_LEAD steps late, and the drained result equals
ConvTranspose2d(4, 2, 1) (enhanced_generator.py:116-123)."""
import pytest
import torch
import torch.nn.functional as F

from multi_style_transfer_gan_b200 import slab


def run_schedule(Cin, Cout, H, W, seg, lead):
    torch.manual_seed(H * 100 + W + Cin)
    KB = Cin // 64
    x = torch.randn(1, Cin, H, W, dtype=torch.float64)
    w = torch.randn(Cin, Cout, 4, 4, dtype=torch.float64)
    ref = F.conv_transpose2d(x, w, stride=2, padding=1)[0]                                  # [Cout, 2H, 2W]
    wall = slab.convt_ring_weights(w, dtype=torch.float64).reshape(Cout // 64, 2, KB, 2, 256, 64)
    xp = F.pad(x[0].permute(1, 2, 0), (0, 0, 4, 4))                                         # [H, W + 8, Cin]: the slab halo (TMA zero fill)
    out = torch.zeros(2 * H, 2 * W, Cout, dtype=torch.float64)
    for g in range(Cout // 64):
        for px in range(2):
            for y0 in range(0, H, seg):                                                     # pieces of input rows, as the CTAs take them
                y1 = min(H, y0 + seg)
                tmem = torch.zeros(W, 512, dtype=torch.float64)
                owner = {}

                def drain(r):
                    for o in (2 * r - 1, 2 * r):                                            # the rows input row r completed
                        if 2 * y0 <= o < 2 * y1:
                            c = (o % slab.CONVT_RING_SLOTS) * 64
                            out[o, px::2, g * 64:(g + 1) * 64] = tmem[:, c:c + 64]
                            tmem[:, c:c + 64] = 0
                            owner.pop(c, None)

                for r in range(y0 - 1, y1 + 1):
                    drain(r - lead)                                                         # worst case: only steps <= r - lead are drained
                    if 0 <= r < H:
                        for e0, n, col in slab.convt_ring_row_mmas(r, y0, y1):
                            for t in range(2):
                                dx = slab.CONVT_RING_DX[px][t]
                                for kb in range(KB):
                                    a = xp[r, 4 + dx:4 + dx + W, kb * 64:(kb + 1) * 64]
                                    tmem[:, col:col + 64 * n] += a @ wall[g, px, kb, t, 64 * e0:64 * (e0 + n)].T
                            for j in range(n):
                                o = 2 * r - 1 + e0 + j
                                assert owner.setdefault(col + 64 * j, o) == o, "two live rows share a ring slot"
                for r in range(y1 + 1 - lead, y1 + 1):
                    drain(r)
                assert not owner and float(tmem.abs().max()) == 0.0
    return out.permute(2, 0, 1), ref


@pytest.mark.parametrize("Cin,Cout,H,W,seg", [(64, 64, 9, 20, 9), (128, 64, 23, 12, 8), (64, 128, 16, 7, 5), (128, 64, 37, 9, 16)])
def test_convt_ring_schedule_equals_conv_transpose(Cin, Cout, H, W, seg):
    out, ref = run_schedule(Cin, Cout, H, W, seg, slab.CONVT_RING_LEAD)
    assert torch.allclose(out, ref, atol=1e-9, rtol=1e-9)


def test_convt_ring_lead_is_tight():
    with pytest.raises(AssertionError):
        run_schedule(64, 64, 37, 5, 37, slab.CONVT_RING_LEAD + 1)
