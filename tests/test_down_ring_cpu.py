"""The row-ring schedule of the stride-2 down conv kernel (slab.py: down_ring_row_mmas / down_ring_weights, restated in
csrc/down_ring.cu) executed on tensors: even / odd pixel slabs, every (ky, kx) contribution reaches its output row exactly once, a
ring slot is never shared by two live rows with the drain running DOWN_RING_LEAD steps late, and the drained result equals
Conv2d(64, Cout, 4, stride 2, padding 1) (enhanced_generator.py:99-104)."""
import pytest
import torch
import torch.nn.functional as F

from multi_style_transfer_gan_b200 import slab


def run_schedule(Cout, H, W, seg, lead):
    torch.manual_seed(H * 100 + W + Cout)
    x = torch.randn(1, 64, H, W, dtype=torch.float64)
    w = torch.randn(Cout, 64, 4, 4, dtype=torch.float64)
    ref = F.conv2d(x, w, stride=2, padding=1)[0]                                            # [Cout, H/2, W/2]
    Ho, Wo = H // 2, W // 2
    wall = slab.down_ring_weights(w, dtype=torch.float64).reshape(Cout // 64, 4, 2, 128, 64)
    xh = x[0].permute(1, 2, 0)                                                              # [H, W, 64]
    even = F.pad(xh[:, 0::2], (0, 0, 4, 4))                                                 # [H, Wo + 8, 64]: slab halo, zero filled
    odd = F.pad(xh[:, 1::2], (0, 0, 4, 4))
    view = {0: (odd, -1), 1: (even, 0), 2: (odd, 0), 3: (even, 1)}                          # kx -> (slab, pixel shift)
    out = torch.zeros(Ho, Wo, Cout, dtype=torch.float64)
    for g in range(Cout // 64):
        for y0 in range(0, Ho, seg):
            y1 = min(Ho, y0 + seg)
            tmem = torch.zeros(Wo, 512, dtype=torch.float64)
            owner = {}

            def drain(r):
                o = r // 2 - 1
                if r % 2 == 0 and y0 <= o < y1:
                    c = (o % 8) * 64
                    out[o, :, g * 64:(g + 1) * 64] = tmem[:, c:c + 64]
                    tmem[:, c:c + 64] = 0
                    owner.pop(c, None)

            steps = list(range(2 * y0 - 1, 2 * y1 + 1))
            for r in steps:
                drain(r - lead)
                if 0 <= r < H:
                    for e0, n, col in slab.down_ring_row_mmas(r, y0, y1):
                        for kx in range(4):
                            sl, dx = view[kx]
                            a = sl[r, 4 + dx:4 + dx + Wo]
                            tmem[:, col:col + 64 * n] += a @ wall[g, kx, r % 2, 64 * e0:64 * (e0 + n)].T
                        for j in range(n):
                            o = (r - 1) // 2 + e0 + j
                            assert owner.setdefault(col + 64 * j, o) == o, "two live rows share a ring slot"
            for r in range(steps[-1] + 1 - lead, steps[-1] + 1):
                drain(r)
            assert not owner and float(tmem.abs().max()) == 0.0
    return out.permute(2, 0, 1), ref


@pytest.mark.parametrize("Cout,H,W,seg", [(64, 10, 20, 5), (128, 24, 12, 4), (64, 16, 6, 3), (64, 60, 8, 30)])
def test_down_ring_schedule_equals_strided_conv(Cout, H, W, seg):
    out, ref = run_schedule(Cout, H, W, seg, slab.DOWN_RING_LEAD)
    assert torch.allclose(out, ref, atol=1e-9, rtol=1e-9)


def test_down_ring_lead_has_a_limit():
    with pytest.raises(AssertionError):
        run_schedule(64, 60, 4, 30, 14)
