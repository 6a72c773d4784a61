"""The row-ring schedule of the fused output-layer kernel (slab.py: out7_ring_row_mmas / out7_ring_weights, restated in
csrc/out7_ring.cu) executed on tensors: every (ky, kx) contribution reaches its output row exactly once, a ring slot is never shared by
two live rows with the drain running OUT7_RING_LEAD steps late, and the drained result equals Conv2d(64, 3, 7, padding 3)
(enhanced_generator.py:130-133)."""
import pytest
import torch
import torch.nn.functional as F

from multi_style_transfer_gan_b200 import slab


@pytest.mark.parametrize("H,W,seg", [(9, 20, 9), (23, 12, 8), (50, 7, 50), (37, 9, 16)])
def test_out7_ring_schedule_equals_conv7(H, W, seg):
    torch.manual_seed(H * 100 + W)
    x = torch.randn(1, 64, H, W, dtype=torch.float64)
    w = torch.randn(3, 64, 7, 7, dtype=torch.float64)
    ref = F.conv2d(x, w, padding=3)[0]
    wall = slab.out7_ring_weights(w, dtype=torch.float64).reshape(7, 128, 64)
    assert float(wall.reshape(7, 8, 16, 64)[:, :, 3:].abs().max()) == 0.0 and float(wall.reshape(7, 8, 16, 64)[:, 7].abs().max()) == 0.0
    xp = F.pad(x[0].permute(1, 2, 0), (0, 0, 4, 4))                                         # [H, W + 8, 64]: slab halo, zero filled
    out = torch.zeros(H, W, 3, dtype=torch.float64)
    lead = slab.OUT7_RING_LEAD
    for y0 in range(0, H, seg):
        y1 = min(H, y0 + seg)
        tmem = torch.zeros(W, 512, dtype=torch.float64)
        owner = {}

        def drain(r):
            o = r - 3
            if y0 <= o < y1:
                c = (o % slab.OUT7_RING_SLOTS) * 16
                out[o] = tmem[:, c:c + 3]
                assert float(tmem[:, c + 3:c + 16].abs().max()) == 0.0                      # padded channels only ever get zero weights
                tmem[:, c:c + 16] = 0
                owner.pop(c, None)

        for r in range(y0 - 3, y1 + 3):
            drain(r - lead)
            if 0 <= r < H:
                for e0, n, col in slab.out7_ring_row_mmas(r, y0, y1):
                    for kx in range(7):
                        a = xp[r, 4 + kx - 3:4 + kx - 3 + W]
                        tmem[:, col:col + 16 * n] += a @ wall[kx, 16 * e0:16 * (e0 + n)].T
                    for j in range(n):
                        o = r - 3 + e0 + j
                        assert owner.setdefault(col + 16 * j, o) == o, "two live rows share a ring slot"
        for r in range(y1 + 3 - lead, y1 + 3):
            drain(r)
        assert not owner and float(tmem.abs().max()) == 0.0
    assert torch.allclose(out.permute(2, 0, 1), ref, atol=1e-9, rtol=1e-9)
