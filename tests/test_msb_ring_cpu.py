"""The row-ring schedule of the C = 64 MultiScaleBlock branch kernel (slab.py: ring_row_mmas / ring_col / msb64_ring_weights,
restated in csrc/msb_ring.cu) executed on tensors: every (branch, ky, kx) contribution reaches its output row exactly once,
a ring slot is never shared by two live rows, and the drained result equals the four branch convolutions
(enhanced_generator.py:52-71)."""
import pytest
import torch
import torch.nn.functional as F

from multi_style_transfer_gan_b200 import slab


@pytest.mark.parametrize("H,W,seg", [(9, 20, 9), (23, 12, 8), (16, 140, 5), (37, 9, 16)])
def test_ring_schedule_equals_the_four_branch_convs(H, W, seg):
    torch.manual_seed(H * 100 + W)
    x = torch.randn(1, 64, H, W, dtype=torch.float64)
    ws = [torch.randn(16, 64, k, k, dtype=torch.float64) for k in (1, 3, 3, 3)]
    ref = torch.cat([F.conv2d(x, w, padding=(w.shape[2] // 2) * d, dilation=d) for w, d in zip(ws, (1, 1, 2, 4))], 1)
    wst = slab.msb64_ring_weights(ws, dtype=torch.float64)                              # [448, 64]
    xp = F.pad(x[0].permute(1, 2, 0), (0, 0, 4, 4))                                     # [H, W + 8, 64]: the slab halo (TMA zero fill)
    out = torch.zeros(H, W, 64, dtype=torch.float64)
    for y0 in range(0, H, seg):                                                         # strip segments, as the CTAs take them
        y1 = min(H, y0 + seg)
        tmem = torch.zeros(W, 512, dtype=torch.float64)                                 # lanes = strip pixels, zeroed accumulators
        owner = {}                                                                      # column -> (branch, row) currently live
        for r in range(y0 - 4, y1 + 4):                                                 # one step per input row, in order
            if 0 <= r < H:
                for b, sx, e0, n, col in slab.ring_row_mmas(r, y0, y1):
                    a = xp[r, 4 + sx:4 + sx + W, :]                                     # shifted slab view [W, 64]
                    w0 = 0 if b == 0 else slab.ring_stack_row(b, sx) + 16 * e0
                    tmem[:, col:col + 16 * n] += a @ wst[w0:w0 + 16 * n].T
                    d = max(1, slab.RING_DIL[b])
                    for j in range(n):
                        y = r if b == 0 else r + (e0 + j - 1) * d
                        assert owner.setdefault(col + 16 * j, (b, y)) == (b, y), "two live rows share a ring slot"
            # the epilogue of step r: rows finished by this input row are drained and their slots zeroed
            for b, lag in enumerate((0, 1, 2, 4)):
                y = r - lag
                if y0 <= y < y1:
                    c = slab.ring_col(b, y)
                    out[y, :, 16 * b:16 * b + 16] = tmem[:, c:c + 16]
                    tmem[:, c:c + 16] = 0
                    owner.pop(c, None)
        assert not owner and float(tmem.abs().max()) == 0.0                             # every slot drained and zero at the end
    assert torch.allclose(out.permute(2, 0, 1), ref[0], atol=1e-9, rtol=1e-9)


def test_ring_geometry_fits_tensor_memory():
    assert slab.RING_BASE[3] + 4 * slab.RING_SLOTS[3] * 16 <= 512
    cols = set()
    for b in range(4):
        for y in range(64):
            cols.add((b, slab.ring_col(b, y)))
    by_branch = {b: sorted(c for bb, c in cols if bb == b) for b in range(4)}
    for b in range(4):
        lo, hi = by_branch[b][0], by_branch[b][-1] + 16
        assert lo == slab.RING_BASE[b] and (b == 3 or hi <= slab.RING_BASE[b + 1])
