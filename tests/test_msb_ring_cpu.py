"""The row-ring schedule of the MultiScaleBlock branch kernel (slab.py: ring_row_mmas / ring_col / msb_ring_weights, restated in
csrc/msb_ring.cu) executed on tensors, for C = 64 (one pass) and C = 128 (three passes): every (branch, ky, kx) contribution
reaches its output row exactly once, a ring slot is never shared by two live rows, and the drained result equals the four
branch convolutions (enhanced_generator.py:52-71)."""
import pytest
import torch
import torch.nn.functional as F

from multi_style_transfer_gan_b200 import slab


@pytest.mark.parametrize("C,H,W,seg", [(64, 9, 20, 9), (64, 23, 12, 8), (64, 16, 140, 5), (64, 37, 9, 16),
                                       (128, 11, 10, 11), (128, 26, 7, 9), (128, 19, 130, 6), (128, 45, 5, 45)])
def test_ring_schedule_equals_the_four_branch_convs(C, H, W, seg):
    torch.manual_seed(H * 100 + W + C)
    Q, KB = C // 4, C // 64
    x = torch.randn(1, C, H, W, dtype=torch.float64)
    ws = [torch.randn(Q, C, k, k, dtype=torch.float64) for k in (1, 3, 3, 3)]
    ref = torch.cat([F.conv2d(x, w, padding=(w.shape[2] // 2) * d, dilation=d) for w, d in zip(ws, (1, 1, 2, 4))], 1)
    wall = slab.msb_ring_weights(ws, C, dtype=torch.float64)
    xp = F.pad(x[0].permute(1, 2, 0), (0, 0, 4, 4))                                     # [H, W + 8, C]: the slab halo (TMA zero fill)
    out = torch.zeros(H, W, C, dtype=torch.float64)
    w_off = 0
    for ps in range(len(slab.RING_PASSES[C])):
        rows = slab.ring_pass_rows(C, ps)
        wst = wall[w_off:w_off + KB * rows].reshape(KB, rows, 64)                       # this pass's stacks, per 64-channel block
        w_off += KB * rows
        lay = slab.ring_layout(C, ps)
        for y0 in range(0, H, seg):                                                     # strip segments, as the CTAs take them
            y1 = min(H, y0 + seg)
            tmem = torch.zeros(W, 512, dtype=torch.float64)                             # lanes = strip pixels, zeroed accumulators
            owner = {}                                                                  # column -> (branch, row) currently live
            lead = slab.RING_LEAD[C][ps]

            def drain(r):
                # the epilogue of step r: rows finished by this input row are drained and their slots zeroed
                for b in lay:
                    y = r - slab.RING_DIL[b]
                    if y0 <= y < y1:
                        c = slab.ring_col(b, y, C, ps)
                        acc = tmem[:, c:c + Q].clone()
                        tmem[:, c:c + Q] = 0
                        owner.pop(c, None)
                        if slab.ring_dup(C, ps, b) == 2:                                # even-row set + odd-row set
                            c2 = c + max(1, slab.RING_DIL[b]) * lay[b][0] * Q
                            acc += tmem[:, c2:c2 + Q]
                            tmem[:, c2:c2 + Q] = 0
                            owner.pop(c2, None)
                        out[y, :, Q * b:Q * b + Q] = acc

            for r in range(y0 - 4, y1 + 4):                                             # one step per input row, in order
                # the issuers run up to `lead` steps ahead of the epilogue: in the worst case only the steps <= r - lead are drained
                drain(r - lead)
                if 0 <= r < H:
                    for b, sx, e0, n, col in slab.ring_row_mmas(r, y0, y1, C, ps):
                        w0 = 0 if b == 0 else slab.ring_stack_row(b, sx, C, ps) + Q * e0
                        if slab.ring_dup(C, ps, b) == 2:
                            col += (r & 1) * max(1, slab.RING_DIL[b]) * lay[b][0] * Q
                        for kb in range(KB):
                            a = xp[r, 4 + sx:4 + sx + W, kb * 64:(kb + 1) * 64]         # shifted slab view [W, 64]
                            tmem[:, col:col + Q * n] += a @ wst[kb, w0:w0 + Q * n].T
                        d = max(1, slab.RING_DIL[b])
                        for j in range(n):
                            y = r if b == 0 else r + (e0 + j - 1) * d
                            assert owner.setdefault(col + Q * j, (b, y)) == (b, y), "two live rows share a ring slot"
            for r in range(y1 + 4 - lead, y1 + 4):
                drain(r)
            assert not owner and float(tmem.abs().max()) == 0.0                         # every slot drained and zero at the end
    assert w_off == wall.shape[0]
    assert torch.allclose(out.permute(2, 0, 1), ref[0], atol=1e-9, rtol=1e-9)


@pytest.mark.parametrize("C", [64, 128])
def test_ring_geometry_fits_tensor_memory(C):
    Q = C // 4
    for ps in range(len(slab.RING_PASSES[C])):
        lay = slab.ring_layout(C, ps)
        spans = []
        for b, (R, base) in lay.items():
            cols = sorted({slab.ring_col(b, y, C, ps) for y in range(80)})
            assert cols[0] == base and len(cols) == max(1, slab.RING_DIL[b]) * R
            spans.append((cols[0], cols[-1] + Q))
        spans.sort()
        assert spans[-1][1] <= 512
        for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
            assert a1 <= b0
