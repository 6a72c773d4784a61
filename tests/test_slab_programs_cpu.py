"""CPU: the host-side "programs" of the row-slab kernels (multi_style_transfer_gan_b200/slab.py) and the ctypes mirrors
of their descriptors.  A program is data (which input-row slabs a tile loads, which taps / accumulator columns each feeds,
how the weights are laid out); here each program + its weight packer is executed by a plain-PyTorch emulator that follows
the semantics documented in include/msg_b200.h, and compared with the convolution it must equal
(enhanced_generator.py:52-71 branches, :92 input conv, :137 output conv, :120/:127 transposed convs).  The CUDA kernels
are checked against the same convolutions in tests/test_gpu_conv_tc.py."""
import ctypes
import os
import subprocess
import sys
import tempfile

import pytest
import torch
import torch.nn.functional as F

from tests.conftest import ROOT


def _slab_mod():
    from multi_style_transfer_gan_b200 import slab
    return slab


def emulate_slab(prog, x, w_slab, out_stride=1, off=(0, 0)):
    """x [N,H,W,Cin_total] fp32, w_slab [n_tiles*ncols, 64] -> y [N,H*os,W*os,n_store] (zeros where this phase does not write)."""
    N, H, W, _ = x.shape
    halo = 8
    xp = F.pad(x, (0, 0, halo, halo + 1, halo, halo))          # zero padding = the TMA out-of-bounds fill
    acc = torch.zeros(N, H, W, prog.Ntot, dtype=torch.float64)
    tp = 0
    for kb, (dy, cb, taps) in enumerate(prog.kblocks):
        for sx, acc_col, kstep, _ in taps:
            if prog.pixel_pair:
                wt = w_slab[kb * prog.ncols:(kb + 1) * prog.ncols, kstep * 16:(kstep + 1) * 16].double()   # [ncols, 2 px * 8 ch]
                for p in range(2):
                    xs = xp[:, halo + dy:halo + dy + H, halo + sx + p:halo + sx + p + W, :8].double()
                    acc[..., acc_col:acc_col + prog.ncols] += xs @ wt[:, p * 8:(p + 1) * 8].t()
            else:
                wt = w_slab[tp * prog.ncols:(tp + 1) * prog.ncols].double()                               # [ncols, 64]
                xs = xp[:, halo + dy:halo + dy + H, halo + sx:halo + sx + W, cb * 64:(cb + 1) * 64].double()
                acc[..., acc_col:acc_col + prog.ncols] += xs @ wt.t()
            tp += 1
    y = torch.zeros(N, H * out_stride, W * out_stride, prog.n_store, dtype=torch.float64)
    y[:, off[0]::out_stride, off[1]::out_stride] = acc[..., :prog.n_store]
    return y


def emulate_shift(prog, x, w_rows):
    """x [N,H,W,Cin] -> y [N,H,W,n_out] following msg_shift_desc (multi-row tiles, shared slabs, epilogue shifts)."""
    N, H, W, _ = x.shape
    R, halo = prog.tile_rows, prog.halo
    pad = 8
    Hy = (H + R - 1) // R
    xp = F.pad(x, (0, 0, pad, pad, pad, pad + R)).double()
    y = torch.zeros(N, Hy * R, W, prog.n_out, dtype=torch.float64)
    for ty in range(Hy):
        y0 = ty * R
        D = torch.zeros(N, W + 2 * pad, prog.Ntot, dtype=torch.float64)      # D'[image column + pad][accumulator column]
        slab = None
        for kb in prog.kblocks:
            dy, cb, col0, ncols, wrow = kb[:5]
            same = kb[6] if len(kb) == 7 else 0
            if not same:
                slab = xp[:, pad + y0 + dy, :, cb * 64:(cb + 1) * 64]              # [N, W + 2 pad, 64]
            D[..., col0:col0 + ncols] += slab @ w_rows[wrow:wrow + ncols].double().t()
        for grp in prog.groups:
            col0, span, oc0, oc, terms = grp[:5]
            row = grp[5] if len(grp) == 6 else 0
            for shift, col in terms:       # out[x] += D'[x - halo + shift]
                y[:, y0 + row, :, oc0:oc0 + oc] += D[:, pad - halo + shift:pad - halo + shift + W, col0 + col:col0 + col + oc]
    return y[:, :H]


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize("C", [64, 128, 256])
def test_msb_program_equals_the_four_branch_convs(C):
    slab = _slab_mod()
    torch.manual_seed(C)
    q, H, W = C // 4, 9, 12
    x = torch.randn(1, C, H, W)
    ws = [torch.randn(q, C, k, k) for k in (1, 3, 3, 3)]
    ref = torch.cat([F.conv2d(x, w, padding=(w.shape[2] // 2) * d, dilation=d) for w, d in zip(ws, (1, 1, 2, 4))], 1)
    prog = slab.msb_program(C)
    y = emulate_slab(prog, nhwc(x), slab.msb_weight_slab(prog, ws, dtype=torch.float32))
    assert torch.allclose(y.permute(0, 3, 1, 2), ref.double(), atol=1e-3, rtol=1e-4)


def test_msb_dgrad_program_is_the_adjoint():
    slab = _slab_mod()
    torch.manual_seed(3)
    C, H, W = 64, 10, 9
    q = C // 4
    ws = [torch.randn(q, C, k, k) for k in (1, 3, 3, 3)]
    x = torch.randn(1, C, H, W, requires_grad=True)
    out = torch.cat([F.conv2d(x, w, padding=(w.shape[2] // 2) * d, dilation=d) for w, d in zip(ws, (1, 1, 2, 4))], 1)
    dB = torch.randn_like(out)
    out.backward(dB)
    prog = slab.msb_dgrad_program(C)
    y = emulate_slab(prog, nhwc(dB), slab.msb_dgrad_weight_slab(prog, ws, dtype=torch.float32))
    assert torch.allclose(y.permute(0, 3, 1, 2), x.grad.double(), atol=1e-3, rtol=1e-4)


def test_conv7_input_program_pixel_pair():
    slab = _slab_mod()
    torch.manual_seed(4)
    c, H, W = 16, 9, 16
    x = torch.randn(1, 3, H, W)
    w = torch.randn(c, 3, 7, 7)
    x8 = F.pad(nhwc(x), (0, 5))
    prog = slab.conv7_in_program(c)
    y = emulate_slab(prog, x8, slab.conv7_in_weight_slab(prog, w, dtype=torch.float32))
    assert torch.allclose(y.permute(0, 3, 1, 2), F.conv2d(x, w, padding=3).double(), atol=1e-3, rtol=1e-4)


@pytest.mark.parametrize("c,H,rows", [(64, 8, 2), (64, 7, 2), (128, 5, 2), (64, 6, 1), (64, 8, 4), (64, 10, 4), (64, 5, 4)])
def test_conv7_output_shift_program(c, H, rows):
    slab = _slab_mod()
    torch.manual_seed(5)
    W = 11
    x = torch.randn(2, c, H, W)
    w = torch.randn(3, c, 7, 7)
    prog = slab.conv7_out_shift_program(c, tile_rows=rows)
    assert prog.tile_rows == rows
    slabs_per_tile = sum(1 for kb in prog.kblocks if not (len(kb) == 7 and kb[6]))
    assert slabs_per_tile == (rows + 6) * (c // 64)          # rows + 6 slab loads per `rows` output rows instead of 7 * rows
    y = emulate_shift(prog, nhwc(x), slab.conv7_out_shift_weights(prog, w, dtype=torch.float32))
    assert torch.allclose(y.permute(0, 3, 1, 2), F.conv2d(x, w, padding=3).double(), atol=1e-3, rtol=1e-4)


def test_msb64_shift_program():
    slab = _slab_mod()
    torch.manual_seed(6)
    C, q, H, W = 64, 16, 9, 13
    x = torch.randn(1, C, H, W)
    ws = [torch.randn(q, C, k, k) for k in (1, 3, 3, 3)]
    ref = torch.cat([F.conv2d(x, w, padding=(w.shape[2] // 2) * d, dilation=d) for w, d in zip(ws, (1, 1, 2, 4))], 1)
    prog = slab.msb64_shift_program()
    y = emulate_shift(prog, nhwc(x), slab.msb64_shift_weights(ws, dtype=torch.float32))
    assert torch.allclose(y.permute(0, 3, 1, 2), ref.double(), atol=1e-3, rtol=1e-4)


def packed_convT_phases(w):
    """CPU statement of the packed phase layout [4][Cout][2][2][Cin] (ops.PACK_CONVT_PHASES): phase (ph, pw), tap (th, tw)
    holds W[ci, co, 1 - ph + 2 (1 - th), 1 - pw + 2 (1 - tw)]."""
    Cin, Cout = w.shape[:2]
    out = torch.zeros(4, Cout, 2, 2, Cin)
    for ph in range(2):
        for pw in range(2):
            for th in range(2):
                for tw in range(2):
                    out[ph * 2 + pw, :, th, tw, :] = w[:, :, 1 - ph + 2 * (1 - th), 1 - pw + 2 * (1 - tw)].t()
    return out.reshape(-1)


@pytest.mark.parametrize("Cin,Cout", [(64, 32), (128, 64)])
def test_convT_phase_programs_equal_conv_transpose(Cin, Cout):
    slab = _slab_mod()
    torch.manual_seed(7)
    H, W = 5, 6
    x = torch.randn(1, Cin, H, W)
    w = torch.randn(Cin, Cout, 4, 4)
    progs = slab.convT_phase_programs(Cin, Cout)
    wsl = slab.convT_phase_weight_slabs(progs, packed_convT_phases(w), Cin, Cout)
    y = torch.zeros(1, 2 * H, 2 * W, Cout, dtype=torch.float64)
    for (ph, pw, prog), ws_ in zip(progs, wsl):
        y += emulate_slab(prog, nhwc(x), ws_, out_stride=2, off=(ph, pw))
    assert torch.allclose(y.permute(0, 3, 1, 2), F.conv_transpose2d(x, w, stride=2, padding=1).double(), atol=1e-3, rtol=1e-4)


def test_descriptor_structs_match_the_header():
    """sizeof / offsetof of msg_slab_desc and msg_shift_desc as gcc sees include/msg_b200.h == the ctypes mirrors."""
    from multi_style_transfer_gan_b200._lib import ShiftDesc, SlabDesc
    fields = {"msg_slab_desc": (SlabDesc, ["Co_total", "out_stride", "out_off_w", "Ntot", "n_chains", "flags", "n_taps", "kb_dy",
                                           "kb_tap_begin", "tap_sx", "tap_kstep"]),
              "msg_shift_desc": (ShiftDesc, ["Co_total", "Ntot", "flags", "n_terms", "kb_dy", "kb_first", "grp_col0",
                                             "grp_term_begin", "term_col", "tile_rows", "kb_same_slab", "grp_row"])}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "msg_b200.h"', 'int main(void) {']
    for name, (_, fs) in fields.items():
        lines.append(f'  printf("{name} %zu\\n", sizeof({name}));')
        for f in fs:
            lines.append(f'  printf("{name}.{f} %zu\\n", offsetof({name}, {f}));')
    lines += ['  return 0;', '}']
    with tempfile.TemporaryDirectory() as td:
        src, exe = os.path.join(td, "t.c"), os.path.join(td, "t")
        open(src, "w").write("\n".join(lines))
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    got = dict(l.split() for l in out.strip().splitlines())
    for name, (cls, fs) in fields.items():
        assert int(got[name]) == ctypes.sizeof(cls), name
        for f in fs:
            assert int(got[f"{name}.{f}"]) == getattr(cls, f).offset, f"{name}.{f}"
