"""Times the C = 64 MultiScaleBlock branch kernels on the bench geometry (16 x 512 x 512 x 64 bf16): the row-ring kernel with /
without IN statistics, and the per-tap row-slab kernel.  Usage: python tools/ring_bench.py [N H W]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_style_transfer_gan_b200 import ops, slab  # noqa: E402


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def bench128(N, H, W):
    C, Q = 128, 32
    x = torch.randn(N, H, W, C, device="cuda").bfloat16()
    ws = [torch.randn(Q, C, k, k, device="cuda") * 0.05 for k in (1, 3, 3, 3)]
    bias = torch.randn(C, device="cuda") * 0.1
    out = torch.empty_like(x)
    st = ops.new_stats(N, C, "cuda")
    wr = slab.msb_ring_weights(ws, C)
    print(f"C=128 {N}x{H}x{W} ring (3 passes), stats : {timed(lambda: slab.msb_ring(x, wr, bias, C, out=out, stats=st)):.4f} ms")
    prog = slab.msb_program(C)
    wsl = slab.msb_weight_slab(prog, ws)
    print(f"C=128 per-tap slab                       : {timed(lambda: slab.conv_slab(prog, x, wsl, bias, out=out, stats=st)):.4f} ms")


def bench_convt(N, H, W, Cin=128, Cout=64):
    x = torch.randn(N, H, W, Cin, device="cuda").bfloat16()
    w = torch.randn(Cin, Cout, 4, 4, device="cuda") * 0.05
    bias = torch.randn(Cout, device="cuda") * 0.1
    out = torch.empty(N, 2 * H, 2 * W, Cout, device="cuda", dtype=torch.bfloat16)
    st = ops.new_stats(N, Cout, "cuda")
    wr = slab.convt_ring_weights(w)
    print(f"convT {Cin}->{Cout} {N}x{H}x{W} ring (2 phases), stats : {timed(lambda: slab.convt_ring(x, wr, bias, Cout, out=out, stats=st)):.4f} ms")
    g = ops.ConvGeom("convT", Cin, Cout, 4, 2, 1)
    progs = slab.convT_phase_programs(Cin, Cout)
    wsl = slab.convT_phase_weight_slabs(progs, g.pack_fwd(w, torch.bfloat16), Cin, Cout)
    print(f"convT row-slab phases (4 launches)              : {timed(lambda: slab.convT_slab(progs, x, wsl, bias, out, stats=st)):.4f} ms")


def bench_out7(N, H, W):
    f = torch.randn(N, H, W, 64, device="cuda").bfloat16()
    a1 = torch.randn(N, H, W, 64, device="cuda").bfloat16()
    w = torch.randn(3, 64, 7, 7, device="cuda") * 0.02
    bias = torch.randn(3, device="cuda") * 0.1
    st = ops.instnorm_stats(f)
    y = torch.empty(N, 3, H, W, device="cuda")
    wr = slab.out7_ring_weights(w)
    print(f"output layer {N}x{H}x{W}: fused ring (apply + conv7 + tanh) : {timed(lambda: slab.out7_ring(f, st, a1, wr, bias, nchw_out=y)):.4f} ms")
    prog = slab.conv7_out_shift_program(64)
    wsl = slab.conv7_out_shift_weights(prog, w)
    a2 = torch.empty_like(f)

    def two():
        ops.instnorm_apply(f, st, ops.ACT_RELU, residual=a1, out=a2)
        slab.conv_shift(prog, a2, wsl, bias, act=ops.ACT_TANH, nchw_out=y)
    print(f"apply kernel + taps-as-N conv (2 launches)                : {timed(two):.4f} ms")


def bench_down(N, H, W):
    x = torch.randn(N, H, W, 64, device="cuda").bfloat16()
    w = torch.randn(128, 64, 4, 4, device="cuda") * 0.03
    bias = torch.randn(128, device="cuda") * 0.1
    sti = ops.instnorm_stats(x)
    st = ops.new_stats(N, 128, "cuda")
    out = torch.empty(N, H // 2, W // 2, 128, device="cuda", dtype=torch.bfloat16)
    wr = slab.down_ring_weights(w)
    print(f"down1 {N}x{H}x{W}: fused ring (IN + ReLU + conv 4x4 s2, 2 launches) : {timed(lambda: slab.down_ring(x, sti, wr, bias, 128, out=out, stats=st)):.4f} ms")
    print(f"      ring without the fused norm                              : {timed(lambda: slab.down_ring(x, None, wr, bias, 128, out=out, stats=st)):.4f} ms")
    g = ops.ConvGeom("conv", 64, 128, 4, 2, 1)
    wp = g.pack_fwd(w, torch.bfloat16)
    a = torch.empty_like(x)

    def two():
        ops.instnorm_apply(x, sti, ops.ACT_RELU, out=a)
        g.forward(a, wp, bias, out=out, stats=st)
    print(f"      apply kernel + per-tap implicit GEMM (2 launches)        : {timed(two):.4f} ms")


def main():
    N, H, W = [int(a) for a in sys.argv[1:4]] if len(sys.argv) >= 4 else (16, 512, 512)
    bench_down(N, H, W)
    bench_out7(N, H, W)
    bench_convt(N, H // 2, W // 2)
    bench128(N, H // 2, W // 2)
    torch.manual_seed(0)
    x = torch.randn(N, H, W, 64, device="cuda").bfloat16()
    ws = [torch.randn(16, 64, k, k, device="cuda") * 0.05 for k in (1, 3, 3, 3)]
    bias = torch.randn(64, device="cuda") * 0.1
    out = torch.empty_like(x)
    st = ops.new_stats(N, 64, "cuda")
    wr = slab.msb64_ring_weights(ws)
    print(f"ring, stats    : {timed(lambda: slab.msb64_ring(x, wr, bias, out=out, stats=st)):.4f} ms")
    print(f"ring, no stats : {timed(lambda: slab.msb64_ring(x, wr, bias, out=out)):.4f} ms")
    prog = slab.msb_program(64)
    wsl = slab.msb_weight_slab(prog, ws)
    print(f"per-tap slab   : {timed(lambda: slab.conv_slab(prog, x, wsl, bias, out=out, stats=st)):.4f} ms")


if __name__ == "__main__":
    main()
