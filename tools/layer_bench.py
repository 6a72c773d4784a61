"""Per-layer micro-benchmark of the c=64 generator's kernels (CUDA events, L2 flushed between reps).
    python tools/layer_bench.py [--batch 4] [--size 512]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_style_transfer_gan_b200 import ops  # noqa: E402


def time_fn(fn, reps=5):
    flush = torch.empty(256 * 1024 * 1024, device="cuda", dtype=torch.uint8)
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    N, S = a.batch, a.size
    dt = torch.bfloat16
    dev = "cuda"
    c = 64
    layers = [
        ("initial 7x7 8->64", "conv", 8, c, 7, 1, 3, 1, S, True),
        ("down1.0 4x4s2 64->128", "conv", c, 2 * c, 4, 2, 1, 1, S, True),
        ("down2.0 4x4s2 128->256", "conv", 2 * c, 4 * c, 4, 2, 1, 1, S // 2, True),
        ("qkv 1x1 128->384 @S/2", "conv", 2 * c, 6 * c, 1, 1, 0, 1, S // 2, False),
        ("qkv 1x1 256->768 @S/4", "conv", 4 * c, 12 * c, 1, 1, 0, 1, S // 4, False),
        ("qkv 1x1 64->192 @S", "conv", c, 3 * c, 1, 1, 0, 1, S, False),
        ("proj 1x1 128->128 @S/2", "conv", 2 * c, 2 * c, 1, 1, 0, 1, S // 2, True),
        ("proj 1x1 64->64 @S", "conv", c, c, 1, 1, 0, 1, S, True),
        ("branch 3x3d2 128->32 @S/2", "conv", 2 * c, c // 2, 3, 1, 2, 2, S // 2, True),
        ("branch 3x3d2 256->64 @S/4", "conv", 4 * c, c, 3, 1, 2, 2, S // 4, True),
        ("branch 3x3d4 64->16 @S", "conv", c, c // 4, 3, 1, 4, 4, S, True),
        ("up1.0 convT 256->128 @S/4", "convT", 4 * c, 2 * c, 4, 2, 1, 1, S // 4, True),
        ("up2.0 convT 128->64 @S/2", "convT", 2 * c, c, 4, 2, 1, 1, S // 2, True),
        ("output 7x7 64->3(8) @S", "conv", c, 8, 7, 1, 3, 1, S, False),
    ]
    print(f"batch {N}, {S}x{S}; per launch: ms, TFLOP/s (algorithmic), GB/s (in+out bytes)")
    for name, kind, Cin, Cout, k, s, p, d, H, stats in layers:
        if a.only and a.only not in name:
            continue
        g = ops.ConvGeom(kind, Cin, Cout, k, s, p, d)
        x = torch.randn(N, H, H, Cin, device=dev).to(dt)
        w = (torch.randn(*g.weight_shape(), device=dev) * 0.05)
        wp = g.pack_fwd(w, dt)
        Ho, Wo = g.out_hw(H, H)
        out = torch.empty(N, Ho, Wo, Cout, device=dev, dtype=dt)
        st = ops.new_stats(N, Cout, dev) if stats else None
        ms = time_fn(lambda: g.forward(x, wp, None, out=out, stats=st))
        if kind == "convT":
            fl = 2.0 * N * Ho * Wo * Cout * Cin * 4
        else:
            fl = 2.0 * N * Ho * Wo * Cout * Cin * k * k
        by = (x.numel() + out.numel()) * 2
        print(f"{name:32s} {ms:8.3f} ms  {fl / ms / 1e9:8.1f} TF/s  {by / ms / 1e6:8.0f} GB/s")
    if a.only == "slabin":
        from multi_style_transfer_gan_b200 import slab
        x8 = torch.randn(N, S, S, 8, device=dev).to(dt)
        prog = slab.conv7_in_program(c)
        wsl = slab.conv7_in_weight_slab(prog, torch.randn(c, 3, 7, 7, device=dev) * 0.1)
        out = torch.empty(N, S, S, c, device=dev, dtype=dt)
        st = ops.new_stats(N, c, dev)
        ms = time_fn(lambda: slab.conv_slab(prog, x8, wsl, None, out=out, stats=st))
        print(f"{'slab input 7x7 3(8)->64 @S':32s} {ms:8.3f} ms  {2.0 * N * S * S * c * 3 * 49 / ms / 1e9:8.1f} TF/s  {(x8.numel() + out.numel()) * 2 / ms / 1e6:8.0f} GB/s")
    if not a.only or a.only == "slab":
        from multi_style_transfer_gan_b200 import slab
        for C, H in ((64, S), (128, S // 2), (256, S // 4)):
            x = torch.randn(N, H, H, C, device=dev).to(dt)
            prog = slab.msb_program(C)
            ws = [torch.randn(C // 4, C, k, k, device=dev) * 0.05 for k in (1, 3, 3, 3)]
            wsl = slab.msb_weight_slab(prog, ws)
            out = torch.empty(N, H, H, C, device=dev, dtype=dt)
            st = ops.new_stats(N, C, dev)
            ms = time_fn(lambda: slab.conv_slab(prog, x, wsl, None, out=out, stats=st))
            fl = 2.0 * N * H * H * (C // 4) * C * 28
            print(f"{'slab MSB branches C=%d @%d' % (C, H):32s} {ms:8.3f} ms  {fl / ms / 1e9:8.1f} TF/s  {2 * x.numel() * 2 / ms / 1e6:8.0f} GB/s")
        x = torch.randn(N, S, S, c, device=dev).to(dt)
        prog = slab.conv7_out_program(c)
        wsl = slab.conv7_out_weight_slab(prog, torch.randn(3, c, 7, 7, device=dev) * 0.02)
        y = torch.empty(N, 3, S, S, device=dev)
        ms = time_fn(lambda: slab.conv_slab(prog, x, wsl, None, act=3, nchw_out=y))
        print(f"{'slab output 7x7 64->3 @S':32s} {ms:8.3f} ms  {2.0 * N * S * S * 3 * c * 49 / ms / 1e9:8.1f} TF/s  {(x.numel() * 2 + y.numel() * 4) / ms / 1e6:8.0f} GB/s")
        x = torch.randn(N, S, S, c, device=dev).to(dt)
        prog = slab.conv7_out_shift_program(c)
        wsl = slab.conv7_out_shift_weights(prog, torch.randn(3, c, 7, 7, device=dev) * 0.02)
        ms = time_fn(lambda: slab.conv_shift(prog, x, wsl, None, act=3, nchw_out=y))
        print(f"{'shift output 7x7 64->3 @S':32s} {ms:8.3f} ms  {2.0 * N * S * S * 3 * c * 49 / ms / 1e9:8.1f} TF/s  {(x.numel() * 2 + y.numel() * 4) / ms / 1e6:8.0f} GB/s")
        prog = slab.msb64_shift_program()
        wsl = slab.msb64_shift_weights([torch.randn(16, 64, k, k, device=dev) * 0.05 for k in (1, 3, 3, 3)])
        out = torch.empty(N, S, S, 64, device=dev, dtype=dt)
        st = ops.new_stats(N, 64, dev)
        ms = time_fn(lambda: slab.conv_shift(prog, x, wsl, None, out=out, stats=st))
        print(f"{'shift MSB branches C=64 @S':32s} {ms:8.3f} ms  {2.0 * N * S * S * 16 * 64 * 28 / ms / 1e9:8.1f} TF/s  {2 * x.numel() * 2 / ms / 1e6:8.0f} GB/s")
        x8 = torch.randn(N, S, S, 8, device=dev).to(dt)
        prog = slab.conv7_in_program(c)
        wsl = slab.conv7_in_weight_slab(prog, torch.randn(c, 3, 7, 7, device=dev) * 0.1)
        out = torch.empty(N, S, S, c, device=dev, dtype=dt)
        st = ops.new_stats(N, c, dev)
        ms = time_fn(lambda: slab.conv_slab(prog, x8, wsl, None, out=out, stats=st))
        print(f"{'slab input 7x7 3(8)->64 @S':32s} {ms:8.3f} ms  {2.0 * N * S * S * c * 3 * 49 / ms / 1e9:8.1f} TF/s  {(x8.numel() + out.numel()) * 2 / ms / 1e6:8.0f} GB/s")
    if not a.only or "norm" in a.only:
        for C, H in ((64, S), (128, S // 2), (256, S // 4)):
            x = torch.randn(N, H, H, C, device=dev).to(dt)
            st = ops.instnorm_stats(x)
            y = torch.empty_like(x)
            ms = time_fn(lambda: ops.instnorm_apply(x, st, 1, out=y))
            print(f"{'instnorm_apply C=%d @%d' % (C, H):32s} {ms:8.3f} ms  {'':8s}       {2 * x.numel() * 2 / ms / 1e6:8.0f} GB/s")
            r = torch.randn_like(x)
            ms = time_fn(lambda: ops.instnorm_apply(x, st, 1, residual=r, out=y))
            print(f"{'instnorm_apply+res C=%d @%d' % (C, H):32s} {ms:8.3f} ms  {'':8s}       {3 * x.numel() * 2 / ms / 1e6:8.0f} GB/s")
            st2 = ops.new_stats(N, C, dev)
            ms = time_fn(lambda: ops.instnorm_stats(x, st2))
            print(f"{'instnorm_stats C=%d @%d' % (C, H):32s} {ms:8.3f} ms  {'':8s}       {x.numel() * 2 / ms / 1e6:8.0f} GB/s")
    if not a.only or "attn" in a.only:
        for C, H in ((64, S), (128, S // 2), (256, S // 4)):
            qkv = torch.randn(N, H, H, 3 * C, device=dev).to(dt)
            ms = time_fn(lambda: ops.local_attn_fwd(qkv))
            fl = 2.0 * 2 * N * (H // 4) ** 2 * C * C * 16
            print(f"{'local_attn C=%d @%d' % (C, H):32s} {ms:8.3f} ms  {fl / ms / 1e9:8.1f} TF/s  {qkv.numel() * 2 * 4 / 3 / ms / 1e6:8.0f} GB/s")
            dy = torch.randn(N, H, H, C, device=dev).to(dt)
            ms = time_fn(lambda: ops.local_attn_bwd(qkv, dy))
            print(f"{'local_attn bwd C=%d @%d' % (C, H):32s} {ms:8.3f} ms  {'':8s}       {qkv.numel() * 2 * 7 / 3 / ms / 1e6:8.0f} GB/s")


if __name__ == "__main__":
    main()
