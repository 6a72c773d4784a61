"""Times the MultiScaleBlock's 1x1 fusion conv with the fused input InstanceNorm + ReLU (conv_tma_kernel<true>) at the three widths of
the c = 64 generator, with its HBM floor (input + output bytes at the measured copy bandwidth).  Usage: python tools/fusion_bench.py [N]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_style_transfer_gan_b200 import ops  # noqa: E402


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


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    for C, H in ((64, 512), (128, 256), (256, 128)):
        x = torch.randn(N, H, H, C, device="cuda").bfloat16()
        w = torch.randn(C, C, 1, 1, device="cuda") * (1.0 / C) ** 0.5
        bias = torch.randn(C, device="cuda") * 0.1
        g = ops.ConvGeom("conv", C, C, 1)
        wp = g.pack_fwd(w, torch.bfloat16)
        sti = ops.instnorm_stats(x)
        st = ops.new_stats(N, C, "cuda")
        out = torch.empty_like(x)
        ms_f = timed(lambda: g.forward(x, wp, bias, out=out, stats=st, in_stats=sti, in_act=ops.ACT_RELU))
        ms_p = timed(lambda: g.forward(x, wp, bias, out=out, stats=st))
        gb = 2 * x.numel() * 2 / 1e9
        print(f"1x1 {C}->{C} at {N}x{H}x{H}: fused IN+ReLU on load {ms_f:.4f} ms ({gb / ms_f:.2f} TB/s), plain {ms_p:.4f} ms ({gb / ms_p:.2f} TB/s); "
              f"in + out = {gb:.3f} GB")


if __name__ == "__main__":
    main()
