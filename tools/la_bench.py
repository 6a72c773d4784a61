"""Times the fused LocalAttention stage kernel on the bench geometries (16 images of the 512x512 step).
Usage: python tools/la_bench.py [C ...]        (default: 64 128 256)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_style_transfer_gan_b200 import ops  # noqa: E402

GEO = {64: (16, 512, 512), 128: (16, 256, 256), 256: (16, 128, 128)}


def main():
    cs = [int(a) for a in sys.argv[1:]] or [64, 128, 256]
    reps = int(os.environ.get("REPS", "10"))
    torch.manual_seed(0)
    for C in cs:
        N, H, W = GEO[C]
        x = torch.randn(N, H, W, C, device="cuda").bfloat16()
        wq = (torch.randn(3 * C * C, device="cuda") * (2.0 / C) ** 0.5).bfloat16()
        wp = (torch.randn(C * C, device="cuda") * (1.0 / C) ** 0.5).bfloat16()
        bq, bp = torch.randn(3 * C, device="cuda") * 0.1, torch.randn(C, device="cuda") * 0.1
        st = ops.instnorm_stats(x)
        if not ops.la_stage_supported(x, wq, wp):
            print(f"C={C}: not supported by the fused stage kernel")
            continue
        for _ in range(2):
            ops.la_stage_fwd(x, wq, bq, wp, bp, in_stats=st, in_act=ops.ACT_RELU)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            ops.la_stage_fwd(x, wq, bq, wp, bp, in_stats=st, in_act=ops.ACT_RELU)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        by = 2 * x.numel() * 2
        fl = 2.0 * N * H * W * C * (3 * C + C) + 2 * 2.0 * (N * H * W / 16) * C * C * 16
        print(f"C={C} {N}x{H}x{W}: {ms:.4f} ms  {by / ms / 1e6:.0f} GB/s (in + out)  {fl / ms / 1e9:.0f} TFLOP/s")


if __name__ == "__main__":
    main()
