"""Launches the fused MultiScaleBlock branch kernel (C=64 at 512^2, C=128 at 256^2) and the four convT phases
128->64 once each after a warm-up: a small target for ncu source-level captures.
    python tools/prof_msb.py [--batch 16]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_style_transfer_gan_b200 import ops, slab  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--reps", type=int, default=2)
    a = ap.parse_args()
    N = a.batch
    torch.manual_seed(0)
    for C, S in ((64, 512), (128, 256)):
        q = C // 4
        prog = slab.msb_program(C)
        ws = [torch.randn(q, C, 1, 1, device="cuda") * 0.05] + [torch.randn(q, C, 3, 3, device="cuda") * 0.05 for _ in range(3)]
        wsl = slab.msb_weight_slab(prog, ws)
        bias = torch.zeros(C, device="cuda")
        x = torch.randn(N, S, S, C, device="cuda").bfloat16()
        out = torch.empty_like(x)
        for _ in range(a.reps):
            st = ops.new_stats(N, C, x.device)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            slab.conv_slab(prog, x, wsl, bias, out=out, stats=st)
            e1.record()
            torch.cuda.synchronize()
        print(f"MSB branches C={C} @{S}: {e0.elapsed_time(e1):.3f} ms")
    g = ops.ConvGeom("convT", 128, 64, 4, 2, 1)
    w = torch.randn(128, 64, 4, 4, device="cuda") * 0.05
    wp = g.pack_fwd(w, torch.bfloat16)
    x = torch.randn(N, 256, 256, 128, device="cuda").bfloat16()
    bias = torch.zeros(64, device="cuda")
    for _ in range(a.reps):
        st = ops.new_stats(N, 64, x.device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.forward(x, wp, bias, stats=st)
        e1.record()
        torch.cuda.synchronize()
    print(f"convT 128->64 @256 (4 phases): {e0.elapsed_time(e1):.3f} ms")


if __name__ == "__main__":
    main()
