"""Does the residual IN apply (two read streams + one write) depend on the relative placement of its tensors?
    python tools/in_apply_skew.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_style_transfer_gan_b200 import ops  # noqa: E402


def main():
    N, H, W, C = 16, 512, 512, 64
    n = N * H * W * C
    pool = torch.empty(3 * n + (64 << 20), device="cuda", dtype=torch.bfloat16)
    x = pool[:n].view(N, H, W, C)
    x.normal_()
    st = ops.new_stats(N, C, x.device)
    st[..., 0] = 0.0
    st[..., 1] = float(H * W)
    flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
    for skew in (0, 64, 1024, 4096, 64 * 1024, 1 << 20, (1 << 20) + 4096, 3 << 20, 17 << 20):
        r = pool[n + skew:2 * n + skew].view(N, H, W, C)
        r.normal_()
        for inplace in (True, False):
            out = x if inplace else pool[2 * n + (32 << 20):3 * n + (32 << 20)].view(N, H, W, C)
            ts = []
            for _ in range(5):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ops.instnorm_apply(x, st, ops.ACT_RELU, residual=r, out=out)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = sorted(ts)[2]
            print(f"residual skew {skew * 2:>10d} B  in-place {int(inplace)}: {ms:.3f} ms  {3 * n * 2 / ms / 1e9:.2f} TB/s")
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.instnorm_apply(x, st, ops.ACT_RELU, out=x)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[2]
    print(f"no residual, in-place: {ms:.3f} ms  {2 * n * 2 / ms / 1e9:.2f} TB/s")


if __name__ == "__main__":
    main()
