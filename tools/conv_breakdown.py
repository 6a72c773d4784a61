"""Per-layer CUDA-event breakdown of one generator forward (bf16): groups the C-ABI calls by entry point and
conv geometry.  Usage: python tools/conv_breakdown.py [--batch 16] [--size 512]"""
import argparse
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--channels", type=int, default=64)
    a = ap.parse_args()
    from multi_style_transfer_gan_b200 import _lib
    from multi_style_transfer_gan_b200.enhanced_generator import EnhancedGenerator
    torch.manual_seed(0)
    g = EnhancedGenerator(channels=a.channels, num_transformer_blocks=3 if a.channels == 64 else 1).cuda().eval()
    g.set_precision("bf16")
    x = torch.rand(a.batch, 3, a.size, a.size, device="cuda") * 2 - 1
    with torch.no_grad():
        for _ in range(3):
            g(x)
    torch.cuda.synchronize()
    recs = []
    orig_call = _lib.call

    def call(name, *args):
        key = name
        d = getattr(args[0], "_obj", None) if args else None
        if d is not None:
            if hasattr(d, "KH"):
                key = f"{name} Cin={d.Cin} Cout={d.Cout} k={d.KH} s={d.in_stride} d={d.dil} plane={d.Hg}x{d.Wg}"
            elif hasattr(d, "n_taps"):
                key = f"{name} Cin={d.Cin} Ntot={d.Ntot} taps={d.n_taps} plane={d.H}x{d.W}"
            elif hasattr(d, "n_groups"):
                key = f"{name} Cin={d.Cin} Ntot={d.Ntot} kblocks={d.n_kblocks} plane={d.H}x{d.W}"
        if name == "msg_la_stage_fwd":
            key = f"{name} C={args[11]} plane={args[9]}x{args[10]}"
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = orig_call(name, *args)
        e1.record()
        recs.append((key, e0, e1))
        return r

    _lib.call = call
    import multi_style_transfer_gan_b200.ops as ops
    import multi_style_transfer_gan_b200.slab as slab
    reps = 5
    with torch.no_grad():
        for _ in range(reps):
            g(x)
    torch.cuda.synchronize()
    _lib.call = orig_call
    agg = collections.OrderedDict()
    for k, e0, e1 in recs:
        v = agg.setdefault(k, [0.0, 0])
        v[0] += e0.elapsed_time(e1)
        v[1] += 1
    tot = sum(v[0] for v in agg.values()) / reps
    print(f"one forward, batch {a.batch} at {a.size}^2: {tot:.3f} ms in kernels")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{v[0] / reps:8.3f} ms  x{v[1] // reps:<3d} {k}")


if __name__ == "__main__":
    main()
