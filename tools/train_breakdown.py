"""Per-geometry CUDA-event breakdown of one EnhancedCycleGAN.train_step (c=64, batch 8 at 256x256, bf16)."""
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from multi_style_transfer_gan_b200 import _lib
    from multi_style_transfer_gan_b200.enhanced_train import EnhancedCycleGAN
    torch.manual_seed(0)
    m = EnhancedCycleGAN(channels=64, num_transformer_blocks=3, precision="bf16")
    A = torch.rand(8, 3, 256, 256) * 2 - 1
    B = torch.rand(8, 3, 256, 256) * 2 - 1
    for _ in range(3):
        m.train_step(A, B)
    torch.cuda.synchronize()
    recs = []
    orig_call = _lib.call

    def call(name, *args):
        key = name
        d = getattr(args[0], "_obj", None) if args else None
        if d is not None and hasattr(d, "KH"):
            key = f"{name} Cin={d.Cin} Cout={d.Cout} k={d.KH} s={d.in_stride} os={d.out_stride} plane={d.Hg}x{d.Wg} N={d.N}"
        elif d is not None and hasattr(d, "n_taps"):
            key = f"{name} Cin={d.Cin} Ntot={d.Ntot} taps={d.n_taps} plane={d.H}x{d.W}"
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = orig_call(name, *args)
        e1.record()
        recs.append((key, e0, e1))
        return r

    _lib.call = call
    reps = 3
    for _ in range(reps):
        m.train_step(A, B)
    torch.cuda.synchronize()
    _lib.call = orig_call
    agg = collections.OrderedDict()
    for k, e0, e1 in recs:
        v = agg.setdefault(k, [0.0, 0])
        v[0] += e0.elapsed_time(e1)
        v[1] += 1
    tot = sum(v[0] for v in agg.values()) / reps
    print(f"one train step: {tot:.2f} ms in kernels, {len(recs) // reps} launches")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:45]:
        print(f"{v[0] / reps:8.3f} ms  x{v[1] // reps:<4d} {k}")


if __name__ == "__main__":
    main()
