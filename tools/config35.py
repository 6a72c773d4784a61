"""BASELINE configs 3 and 5 (not bench lines): Gram style loss fwd+bwd on the five VGG-19 tap shapes at batch 16,
256x256, and the 1024x1024 batch-8 four-style stylisation (peak memory, throughput)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def timed(fn, it=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it


def main():
    from multi_style_transfer_gan_b200 import ops
    from multi_style_transfer_gan_b200.enhanced_generator import EnhancedGenerator
    from multi_style_transfer_gan_b200.style_loss import GramStyleLoss, VGG19Features
    from multi_style_transfer_gan_b200.stylize import MultiStyleStylizer
    dev = "cuda"
    torch.manual_seed(0)
    shapes = [(64, 256), (128, 128), (256, 64), (512, 32), (512, 16)]
    feats = [torch.relu(torch.randn(16, s, s, c, device=dev)).to(torch.bfloat16) for c, s in shapes]
    tgts = [ops.gram(f) for f in feats]

    def gram_fb():
        loss = torch.zeros(1, device=dev)
        for f, t in zip(feats, tgts):
            _, g = ops.gram_loss_fwd(f, t, 1.0, loss)
            ops.gram_loss_bwd(f, g, t, 1.0)

    ms = timed(gram_fb)
    fl = 2 * sum(2.0 * c * c * s * s * 16 for c, s in shapes)
    by = sum(f.numel() * 2 * 3 for f in feats)
    print(f"config 3: Gram + loss fwd+bwd, 5 taps, batch 16 @256^2: {ms:.3f} ms  ({fl / ms / 1e9:.1f} TFLOP/s algorithmic, "
          f"{by / ms / 1e6:.0f} GB/s feature traffic)")
    vgg = VGG19Features(dev, seed=0)
    sl = GramStyleLoss(vgg, "bf16").set_style(torch.rand(16, 3, 256, 256, device=dev) * 2 - 1)
    y = (torch.rand(16, 3, 256, 256, device=dev) * 2 - 1).requires_grad_(True)

    def full():
        y.grad = None
        sl(y).backward()

    ms = timed(full, 5)
    print(f"config 3 with the VGG-19 trunk (fwd + bwd to the image): {ms:.2f} ms per batch of 16 = {16e3 / ms:.0f} images/s")

    gens = []
    for s in range(4):
        torch.manual_seed(s)
        gens.append(EnhancedGenerator(64, 3).to(dev))
    st = MultiStyleStylizer(gens, precision="bf16", micro_batch=4)
    x = torch.rand(8, 3, 1024, 1024, device=dev) * 2 - 1
    torch.cuda.reset_peak_memory_stats()
    ms = timed(lambda: st(x, [0.4, 0.3, 0.2, 0.1]), 3)
    print(f"config 5: 1024^2, batch 8, 4 styles (c=64): {ms:.1f} ms per batch = {8e3 / ms:.1f} images/s, peak memory "
          f"{torch.cuda.max_memory_allocated() / 2 ** 30:.2f} GiB")


if __name__ == "__main__":
    main()
