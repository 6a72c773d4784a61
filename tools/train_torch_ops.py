"""Which Python lines of the train step launch torch's own kernels (fills, adds, copies, cats)?  torch.profiler with stacks over one
eager train step; device kernels whose name starts with 'void at::' are attributed to the innermost frame inside this repository.
Usage: python tools/train_torch_ops.py"""
import collections
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from multi_style_transfer_gan_b200.enhanced_train import EnhancedCycleGAN
    torch.manual_seed(0)
    m = EnhancedCycleGAN(channels=64, num_transformer_blocks=3, precision="bf16", use_graph=False)
    A = torch.rand(8, 3, 256, 256) * 2 - 1
    B = torch.rand(8, 3, 256, 256) * 2 - 1
    for _ in range(2):
        m.train_step(A, B)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
        m.train_step(A, B)
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0, collections.Counter()])
    for ev in prof.events():
        if not ev.kernels:
            continue
        tk = [k for k in ev.kernels if k.name.startswith("void at::") or "at::native" in k.name]
        if not tk:
            continue
        site = next((s for s in ev.stack if "multi_style_transfer_gan_b200" in s or "/repo/" in s), ev.stack[0] if ev.stack else "?")
        site = site.replace(ROOT + "/", "")
        a = agg[site]
        a[0] += len(tk)
        a[1] += sum(k.duration for k in tk)
        a[2][ev.name] += len(tk)
    tot_n = sum(v[0] for v in agg.values())
    tot_t = sum(v[1] for v in agg.values())
    print(f"torch kernels in one eager train step: {tot_n} launches, {tot_t / 1e3:.2f} ms of device time")
    for site, (n, t, ops_) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"{n:5d} {t / 1e3:7.3f} ms  {site[:110]}  {dict(ops_.most_common(3))}")


if __name__ == "__main__":
    main()
