"""Batch-independence probe: generator on odd planes (batch 2 vs image 0 alone) with the fused LocalAttention stage on / off,
and the stage kernel itself on partial tiles.  Usage: python tools/batch_indep.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_style_transfer_gan_b200 import ops
from multi_style_transfer_gan_b200.enhanced_generator import EnhancedGenerator

torch.manual_seed(5)
G = EnhancedGenerator(64, 3).cuda().eval().set_precision("bf16")
for hw in [(80, 48), (192, 320), (128, 384)]:
    g = torch.Generator().manual_seed(9)
    x = (torch.rand(2, 3, *hw, generator=g) * 2 - 1).cuda()
    for fuse in (True, False):
        G._engine.fuse_la = fuse
        with torch.no_grad():
            y2 = G(x)
            y2b = G(x)
            y1 = G(x[:1].contiguous())
        d = (y1 - y2[:1]).float()
        print(hw, "fuse_la", fuse, "rerun diff", float((y2 - y2b).abs().max()), "batch-1 vs batch-2 rel-l2", float(d.norm() / y2[:1].float().norm()),
              "max", float(d.abs().max()))

torch.manual_seed(0)
for C, N, H, W in [(64, 2, 80, 48), (128, 2, 40, 24), (128, 2, 8, 8), (64, 3, 12, 40)]:
    x = torch.randn(N, H, W, C, device="cuda").bfloat16()
    wq = (torch.randn(3 * C * C, device="cuda") * (2.0 / C) ** 0.5).bfloat16()
    wp = (torch.randn(C * C, device="cuda") * (1.0 / C) ** 0.5).bfloat16()
    bq = torch.randn(3 * C, device="cuda") * 0.1
    bp = torch.randn(C, device="cuda") * 0.1
    st = ops.instnorm_stats(x)
    o = [ops.la_stage_fwd(x, wq, bq, wp, bp, in_stats=st, in_act=ops.ACT_RELU) for _ in range(3)]
    o1 = ops.la_stage_fwd(x[:1].contiguous(), wq, bq, wp, bp, in_stats=st[:1].contiguous(), in_act=ops.ACT_RELU)
    torch.cuda.synchronize()
    print(C, N, H, W, "rerun", float((o[0].float() - o[1].float()).abs().max()), float((o[0].float() - o[2].float()).abs().max()),
          "single vs batch", float((o1.float() - o[0][:1].float()).abs().max()))

# ---- which launch is the first to depend on the batch?  (records image 0 of every op's output)
from multi_style_transfer_gan_b200 import slab  # noqa: E402
import multi_style_transfer_gan_b200.generator_engine as ge  # noqa: E402

rec = []


def wrap(mod, name):
    f = getattr(mod, name)

    def w(*a, **k):
        r = f(*a, **k)
        out = k.get("out") if k.get("out") is not None else k.get("nchw_out") if k.get("nchw_out") is not None else r
        st = k.get("stats")
        rec.append((name, tuple(out.shape), out[:1].clone(), None if st is None else st[:1].clone()))
        return r
    setattr(mod, name, w)


for m, n in ((ops, "instnorm_apply"), (ops, "la_stage_fwd"), (ops, "local_attn_fwd"), (slab, "conv_slab"), (slab, "conv_shift"), (slab, "convT_slab"),
             (ops.ConvGeom, "forward")):
    wrap(m, n)
g = torch.Generator().manual_seed(9)
x = (torch.rand(2, 3, 80, 48, generator=g) * 2 - 1).cuda()
G._engine.fuse_la = True
with torch.no_grad():
    G(x)
    r2 = list(rec)
    rec.clear()
    G(x[:1].contiguous())
    r1 = list(rec)
for (n2, s2, o2, st2), (n1, s1, o1, st1) in zip(r2, r1):
    d = float((o2.float() - o1.float()).abs().max())
    ds = None if st2 is None else float(((st2 - st1).abs() / (st1.abs() + 1e-30)).max())
    print(f"{n2:16s} {str(s2):24s} out diff {d:.3e}  stats rel diff {ds}")
