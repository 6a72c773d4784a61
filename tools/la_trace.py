"""Role timelines of the fused LocalAttention stage kernel (development tool).  Needs a trace build:
    MSG_LA_TRACE=1 python -m multi_style_transfer_gan_b200.build --force
Usage: python tools/la_trace.py C H W [N] [tile]   -> prints the events of CTA 0 for one steady-state tile."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multi_style_transfer_gan_b200 import _lib, ops  # noqa: E402

ROLES = ["producer", "issuerG", "issuerA", "drain0", "drain1", "drain2", "drain3", "sm0.0", "sm0.1", "sm0.2", "sm0.3",
         "sm1.0", "sm1.1", "sm1.2", "sm1.3", "sm2.0", "sm2.1", "sm2.2", "sm2.3", "epi0", "epi1", "epi2", "epi3", "issuerPV"]
EV = {"producer": {1: "x load issue"},
      "issuerG": {1: "x ready", 2: "chunk slot free", 3: "chunk issued", 10: "proj start", 11: "proj issued"},
      "issuerA": {1: "qk ready", 2: "S buf free -> issue S", 4: "P ready -> issue PV"},
      "issuerPV": {4: "P ready -> issue PV"},
      "drain": {1: "acc full", 2: "pass1 done", 3: "windows free", 4: "chunk stored"},
      "sm": {1: "S full", 2: "P stored", 3: "O full", 4: "O drained"},
      "epi": {3: "proj full", 4: "store done"}}


def main():
    C, H, W = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    N = int(sys.argv[4]) if len(sys.argv) > 4 else 16
    tile = int(sys.argv[5]) if len(sys.argv) > 5 else 3
    torch.manual_seed(0)
    x = torch.randn(N, H, W, C, device="cuda").bfloat16()
    wq = (torch.randn(3 * C * C, device="cuda") * (2.0 / C) ** 0.5).bfloat16()
    wp = (torch.randn(C * C, device="cuda") * (1.0 / C) ** 0.5).bfloat16()
    bq, bp = torch.randn(3 * C, device="cuda") * 0.1, torch.randn(C, device="cuda") * 0.1
    st = ops.instnorm_stats(x)
    for _ in range(2):
        ops.la_stage_fwd(x, wq, bq, wp, bp, in_stats=st, in_act=ops.ACT_RELU)
    buf = torch.zeros(24 * 4096, device="cuda", dtype=torch.int64)
    _lib.load().msg_la_stage_set_trace(ctypes.c_void_p(buf.data_ptr()))
    ops.la_stage_fwd(x, wq, bq, wp, bp, in_stats=st, in_act=ops.ACT_RELU)
    torch.cuda.synchronize()
    _lib.load().msg_la_stage_set_trace(None)
    b = buf.cpu().view(24, 4096)
    evs = []
    for r, name in enumerate(ROLES):
        n = int(b[r, 4094])
        kind = "drain" if name.startswith("drain") else "sm" if name.startswith("sm") else "epi" if name.startswith("epi") else name
        for i in range(n):
            tag, clk = int(b[r, 2 * i]), int(b[r, 2 * i + 1])
            ev, lt, u = tag >> 32, (tag >> 8) & 0xffffff, tag & 0xff
            evs.append((clk, name, EV[kind].get(ev, str(ev)), lt, u))
    evs.sort()
    if not evs:
        print("no events: is this a MSG_LA_TRACE build?")
        return
    t0 = min(c for c, n, e, lt, u in evs if lt == tile)
    per_tile = {}
    for c, n, e, lt, u in evs:
        if n == "issuerA" and e == "S buf free -> issue S" and u == 0:
            per_tile[lt] = c
    ks = sorted(per_tile)
    print("tiles of CTA 0:", len(ks), " cycles between 'qk ready' of consecutive tiles:", [per_tile[b_] - per_tile[a] for a, b_ in zip(ks, ks[1:])][:12])
    only = {"producer", "issuerG", "issuerA", "issuerPV", "drain0", "drain3", "sm0.0", "sm1.0", "sm2.0", "epi0"}
    for c, n, e, lt, u in evs:
        if lt in (tile, tile + 1) and n in only:
            print(f"{c - t0:8d}  {n:9s} tile {lt} unit {u}  {e}")


if __name__ == "__main__":
    main()
