"""tcgen05 implicit-GEMM conv (conv_tc.cu) against the SIMT engine (conv_simt.cu) and against
PyTorch fp32, on the layer shapes of the c=64 generator (scaled-down planes)."""
import pytest
import torch
import torch.nn.functional as F

from tests.util import assert_parity

pytestmark = pytest.mark.gpu
DEV = "cuda"


def nhwc(x, dtype=torch.bfloat16):
    return x.permute(0, 2, 3, 1).contiguous().to(dtype)


def nchw(x):
    return x.float().permute(0, 3, 1, 2).contiguous()


LAYERS = [
    # kind, Cin, Cout, k, s, p, d, H, W   (c=64 generator layers)
    ("conv", 8, 64, 7, 1, 3, 1, 64, 64),      # initial (image padded to 8 channels)
    ("conv", 64, 128, 4, 2, 1, 1, 64, 64),    # down1.0
    ("conv", 128, 256, 4, 2, 1, 1, 32, 32),   # down2.0
    ("conv", 128, 384, 1, 1, 0, 1, 32, 32),   # qkv
    ("conv", 64, 192, 1, 1, 0, 1, 64, 64),    # qkv (C=64): BN=96 tiles
    ("conv", 256, 256, 1, 1, 0, 1, 16, 16),   # proj / fusion
    ("conv", 256, 64, 3, 1, 1, 1, 16, 16),    # branch2
    ("conv", 128, 32, 3, 1, 2, 2, 32, 32),    # branch3
    ("conv", 64, 16, 3, 1, 4, 4, 64, 64),     # branch4 (N=16)
    ("convT", 256, 128, 4, 2, 1, 1, 16, 16),  # up1.0
    ("convT", 128, 64, 4, 2, 1, 1, 32, 32),   # up2.0
    ("conv", 64, 8, 7, 1, 3, 1, 64, 64),      # output (padded filters)
    ("conv", 64, 128, 4, 2, 1, 1, 48, 80),    # plane not a multiple of 128 -> partial tiles
    ("conv", 64, 64, 3, 1, 1, 1, 16, 256),    # Wg = 256: two 128-pixel tiles per row
    ("conv", 64, 320, 1, 1, 0, 1, 32, 32),    # Cout > 256: two N tiles of 160
    ("conv", 128, 64, 4, 2, 1, 1, 32, 512),   # stride 2 with Wg = 256
]


@pytest.mark.parametrize("case", LAYERS, ids=lambda c: "-".join(map(str, c)))
def test_tc_vs_simt_and_torch(case):
    from multi_style_transfer_gan_b200 import ops
    kind, Cin, Cout, k, s, p, d, H, W = case
    torch.manual_seed(0)
    g = ops.ConvGeom(kind, Cin, Cout, k, s, p, d)
    N = 2
    dt = torch.bfloat16
    x = (torch.randn(N, Cin, H, W, device=DEV)).to(dt).float()
    w = (torch.randn(*g.weight_shape(), device=DEV) * (1.0 / (Cin * k * k) ** 0.5)).to(dt).float()
    b = torch.randn(Cout, device=DEV)
    ref = F.conv_transpose2d(x, w, b, stride=2, padding=1) if kind == "convT" else F.conv2d(x, w, b, stride=s, padding=p, dilation=d)
    wp = g.pack_fwd(w.contiguous(), dt)
    xh = nhwc(x)
    Ho, Wo = g.out_hw(H, W)
    use_stats = (H * W if kind == "convT" else Ho * Wo) % 128 == 0
    st_tc = ops.new_stats(N, Cout, DEV) if use_stats else None
    st_si = ops.new_stats(N, Cout, DEV) if use_stats else None
    y_tc = g.forward(xh, wp, b, stats=st_tc)                    # TMA kernel when the geometry allows, else gather
    y_si = g.forward(xh, wp, b, stats=st_si, extra_flags=ops.CONV_FORCE_SIMT)
    st_ga = ops.new_stats(N, Cout, DEV) if use_stats else None
    y_ga = g.forward(xh, wp, b, stats=st_ga, extra_flags=ops.CONV_FORCE_GATHER)   # cp.async gather kernel
    assert_parity(nchw(y_tc), ref, 1e-2, "tc vs torch")
    assert_parity(nchw(y_tc), nchw(y_si), 1e-2, "tc vs simt")
    assert_parity(nchw(y_ga), nchw(y_si), 1e-2, "gather vs simt")
    if use_stats:
        assert_parity(st_tc, st_si, 1e-4, "stats tc vs simt")
        assert_parity(st_ga, st_si, 1e-4, "stats gather vs simt")
    # fused activation epilogues
    y_act = g.forward(xh, wp, b, act=ops.ACT_LRELU)
    assert_parity(nchw(y_act), F.leaky_relu(ref, 0.2), 1e-2, "lrelu epilogue")


def test_tc_nchw_tanh_epilogue_and_accumulate():
    from multi_style_transfer_gan_b200 import ops
    torch.manual_seed(1)
    dt = torch.bfloat16
    N, C, H, W = 2, 64, 32, 32
    x = torch.randn(N, C, H, W, device=DEV).to(dt).float()
    w = (torch.randn(4, C, 7, 7, device=DEV) * 0.02).to(dt).float()
    w[3] = 0
    b = torch.randn(4, device=DEV)
    g = ops.ConvGeom("conv", C, 3, 7, 1, 3)
    y = torch.empty(N, 3, H, W, device=DEV)
    g.forward(nhwc(x), ops.pack_weight(w, ops.PACK_FWD, dt), b, act=ops.ACT_TANH, nchw_out=y)
    assert_parity(y, torch.tanh(F.conv2d(x, w[:3], b[:3], padding=3)), 1e-2, "nchw tanh")
    # accumulate flag: y += conv
    g2 = ops.ConvGeom("conv", C, 32, 3, 1, 1)
    w2 = (torch.randn(32, C, 3, 3, device=DEV) * 0.05).to(dt).float()
    base = torch.randn(N, 32, H, W, device=DEV).to(dt)
    out = nhwc(base.float())
    g2.forward(nhwc(x), g2.pack_fwd(w2, dt), None, out=out, extra_flags=ops.CONV_ACCUM)
    assert_parity(nchw(out), base.float() + F.conv2d(x, w2, None, padding=1), 1e-2, "accumulate")


# ---- row-slab kernel (csrc/conv_slab.cu) ---------------------------------------------------------
@pytest.mark.parametrize("C,H,W", [(64, 16, 128), (64, 8, 256), (128, 8, 128), (64, 12, 72), (256, 4, 128), (128, 7, 72), (128, 6, 200)])
def test_slab_msb_branches(C, H, W):
    """fused 1x1 + 3x3 dil 1/2/4 branches == the four separate convs + cat, incl. shared IN statistics"""
    from multi_style_transfer_gan_b200 import ops, slab
    torch.manual_seed(0)
    N, q, dt = 2, C // 4, torch.bfloat16
    x = torch.randn(N, C, H, W, device=DEV).to(dt).float()
    ws = [(torch.randn(q, C, k, k, device=DEV) * (1.0 / (C * k * k) ** 0.5)).to(dt).float() for k in (1, 3, 3, 3)]
    bs = [torch.randn(q, device=DEV) for _ in range(4)]
    ref = torch.cat([F.conv2d(x, w, b, padding=(w.shape[2] // 2) * d, dilation=d)
                     for w, b, d in zip(ws, bs, (1, 1, 2, 4))], 1)
    prog = slab.msb_program(C)
    wsl = slab.msb_weight_slab(prog, ws)
    st = ops.new_stats(N, C, DEV)
    y = slab.conv_slab(prog, nhwc(x), wsl, torch.cat(bs).contiguous(), stats=st)
    assert_parity(nchw(y), ref, 1e-2, "fused branches")
    st_ref = torch.stack([ref.sum((2, 3)), (ref * ref).sum((2, 3))], -1)
    assert_parity(st, st_ref, 3e-3, "stats")


@pytest.mark.parametrize("Cin,Cout,H,W", [(128, 64, 16, 128), (256, 128, 8, 64), (64, 32, 6, 40), (128, 64, 7, 136)])
def test_slab_convT_phases(Cin, Cout, H, W):
    """4x4 stride-2 transposed conv as four row-slab programs (strided output) == conv_transpose2d, incl. IN statistics;
    covers resident and streamed weights, one and two output rows per tile, ragged widths."""
    from multi_style_transfer_gan_b200 import ops, slab
    torch.manual_seed(0)
    N, dt = 2, torch.bfloat16
    x = torch.randn(N, Cin, H, W, device=DEV).to(dt).float()
    w = (torch.randn(Cin, Cout, 4, 4, device=DEV) * (1.0 / (Cin * 4) ** 0.5)).to(dt).float()
    b = torch.randn(Cout, device=DEV)
    ref = F.conv_transpose2d(x, w, b, stride=2, padding=1)
    g = ops.ConvGeom("convT", Cin, Cout, 4, 2, 1)
    progs = slab.convT_phase_programs(Cin, Cout)
    wsl = slab.convT_phase_weight_slabs(progs, g.pack_fwd(w, dt), Cin, Cout)
    st = ops.new_stats(N, Cout, DEV)
    y = torch.zeros(N, 2 * H, 2 * W, Cout, device=DEV, dtype=dt)
    slab.convT_slab(progs, nhwc(x), wsl, b, y, stats=st)
    assert_parity(nchw(y), ref, 1e-2, "convT phases")
    st_ref = torch.stack([ref.sum((2, 3)), (ref * ref).sum((2, 3))], -1)
    assert_parity(st, st_ref, 3e-3, "stats")
    # and against the per-tap implicit-GEMM kernel (same packed weights)
    y2 = g.forward(nhwc(x), g.pack_fwd(w, dt), b)
    assert_parity(y, y2.float(), 1e-2, "vs conv_tma phases")


@pytest.mark.parametrize("c,H,W", [(64, 16, 128), (64, 8, 384), (128, 8, 128), (64, 24, 40)])
def test_slab_output_conv(c, H, W):
    from multi_style_transfer_gan_b200 import ops, slab
    torch.manual_seed(1)
    N, dt = 2, torch.bfloat16
    x = torch.randn(N, c, H, W, device=DEV).to(dt).float()
    w = (torch.randn(3, c, 7, 7, device=DEV) * 0.02).to(dt).float()
    b = torch.randn(3, device=DEV)
    prog = slab.conv7_out_program(c)
    y = torch.empty(N, 3, H, W, device=DEV)
    bias16 = torch.zeros(16, device=DEV)
    bias16[:3] = b
    slab.conv_slab(prog, nhwc(x), slab.conv7_out_weight_slab(prog, w), bias16, act=ops.ACT_TANH, nchw_out=y)
    assert_parity(y, torch.tanh(F.conv2d(x, w, b, padding=3)), 1e-2, "7x7 output conv + tanh")


@pytest.mark.parametrize("c,H,W", [(64, 16, 128), (64, 8, 256), (16, 8, 128), (64, 20, 56)])
def test_slab_input_conv(c, H, W):
    from multi_style_transfer_gan_b200 import ops, slab
    torch.manual_seed(2)
    N, dt = 2, torch.bfloat16
    x = (torch.rand(N, 3, H, W, device=DEV) * 2 - 1).to(dt).float()
    w = (torch.randn(c, 3, 7, 7, device=DEV) * 0.1).to(dt).float()
    b = torch.randn(c, device=DEV)
    prog = slab.conv7_in_program(c)
    x8 = ops.nchw_to_nhwc(x, dt, 8)
    st = ops.new_stats(N, c, DEV)
    y = slab.conv_slab(prog, x8, slab.conv7_in_weight_slab(prog, w), b, stats=st)
    ref = F.conv2d(x, w, b, padding=3)
    assert_parity(nchw(y), ref, 1e-2, "7x7 input conv")
    assert_parity(st, torch.stack([ref.sum((2, 3)), (ref * ref).sum((2, 3))], -1), 3e-3, "stats")


# ---- taps-as-N kernel (csrc/conv_shift.cu) ----------------------------------------------------------
@pytest.mark.parametrize("H,W", [(16, 128), (8, 256), (12, 72), (6, 512), (5, 120)])
def test_shift_msb64(H, W):
    from multi_style_transfer_gan_b200 import ops, slab
    torch.manual_seed(0)
    N, C, q, dt = 2, 64, 16, torch.bfloat16
    x = torch.randn(N, C, H, W, device=DEV).to(dt).float()
    ws = [(torch.randn(q, C, k, k, device=DEV) * (1.0 / (C * k * k) ** 0.5)).to(dt).float() for k in (1, 3, 3, 3)]
    bs = [torch.randn(q, device=DEV) for _ in range(4)]
    ref = torch.cat([F.conv2d(x, w, b, padding=(w.shape[2] // 2) * d, dilation=d)
                     for w, b, d in zip(ws, bs, (1, 1, 2, 4))], 1)
    prog = slab.msb64_shift_program()
    st = ops.new_stats(N, C, DEV)
    y = slab.conv_shift(prog, nhwc(x), slab.msb64_shift_weights(ws), torch.cat(bs).contiguous(), stats=st)
    assert_parity(nchw(y), ref, 1e-2, "taps-as-N fused branches")
    assert_parity(st, torch.stack([ref.sum((2, 3)), (ref * ref).sum((2, 3))], -1), 3e-3, "stats")


@pytest.mark.parametrize("c,H,W", [(64, 16, 128), (64, 8, 384), (128, 8, 128), (64, 24, 40), (64, 4, 122)])
def test_shift_output_conv(c, H, W):
    from multi_style_transfer_gan_b200 import ops, slab
    torch.manual_seed(1)
    N, dt = 2, torch.bfloat16
    x = torch.randn(N, c, H, W, device=DEV).to(dt).float()
    w = (torch.randn(3, c, 7, 7, device=DEV) * 0.02).to(dt).float()
    b = torch.randn(3, device=DEV)
    prog = slab.conv7_out_shift_program(c)
    y = torch.empty(N, 3, H, W, device=DEV)
    slab.conv_shift(prog, nhwc(x), slab.conv7_out_shift_weights(prog, w), b, act=ops.ACT_TANH, nchw_out=y)
    assert_parity(y, torch.tanh(F.conv2d(x, w, b, padding=3)), 1e-2, "7x7 output conv + tanh (taps-as-N)")
    # one output row per tile (14 slabs per 2 rows) gives the same result as the default two-row tiles (8 slabs)
    prog1 = slab.conv7_out_shift_program(c, tile_rows=1)
    y1 = torch.empty(N, 3, H, W, device=DEV)
    slab.conv_shift(prog1, nhwc(x), slab.conv7_out_shift_weights(prog1, w), b, act=ops.ACT_TANH, nchw_out=y1)
    assert_parity(y, y1, 1e-5, "two-row vs one-row tiles")


# ---- tcgen05 wgrad (csrc/conv_wgrad_tc.cu) -----------------------------------------------------------
WGRAD_CASES = [
    ("conv", 64, 128, 4, 2, 1, 1, 64, 64),     # down1.0
    ("conv", 128, 256, 4, 2, 1, 1, 32, 32),    # down2.0
    ("conv", 128, 384, 1, 1, 0, 1, 32, 32),    # qkv
    ("conv", 64, 64, 1, 1, 0, 1, 32, 128),     # proj / fusion at C=64
    ("conv", 128, 32, 3, 1, 2, 2, 32, 32),     # branch (N small: padded to M=128 by TMA zero fill)
    ("conv", 64, 16, 3, 1, 4, 4, 16, 128),     # branch4 at C=64
    ("convT", 256, 128, 4, 2, 1, 1, 16, 16),   # up1.0
    ("convT", 128, 64, 4, 2, 1, 1, 32, 32),    # up2.0
    ("conv", 64, 8, 7, 1, 3, 1, 32, 64),       # output conv (padded filters)
    ("conv", 512, 512, 3, 1, 1, 1, 16, 16),    # D structure head
]


@pytest.mark.parametrize("case", WGRAD_CASES, ids=lambda c: "-".join(map(str, c)))
def test_wgrad_tc_vs_simt_and_torch(case):
    from multi_style_transfer_gan_b200 import ops
    kind, Cin, Cout, k, s, p, d, H, W = case
    torch.manual_seed(0)
    g = ops.ConvGeom(kind, Cin, Cout, k, s, p, d)
    N, dt = 2, torch.bfloat16
    x = torch.randn(N, Cin, H, W, device=DEV).to(dt).float()
    w = torch.zeros(*g.weight_shape(), device=DEV, requires_grad=True)
    out = F.conv_transpose2d(x, w, None, stride=2, padding=1) if kind == "convT" else F.conv2d(x, w, None, stride=s, padding=p, dilation=d)
    dy = torch.randn_like(out).to(dt).float()
    out.backward(dy)
    dw_tc = torch.zeros_like(w)
    g.wgrad(nhwc(x), nhwc(dy), dw_tc)
    dw_si = torch.zeros_like(w)
    g.wgrad(nhwc(x), nhwc(dy), dw_si, extra_flags=ops.CONV_FORCE_SIMT)
    assert_parity(dw_si, w.grad, 1e-2, "simt wgrad vs torch")
    assert_parity(dw_tc, w.grad, 1e-2, "tcgen05 wgrad vs torch")
    assert_parity(dw_tc, dw_si, 2e-3, "tcgen05 wgrad vs simt")


@pytest.mark.parametrize("Cin,Cout,H,W,N,act", [(64, 192, 64, 64, 3, "relu"), (128, 384, 32, 32, 2, "relu"),
                                                (256, 768, 16, 32, 2, "relu"), (64, 64, 16, 128, 5, "lrelu"),
                                                (128, 128, 8, 64, 2, "none")])
def test_fused_input_instnorm_1x1_is_bit_identical(Cin, Cout, H, W, N, act):
    """conv_tma's transform warps (MSG_CONV_IN_NORM on a 1x1 conv) normalise every A tile in shared memory with the
    apply kernel's arithmetic: the result must equal IN-apply followed by the plain conv BIT FOR BIT, including
    the output statistics, for tiles of several images per CTA, several channel blocks and several N tiles."""
    from multi_style_transfer_gan_b200 import ops
    torch.manual_seed(21)
    A = {"relu": ops.ACT_RELU, "lrelu": ops.ACT_LRELU, "none": ops.ACT_NONE}[act]
    x = (torch.randn(N, H, W, Cin, device=DEV) * 2 + 0.5).to(torch.bfloat16)
    w = torch.randn(Cout, Cin, 1, 1, device=DEV) * 0.1
    bias = torch.randn(Cout, device=DEV)
    g = ops.ConvGeom("conv", Cin, Cout, 1, 1, 0, 1)
    wp = g.pack_fwd(w, torch.bfloat16)
    assert g.fused_in_norm_ok(x, wp), "this geometry should take the TMA kernel with a fused input norm"
    st = ops.new_stats(N, Cin, DEV)
    ops.instnorm_stats(x, st)
    xn = ops.instnorm_apply(x, st, A)
    s_ref, s_fused = ops.new_stats(N, Cout, DEV), ops.new_stats(N, Cout, DEV)
    ref = g.forward(xn, wp, bias, stats=s_ref)
    fused = g.forward(x, wp, bias, stats=s_fused, in_stats=st, in_act=A)
    assert torch.equal(ref, fused)
    assert torch.allclose(s_ref, s_fused, rtol=1e-12, atol=1e-9)
    # and it is a real InstanceNorm: compare with torch on the same bf16 input
    t = F.instance_norm(nchw(x), eps=1e-5)
    t = {"relu": torch.relu, "lrelu": lambda v: F.leaky_relu(v, 0.2), "none": lambda v: v}[act](t)
    t = F.conv2d(t.to(torch.bfloat16).float(), w.to(torch.bfloat16).float(), bias)
    assert_parity(nchw(fused), t, 2e-2, "fused IN + 1x1 conv vs torch")


@pytest.mark.parametrize("C,H,W", [(64, 16, 128), (128, 8, 128), (256, 4, 128), (64, 12, 72)])
def test_slab_msb_dgrad(C, H, W):
    """Fused data gradient of the four MultiScaleBlock branches (one row-slab launch) against autograd of the four
    torch convs on the same bf16-rounded operands."""
    from multi_style_transfer_gan_b200 import slab
    torch.manual_seed(31)
    q, N = C // 4, 2
    ws = [torch.randn(q, C, k, k, device=DEV) * 0.05 for k in (1, 3, 3, 3)]
    db = torch.randn(N, C, H, W, device=DEV).to(torch.bfloat16)
    x = torch.zeros(N, C, H, W, device=DEV, requires_grad=True)
    outs = [F.conv2d(x, w.to(torch.bfloat16).float(), padding=(w.shape[2] // 2) * d, dilation=d)
            for w, d in zip(ws, (1, 1, 2, 4))]
    torch.cat(outs, 1).backward(db.float())
    prog = slab.msb_dgrad_program(C)
    got = slab.conv_slab(prog, nhwc(db.float()), slab.msb_dgrad_weight_slab(prog, ws), None)
    assert_parity(nchw(got), x.grad, 1e-2, "fused MSB dgrad")


@pytest.mark.parametrize("N,H,W", [(1, 16, 128), (2, 40, 64), (1, 9, 20), (3, 33, 200), (2, 64, 256), (1, 130, 136)])
def test_msb64_ring_matches_the_four_branch_convs(N, H, W):
    """csrc/msb_ring.cu (row ring of TMEM accumulators, vertical taps stacked along N) vs the four MultiScaleBlock branch
    convolutions of the reference (enhanced_generator.py:52-71) on bf16-rounded operands, with the IN statistics of its epilogue;
    planes with partial strips, segments shorter than the dilation halo and several strip segments per image."""
    from multi_style_transfer_gan_b200 import ops, slab
    torch.manual_seed(N * 1000 + H + W)
    x = torch.randn(N, 64, H, W, device=DEV).bfloat16().float()
    ws = [(torch.randn(16, 64, k, k, device=DEV) * (2.0 / (64 * k * k)) ** 0.5).bfloat16().float() for k in (1, 3, 3, 3)]
    bias = torch.randn(64, device=DEV) * 0.1
    ref = torch.cat([F.conv2d(x, w, bias[16 * i:16 * i + 16], padding=(w.shape[2] // 2) * d, dilation=d)
                     for i, (w, d) in enumerate(zip(ws, (1, 1, 2, 4)))], 1)
    st = ops.new_stats(N, 64, DEV)
    out = slab.msb64_ring(nhwc(x).bfloat16(), slab.msb64_ring_weights(ws), bias, stats=st)
    torch.cuda.synchronize()
    assert_parity(out.float().permute(0, 3, 1, 2), ref, 1e-2, f"msb64 ring {N}x{H}x{W}")
    # statistics of the fp32 accumulators (before the bf16 rounding of the store)
    assert_parity(st[..., 0].float(), ref.sum(dim=(2, 3)), 2e-3, "ring sum", floor=1e-2)
    assert_parity(st[..., 1].float(), (ref * ref).sum(dim=(2, 3)), 2e-3, "ring sum of squares")
    # a second launch gives the same bits (the ring is zeroed and drained deterministically)
    out2 = slab.msb64_ring(nhwc(x).bfloat16(), slab.msb64_ring_weights(ws), bias)
    assert torch.equal(out, out2)


@pytest.mark.parametrize("N,H,W", [(1, 16, 128), (2, 40, 64), (1, 9, 20), (3, 33, 200), (2, 64, 256)])
def test_msb128_ring_matches_the_four_branch_convs(N, H, W):
    """the C = 128 row ring (three passes: branches 1 + 2 | 3 | 4, 32-column row accumulators, two 64-channel K blocks per slab)
    vs the four branch convolutions (enhanced_generator.py:52-71) with the IN statistics of its epilogue, written as a 128-channel
    slice of a wider tensor."""
    from multi_style_transfer_gan_b200 import ops, slab
    torch.manual_seed(N * 1000 + H + W + 7)
    C, Q = 128, 32
    x = torch.randn(N, C, H, W, device=DEV).bfloat16().float()
    ws = [(torch.randn(Q, C, k, k, device=DEV) * (2.0 / (C * k * k)) ** 0.5).bfloat16().float() for k in (1, 3, 3, 3)]
    bias = torch.randn(C, device=DEV) * 0.1
    ref = torch.cat([F.conv2d(x, w, bias[Q * i:Q * i + Q], padding=(w.shape[2] // 2) * d, dilation=d)
                     for i, (w, d) in enumerate(zip(ws, (1, 1, 2, 4)))], 1)
    st = ops.new_stats(N, 2 * C, DEV)
    out = torch.zeros(N, H, W, 2 * C, device=DEV, dtype=torch.bfloat16)
    slab.msb_ring(nhwc(x).bfloat16(), slab.msb_ring_weights(ws, C), bias, C, out=out, co_off=C, stats=st)
    torch.cuda.synchronize()
    assert float(out[..., :C].abs().max()) == 0.0                          # only the slice is written
    assert_parity(out[..., C:].float().permute(0, 3, 1, 2), ref, 1e-2, f"msb128 ring {N}x{H}x{W}")
    assert_parity(st[:, C:, 0].float(), ref.sum(dim=(2, 3)), 2e-3, "ring sum", floor=1e-2)
    assert_parity(st[:, C:, 1].float(), (ref * ref).sum(dim=(2, 3)), 2e-3, "ring sum of squares")
    out2 = torch.zeros_like(out)
    slab.msb_ring(nhwc(x).bfloat16(), slab.msb_ring_weights(ws, C), bias, C, out=out2, co_off=C)
    assert torch.equal(out, out2)


@pytest.mark.parametrize("N,H,W,Cin,Cout", [(1, 16, 128, 128, 64), (2, 40, 64, 128, 64), (1, 9, 20, 64, 64), (3, 33, 200, 128, 64),
                                            (2, 64, 256, 128, 64), (1, 21, 136, 64, 128)])
def test_convt_ring_matches_conv_transpose(N, H, W, Cin, Cout):
    """csrc/convt_ring.cu (row ring of TMEM accumulators, the four vertical taps of a horizontal tap as one N = 256 MMA, one launch
    per horizontal output phase) vs ConvTranspose2d(4, 2, 1) of the reference's decoder (enhanced_generator.py:116-123) on
    bf16-rounded operands, with the IN statistics of its epilogue, written as a channel slice of a wider tensor."""
    from multi_style_transfer_gan_b200 import ops, slab
    torch.manual_seed(N * 1000 + H + W + Cin)
    x = torch.randn(N, Cin, H, W, device=DEV).bfloat16().float()
    w = (torch.randn(Cin, Cout, 4, 4, device=DEV) * (2.0 / (Cin * 4)) ** 0.5).bfloat16().float()
    bias = torch.randn(Cout, device=DEV) * 0.1
    ref = F.conv_transpose2d(x, w, bias, stride=2, padding=1)
    st = ops.new_stats(N, Cout + 64, DEV)
    out = torch.zeros(N, 2 * H, 2 * W, Cout + 64, device=DEV, dtype=torch.bfloat16)
    ws = slab.convt_ring_weights(w)
    slab.convt_ring(nhwc(x).bfloat16(), ws, bias, Cout, out=out, co_off=64, stats=st)
    torch.cuda.synchronize()
    assert float(out[..., :64].abs().max()) == 0.0                          # only the slice is written
    assert_parity(out[..., 64:].float().permute(0, 3, 1, 2), ref, 1e-2, f"convT ring {N}x{H}x{W} {Cin}->{Cout}")
    assert_parity(st[:, 64:, 0].float(), ref.sum(dim=(2, 3)), 2e-3, "ring sum", floor=1e-2)
    assert_parity(st[:, 64:, 1].float(), (ref * ref).sum(dim=(2, 3)), 2e-3, "ring sum of squares")
    out2 = torch.zeros_like(out)
    slab.convt_ring(nhwc(x).bfloat16(), ws, bias, Cout, out=out2, co_off=64)
    assert torch.equal(out, out2)


@pytest.mark.parametrize("N,H,W", [(1, 16, 128), (2, 40, 64), (1, 9, 20), (3, 33, 200), (2, 64, 256), (1, 130, 136)])
def test_out7_ring_matches_apply_plus_output_conv(N, H, W):
    """csrc/out7_ring.cu: y = tanh(conv7x7(a1 + ReLU(IN(f))) + b) in one launch (IN + ReLU + residual applied to the landed row slabs,
    7 vertical taps as one N = 112 MMA through a ring of TMEM row accumulators) vs the reference's modules
    (enhanced_generator.py:78-84, 130-133) on the same bf16-rounded a2, and vs the product's two-kernel path (apply kernel +
    conv_shift), whose a2 it must reproduce bit for bit."""
    from multi_style_transfer_gan_b200 import ops, slab
    torch.manual_seed(N * 1000 + H + W + 3)
    f = (torch.randn(N, 64, H, W, device=DEV) * 1.7 + 0.3).bfloat16()
    a1 = torch.randn(N, 64, H, W, device=DEV).bfloat16()
    w = (torch.randn(3, 64, 7, 7, device=DEV) * (1.0 / (64 * 49)) ** 0.5).bfloat16().float()
    bias = torch.randn(3, device=DEV) * 0.1
    fh, ah = nhwc(f), nhwc(a1)
    st = ops.instnorm_stats(fh)
    a2 = ops.instnorm_apply(fh, st, ops.ACT_RELU, residual=ah)              # the product's apply kernel (bf16 out)
    ref = torch.tanh(F.conv2d(a2.float().permute(0, 3, 1, 2), w, bias, padding=3))
    got = slab.out7_ring(fh, st, ah, slab.out7_ring_weights(w), bias)
    torch.cuda.synchronize()
    assert_parity(got, ref, 2e-3, f"out7 ring {N}x{H}x{W}")
    # and against fp32 modules end to end (bf16 operand rounding only)
    x32 = a1.float() + torch.relu(F.instance_norm(f.float()))
    ref32 = torch.tanh(F.conv2d(x32, w, bias, padding=3))
    assert_parity(got, ref32, 1e-2, "out7 ring vs fp32 modules")
    got2 = slab.out7_ring(fh, st, ah, slab.out7_ring_weights(w), bias)
    assert torch.equal(got, got2)


@pytest.mark.parametrize("N,H,W,Cout,fused", [(1, 16, 256, 128, True), (2, 40, 64, 128, True), (1, 10, 20, 64, True), (3, 34, 200, 128, True),
                                              (2, 64, 512, 128, True), (1, 22, 272, 64, False)])
def test_down_ring_matches_apply_plus_strided_conv(N, H, W, Cout, fused):
    """csrc/down_ring.cu: Conv2d(64, Cout, 4, stride 2, padding 1) of ReLU(IN(x)) in one launch per 64 output channels (even / odd pixel
    slabs normalised in shared memory, the two vertical taps of a horizontal tap as one N = 128 MMA through a ring of TMEM row
    accumulators) vs the reference's modules (enhanced_generator.py:93-104) on the product apply kernel's bf16 a, with the IN
    statistics of its epilogue; and without the fused norm."""
    from multi_style_transfer_gan_b200 import ops, slab
    torch.manual_seed(N * 1000 + H + W + Cout)
    x = (torch.randn(N, 64, H, W, device=DEV) * 1.3 - 0.2).bfloat16()
    w = (torch.randn(Cout, 64, 4, 4, device=DEV) * (2.0 / (64 * 16)) ** 0.5).bfloat16().float()
    bias = torch.randn(Cout, device=DEV) * 0.1
    xh = nhwc(x)
    if fused:
        sti = ops.instnorm_stats(xh)
        a = ops.instnorm_apply(xh, sti, ops.ACT_RELU)                      # the product's apply kernel (bf16 out)
    else:
        sti, a = None, xh
    ref = F.conv2d(a.float().permute(0, 3, 1, 2), w, bias, stride=2, padding=1)
    st = ops.new_stats(N, Cout + 64, DEV)
    out = torch.zeros(N, H // 2, W // 2, Cout + 64, device=DEV, dtype=torch.bfloat16)
    ws = slab.down_ring_weights(w)
    slab.down_ring(xh, sti, ws, bias, Cout, out=out, co_off=64, stats=st)
    torch.cuda.synchronize()
    assert float(out[..., :64].abs().max()) == 0.0
    assert_parity(out[..., 64:].float().permute(0, 3, 1, 2), ref, 1e-2, f"down ring {N}x{H}x{W}->{Cout}")
    assert_parity(st[:, 64:, 0].float(), ref.sum(dim=(2, 3)), 2e-3, "ring sum", floor=1e-2)
    assert_parity(st[:, 64:, 1].float(), (ref * ref).sum(dim=(2, 3)), 2e-3, "ring sum of squares")
    out2 = torch.zeros_like(out)
    slab.down_ring(xh, sti, ws, bias, Cout, out=out2, co_off=64)
    assert torch.equal(out, out2)
