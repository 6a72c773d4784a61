"""GPU parity of the drop-in models against the golden vectors produced from the UNMODIFIED
reference (oracle/make_golden.py) and against the CPU oracle (oracle/restate.py) on fresh inputs.
All calls go model -> ctypes -> C-ABI -> sm_100a kernels.

Tolerances (tests/util.py): forward outputs / losses <= 1e-4 (fp32 mode), image <= 2e-2 max-abs
(bf16 mode) -- BASELINE.json.  Gradients: 1e-2, because the reference's own fp32 gradients are
only reproducible to 1e-3..4e-3 against an fp64 evaluation of the same graph."""
import math

import pytest
import torch

from tests.test_oracle_golden import check_generator_grads
from tests.util import assert_parity, parity_errors

pytestmark = pytest.mark.gpu
DEV = "cuda"


def synth_images(B, H, W, seed=1234):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(B, 3, H, W, generator=g) * 2 - 1


def make_G(c, nb, sd=None, seed=None):
    from multi_style_transfer_gan_b200.enhanced_generator import EnhancedGenerator
    if seed is not None:
        torch.manual_seed(seed)
    G = EnhancedGenerator(channels=c, num_transformer_blocks=nb)
    if sd is not None:
        G.load_state_dict(sd, strict=True)
    return G.to(DEV)


def test_native_library_is_loaded():
    from multi_style_transfer_gan_b200 import _lib
    lib = _lib.load()
    assert lib.msg_check_device() == 0, lib.msg_last_error()
    assert lib.msg_sm_count() > 0
    maps = open("/proc/self/maps").read()
    assert "libmsg_b200.so" in maps


@pytest.mark.parametrize("checkpointing", [False, True])
def test_generator_golden_fp32(golden, checkpointing):
    g = golden("gen_c16_b1_64x48.pt")
    G = make_G(16, 1, g["state_dict"])
    if checkpointing:
        G.gradient_checkpointing_enable()
    x = g["x"].to(DEV).requires_grad_(True)
    y = G(x)
    assert y.shape == g["y"].shape and y.dtype == torch.float32
    assert_parity(y, g["y"], 1e-4, "y")
    r = g["r"].to(DEV)
    loss = (y * r).sum() / y.numel() + ((y - r) ** 2).mean()
    assert_parity(loss, g["loss"], 1e-4, "loss")
    loss.backward()
    grads = {k: (p.grad if p.grad is not None else None) for k, p in G.named_parameters()}
    check_generator_grads(g, x.grad, grads, rel=2e-2)


def stock_bf16_error(sd, x, ref):
    """Error of a stock-PyTorch bf16 execution (weights, input and every activation in bf16, torch
    ops) of the oracle restatement against the fp32 reference output: the comparator for our bf16
    path.  BASELINE.json's "<= 2e-2 max-abs in bf16" is NOT attainable by any bf16 execution of this
    network at random init: rounding only the INPUT IMAGE to bf16 already moves the output by 0.14,
    and the unmodified reference under torch.autocast(bf16) is off by 0.67 max-abs / 0.105 rel-L2 on
    this golden (measured; DESIGN.md "Numerics").  So the end-to-end bf16 gate is "at least as
    accurate as stock PyTorch bf16", while each kernel is held to bf16 rounding of the exact result
    in tests/test_gpu_ops.py and tests/test_gpu_conv_tc.py."""
    from oracle import restate as R
    sdb = {k: v.to(DEV).bfloat16() for k, v in sd.items()}
    with torch.no_grad():
        y = R.generator_forward(sdb, x.to(DEV).bfloat16()).float().cpu()
    return parity_errors(y, ref)


def test_generator_golden_bf16(golden):
    g = golden("gen_c16_b1_64x48.pt")
    G = make_G(16, 1, g["state_dict"]).set_precision("bf16").eval()
    with torch.no_grad():
        y = G(g["x"].to(DEV))
    mine = parity_errors(y, g["y"])
    stock = stock_bf16_error(g["state_dict"], g["x"], g["y"])
    print(f"bf16 end-to-end error vs fp32 reference: ours max {mine[0]:.3e} relL2 {mine[1]:.3e}; "
          f"stock torch bf16 max {stock[0]:.3e} relL2 {stock[1]:.3e}")
    assert mine[1] <= 1.0 * stock[1] + 1e-3, (mine, stock)      # at least as accurate as stock PyTorch bf16
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        y2 = make_G(16, 1, g["state_dict"]).eval()(g["x"].to(DEV))   # autocast selects the bf16 path
    assert torch.equal(y2, y) or parity_errors(y2, y)[1] < 1e-2


@pytest.mark.parametrize("c,nb", [(16, 1), (64, 3)])
def test_config1_two_style_blend(golden, c, nb):
    """BASELINE.json configs[0]: 1x3x256x256 fp32, two state_dicts (seeds 0, 1), 0.7/0.3 blend."""
    from multi_style_transfer_gan_b200 import ops
    g = golden(f"config1_c{c}_256.pt")
    x = synth_images(1, 256, 256).to(DEV)
    ys = []
    with torch.no_grad():
        for seed in g["seeds"]:
            ys.append(make_G(c, nb, seed=seed).eval()(x))
    # rel-L2 <= 1e-4.  Normalised max error: the reference's own fp32 output is only reproducible to
    # 1.6e-4 against an fp64 evaluation for the c=64 case (4.7e-5 for c=16), measured -- so the max
    # bound is 1e-4 for c=16 and 5e-4 (3x the reference's own noise floor) for c=64.
    mx = 1e-4 if c == 16 else 5e-4
    assert_parity(ys[0], g["y0"], 1e-4, "y0", max_rel=mx)
    assert_parity(ys[1], g["y1"], 1e-4, "y1", max_rel=mx)
    out = ops.blend_outputs(ys, g["w"])
    assert_parity(out, 0.7 * g["y0"] + 0.3 * g["y1"], 1e-4, "blend", max_rel=mx)
    # bf16 arm of the same config, gated against stock PyTorch bf16 (see stock_bf16_error)
    Gb = make_G(c, nb, seed=g["seeds"][0]).set_precision("bf16").eval()
    with torch.no_grad():
        yb = Gb(x)
    mine = parity_errors(yb, g["y0"])
    stock = stock_bf16_error({k: v.detach().cpu() for k, v in Gb.state_dict().items()}, x.cpu(), g["y0"])
    print(f"config1 c={c} bf16: ours max {mine[0]:.3e} relL2 {mine[1]:.3e}; stock torch bf16 max {stock[0]:.3e} relL2 {stock[1]:.3e}")
    assert mine[1] <= 1.0 * stock[1] + 1e-3, (mine, stock)      # at least as accurate as stock PyTorch bf16


def test_bad_sizes_raise():
    G = make_G(8, 1, seed=0)
    for hw in ((130, 130), (250, 256), (64, 40)):
        with pytest.raises(RuntimeError):
            G(torch.zeros(1, 3, *hw, device=DEV))
    with pytest.raises(RuntimeError):
        G(torch.zeros(1, 4, 64, 64, device=DEV))


def test_generator_sizes_and_batch_independence():
    """direct_transform.py:86 sizes; images are independent (IN per-sample, attention per-window):
    a batch equals the per-image results -> image sharding needs no communication."""
    G = make_G(8, 1, seed=3).eval()
    with torch.no_grad():
        for hw in ((128, 128), (64, 96)):
            x = synth_images(3, *hw, seed=5).to(DEV)
            yb = G(x)
            for i in range(3):
                yi = G(x[i:i + 1])
                assert_parity(yb[i:i + 1], yi, 2e-6, f"batch independence {hw} #{i}")


def test_generator_vs_oracle_fresh_weights():
    from oracle import restate as R
    G = make_G(8, 2, seed=11).eval()
    x = synth_images(2, 32, 64, seed=7)
    with torch.no_grad():
        y = G(x.to(DEV))
        ref = R.generator_forward({k: v.cpu() for k, v in G.state_dict().items()}, x)
    assert_parity(y, ref, 1e-4, "G vs oracle")


def test_user_supplied_transformer_block():
    """pluggable StructuralTransformerBlock slot: the style encoder becomes live
    (enhanced_generator.py:216-225)."""
    from oracle import restate as R
    import torch.nn as nn

    class Blk(nn.Module):
        def __init__(self, dim):
            super().__init__()
            self.lin = nn.Linear(dim, dim)

        def forward(self, x, style, orig):
            return x + 0.1 * torch.tanh(self.lin(x)) * style[:, None, :] + orig.mean()

    G = make_G(8, 1, seed=2)
    torch.manual_seed(0)
    G.transformer_blocks[0] = Blk(32).to(DEV)
    x = synth_images(2, 32, 32, seed=9)
    y = G(x.to(DEV))
    sd = {k: v.detach().cpu() for k, v in G.state_dict().items()}
    blk = Blk(32)
    blk.load_state_dict({k.split("transformer_blocks.0.")[1]: v for k, v in sd.items() if k.startswith("transformer_blocks.0.")})
    ref = R.generator_forward(sd, x, blocks=[blk])
    assert_parity(y, ref, 1e-4, "with block")
    y.square().mean().backward()
    assert G.style_encoder[2].weight.grad is not None and float(G.style_encoder[2].weight.grad.abs().max()) > 0


def test_discriminator_golden(golden):
    from multi_style_transfer_gan_b200.enhanced_generator import EnhancedDiscriminator
    g = golden("disc_c8_64.pt")
    D = EnhancedDiscriminator(8)
    D.load_state_dict(g["state_dict_before"], strict=True)
    D = D.to(DEV).train()
    x = g["x"].to(DEV).requires_grad_(True)
    score, struct = D(x)
    assert_parity(score, g["score"], 1e-4, "score")
    assert_parity(struct, g["struct"], 1e-4, "struct")
    after = D.state_dict()
    for k, v in g["state_dict_after"].items():
        if k.endswith("weight_u") or k.endswith("weight_v"):
            assert_parity(after[k], v, 1e-4, k)     # power iteration advanced in place
    loss = ((score - 1) ** 2).mean() + struct.abs().mean()
    assert_parity(loss, g["loss"], 1e-4, "loss")
    loss.backward()
    check_generator_grads(g, x.grad, {k: p.grad for k, p in D.named_parameters()}, rel=2e-2)
    # eval mode, B=1: no power iteration, 0-dim score (.squeeze() quirk, :275)
    e = golden("disc_c8_64_eval_b1.pt")
    D.eval()
    before = {k: v.clone() for k, v in D.state_dict().items()}
    with torch.no_grad():
        s1, st1 = D(synth_images(1, 64, 64).to(DEV))
    assert list(s1.shape) == [] and st1.shape == (1, 1, 3, 3)
    assert_parity(s1, e["score"], 1e-4, "eval score")
    assert_parity(st1, e["struct"], 1e-4, "eval struct")
    for k, v in D.state_dict().items():
        assert torch.equal(v, before[k]), k


def test_train_step_golden(golden):
    """Two consecutive EnhancedCycleGAN.train_step calls vs the reference's (c=8 members)."""
    from multi_style_transfer_gan_b200.enhanced_train import EnhancedCycleGAN
    g = golden("train_step_c8_64.pt")
    m = EnhancedCycleGAN(channels=8, num_transformer_blocks=1, precision="fp32")
    m.load_state_dicts(**g["init"])
    assert {"G_AB", "G_BA", "D_A", "D_B", "g_optimizer", "d_optimizer", "device"} <= set(vars(m))
    for step, ref in enumerate(g["losses"]):
        got = m.train_step(g["real_A"], g["real_B"])
        assert list(got.keys()) == ["d_loss", "g_loss", "cycle_loss", "identity_loss", "structure_loss"]
        tol = 1e-4 if step == 0 else 2e-3     # see tests/test_oracle_golden.py::test_train_step
        for k in ref:
            assert abs(got[k] - ref[k]) <= tol * abs(ref[k]) + 1e-6, (step, k, got[k], ref[k])
    for n in ("D_A", "D_B"):
        sd = getattr(m, n).state_dict()
        for k, v in g["final"][n].items():
            if k.endswith("weight_u") or k.endswith("weight_v"):
                assert_parity(sd[k], v, 2e-3, f"{n}.{k}")
    # weights moved, by about lr per element (Adam's first steps)
    sd = m.G_AB.state_dict()
    moved = (sd["output.0.weight"].cpu() - g["init"]["G_AB"]["output.0.weight"]).abs().max()
    assert 1e-5 < float(moved) < 2e-4
    assert_parity(sd["output.0.weight"], g["final"]["G_AB"]["output.0.weight"], 1e-3, "updated output.0.weight")


def test_train_step_bf16_runs_and_tracks_fp32(golden):
    from multi_style_transfer_gan_b200.enhanced_train import EnhancedCycleGAN
    g = golden("train_step_c8_64.pt")
    m = EnhancedCycleGAN(channels=8, num_transformer_blocks=1, precision="bf16")
    m.load_state_dicts(**g["init"])
    got = m.train_step(g["real_A"], g["real_B"])
    ref = g["losses"][0]
    # Single step only (bf16 vs fp32 diverge chaotically afterwards).  This is a sanity bound, not a parity
    # gate (those are the fp32 golden tests): measured bf16-vs-fp32 deviations of the step-1 losses are 0.01 % (identity),
    # 0.1 % (cycle), 1 % (d), 2-3 % (g) and 3-4 % (structure: an L1 between two nearly equal discriminator maps), and they
    # move by ~1 % between equally accurate kernel variants (e.g. bf16 vs fp16 softmax weights in LocalAttention).
    for k in ref:
        assert abs(got[k] - ref[k]) <= 6e-2 * abs(ref[k]) + 1e-3, (k, got[k], ref[k])


def test_train_step_cuda_graph_matches_eager(golden):
    """use_graph=True: after the eager warm-up steps the whole step is captured once and replayed.  Same kernels in the same
    order, so the losses of every step -- and the weights after four steps -- agree with the eager run to fp32 summation-order
    noise (the weight-gradient reductions are order-free adds), including across the capture boundary (Adam's step count lives
    on the device)."""
    from multi_style_transfer_gan_b200.enhanced_train import EnhancedCycleGAN
    g = golden("train_step_c8_64.pt")
    runs = []
    for use_graph in (False, True, "segmented"):
        m = EnhancedCycleGAN(channels=8, num_transformer_blocks=1, precision="fp32", use_graph=bool(use_graph), graph_warmup=1)
        m.segment_at_syncs = use_graph == "segmented"      # the data-parallel form: four graph segments, the all-reduces eager between them
        m.load_state_dicts(**g["init"])
        losses = [m.train_step(g["real_A"], g["real_B"]) for _ in range(4)]
        if use_graph:
            assert m._graph is not None and m.graph_error is None
            assert len(m._graph["graphs"]) == (4 if use_graph == "segmented" else 1)
            assert m.g_optimizer.step_count == 4 and int(m.g_optimizer.step_dev.item()) == 4
        runs.append((losses, m.g_optimizer.flat.clone(), m.d_optimizer.flat.clone()))
    # Step 1 is eager in both runs and step 2 is the first replay: both agree to the noise of the order-free reductions.  From
    # step 3 on two EAGER runs of this c=8 network already differ by 5e-3 .. 5e-2 (measured, tools/graph_vs_eager.py: Adam's
    # first updates are ~lr * sign(g), so gradient noise on near-zero gradients flips whole updates) -- the graph run sits in
    # the same band, which is all that can be asserted there.
    for other in (1, 2):
        for step, (a, b) in enumerate(zip(runs[0][0], runs[other][0])):
            tol = 1e-5 if step == 0 else 1e-4 if step == 1 else 0.15
            for k in a:
                assert abs(a[k] - b[k]) <= tol * abs(a[k]) + 1e-6, (other, step, k, a[k], b[k])
        moved = float((runs[0][1] - runs[other][1]).abs().mean()), float((runs[0][2] - runs[other][2]).abs().mean())
        assert moved[0] <= 2e-4 and moved[1] <= 8e-4, moved          # (4 steps x lr: 2e-4 / 8e-4 is "every update flipped")


def test_save_models_layout(tmp_path):
    from multi_style_transfer_gan_b200.enhanced_train import EnhancedCycleGAN
    m = EnhancedCycleGAN(channels=8, num_transformer_blocks=1, precision="fp32")
    m.save_models(tmp_path, 3)
    a = torch.load(tmp_path / "G_AB_epoch_3.pth", weights_only=False)
    assert set(a) == {"epoch", "G_AB_state_dict"} and len(a["G_AB_state_dict"]) == 70
    d = torch.load(tmp_path / "discriminators_epoch_3.pth", weights_only=False)
    assert set(d) == {"epoch", "D_A_state_dict", "D_B_state_dict"} and len(d["D_A_state_dict"]) == 28


def test_train_step_with_gram_style_term():
    """north_star / BASELINE config 4 extension: lambda_style > 0 adds the VGG Gram term to the generator objective
    (sixth dict key) and changes the generator update; the discriminator phase, which precedes it, is unaffected."""
    from multi_style_transfer_gan_b200.enhanced_train import EnhancedCycleGAN
    from multi_style_transfer_gan_b200.style_loss import GramStyleLoss, VGG19Features
    torch.manual_seed(3)
    A = torch.rand(2, 3, 64, 64) * 2 - 1
    B = torch.rand(2, 3, 64, 64) * 2 - 1
    outs, params = [], []
    for lam in (0.0, 10.0):
        torch.manual_seed(0)
        style = GramStyleLoss(VGG19Features(DEV, seed=0), precision="fp32") if lam else None
        m = EnhancedCycleGAN(channels=8, precision="fp32", device=DEV, style_loss=style, lambda_style=lam)
        outs.append(m.train_step(A, B))
        params.append(m.g_optimizer.flat.clone())
    assert set(outs[0]) == {"d_loss", "g_loss", "cycle_loss", "identity_loss", "structure_loss"}
    assert set(outs[1]) == set(outs[0]) | {"style_loss"}
    assert outs[1]["style_loss"] > 0 and math.isfinite(outs[1]["style_loss"])
    for k in outs[0]:
        assert abs(outs[0][k] - outs[1][k]) <= 1e-5 * max(1.0, abs(outs[0][k])), k
    assert float((params[0] - params[1]).abs().max()) > 0
    with pytest.raises(ValueError):
        EnhancedCycleGAN(channels=8, device=DEV, lambda_style=1.0)


def test_config5_highres_four_style_blend_properties():
    """BASELINE config 5 (1024x1024, 4 blended styles, bf16) at B=2 through size-independent properties: the
    stylizer's output is the linear blend of the four single-style outputs, it does not depend on how the batch is
    cut into micro-batches or on the other images in the batch, host-in/uint8-out equals device-in + manual
    conversion, and the peak activation footprint stays far below one GPU's HBM."""
    from multi_style_transfer_gan_b200 import ops
    from multi_style_transfer_gan_b200.enhanced_generator import EnhancedGenerator
    from multi_style_transfer_gan_b200.stylize import MultiStyleStylizer
    from oracle import restate as R
    gens = []
    for s in range(4):
        torch.manual_seed(s)
        gens.append(EnhancedGenerator(channels=16, num_transformer_blocks=1).to(DEV))
    w = [0.4, 0.3, 0.2, 0.1]
    torch.manual_seed(1234)
    x = (torch.rand(2, 3, 1024, 1024) * 2 - 1).to(DEV)
    torch.cuda.reset_peak_memory_stats()
    st = MultiStyleStylizer(gens, precision="bf16", micro_batch=2)
    y = st(x, w)
    peak = torch.cuda.max_memory_allocated()
    assert y.shape == x.shape and torch.isfinite(y).all() and float(y.abs().max()) <= 1.0 + 1e-6
    assert peak < 40 * 2 ** 30, peak
    with torch.no_grad():          # the stylizer's (inference) schedule; the training forward keeps qkv / attention for backward and
        singles = [st.generators[s](x) for s in range(4)]      # runs the unfused kernels, equal only to bf16 rounding
    ref = R.blend_outputs([t.float().cpu() for t in singles], w)
    assert_parity(y, ref, 1e-6, "blend of singles")
    y1 = MultiStyleStylizer(gens, precision="bf16", micro_batch=1)(x, w)
    assert torch.equal(y1, y), "micro-batching changed the result"
    y0 = MultiStyleStylizer(gens, precision="bf16", micro_batch=2)(x[:1].contiguous(), w)
    assert torch.equal(y0, y[:1]), "an image's result depends on its batch neighbours"
    u8 = torch.empty(2, 3, 1024, 1024, dtype=torch.uint8).pin_memory()
    st(x.cpu().pin_memory(), w, out_uint8=True, out=u8)
    torch.cuda.synchronize()
    assert (u8.int() - R.to_uint8_image(y.cpu()).int()).abs().max() <= 1


def test_stylizer_schedules_agree():
    """Style streams and CUDA-graph replay only change WHEN kernels run: results are bit-identical to the serial
    schedule, also on a second call (replay) with different input and after a weight update (graph re-capture)."""
    from multi_style_transfer_gan_b200.enhanced_generator import EnhancedGenerator
    from multi_style_transfer_gan_b200.stylize import MultiStyleStylizer
    gens = []
    for s in range(3):
        torch.manual_seed(s)
        gens.append(EnhancedGenerator(channels=64, num_transformer_blocks=3).to(DEV))
    w = [0.2, 0.3, 0.5]
    torch.manual_seed(7)
    xs = [(torch.rand(3, 3, 128, 128) * 2 - 1).to(DEV) for _ in range(2)]
    serial = MultiStyleStylizer(gens, micro_batch=2, style_streams=False, use_graph=False)
    streams = MultiStyleStylizer(gens, micro_batch=2, style_streams=True, use_graph=False)
    graph = MultiStyleStylizer(gens, micro_batch=2, style_streams=True, use_graph=True)
    for x in xs:
        ref = serial(x, w)
        assert torch.equal(streams(x, w), ref), "style streams changed the result"
        assert torch.equal(graph(x, w), ref), "graph replay changed the result"
        u8 = graph(x, w, out_uint8=True)
        assert torch.equal(u8, serial(x, w, out_uint8=True))
    with torch.no_grad():
        gens[1].output[0].weight.mul_(0.5)
    ref = serial(xs[0], w)
    assert torch.equal(graph(xs[0], w), ref), "graph replay used stale weights"


def test_fused_input_norm_does_not_change_the_generator():
    """Inference fuses ReLU(IN(.)) into its only consumer (the 1x1 qkv / fusion convs) where the TMA kernel supports
    it; the output must be identical to the unfused schedule."""
    from multi_style_transfer_gan_b200.enhanced_generator import EnhancedGenerator
    torch.manual_seed(0)
    g = EnhancedGenerator(channels=64, num_transformer_blocks=3).to(DEV).eval()
    g.set_precision("bf16")
    x = torch.rand(2, 3, 128, 128, device=DEV) * 2 - 1
    eng = g._engine if hasattr(g, "_engine") else g.engine
    with torch.no_grad():
        assert eng.fuse_in_norm
        y1 = g(x)
        eng.fuse_in_norm = False
        y0 = g(x)
        eng.fuse_in_norm = True
    assert torch.equal(y0, y1)


@pytest.mark.parametrize("hw", [(192, 320), (128, 384), (80, 48)])
def test_c64_bf16_odd_plane_geometries_track_fp32(hw):
    """c=64 generator on planes the TMA kernels cannot box at every level (widths 320 / 80 / 48: not a power of two
    and not a multiple of 128 after down-sampling), so the gather (cp.async + tcgen05) and SIMT engines and the
    partial-tile paths of the slab kernels are mixed with the TMA ones.  bf16 must track the (oracle-checked) fp32
    engine within the bf16 noise measured for this network (rel-L2 0.12 at c=64), and stay batch-independent."""
    G = make_G(64, 3, seed=5).eval()
    x = synth_images(2, *hw, seed=9).to(DEV)
    with torch.no_grad():
        G.set_precision("fp32")
        y32 = G(x)
        G.set_precision("bf16")
        y16 = G(x)
        y16_0 = G(x[:1].contiguous())
    _, l2 = parity_errors(y16, y32)
    assert l2 <= 0.2, l2
    assert torch.isfinite(y16).all()
    _, l2b = parity_errors(y16_0, y16[:1])
    assert l2b <= 1e-2, l2b


def _pretrain_model(g, precision):
    from multi_style_transfer_gan_b200.pretrain import Generator
    torch.manual_seed(0)
    G = Generator(channels=8).to(DEV)
    G.load_state_dict(g["init"], strict=True)
    return G.set_precision(precision)


def test_pretrain_generator_golden_fp32(golden):
    """pretrain.Generator (pretrain.py:60-97, BatchNorm auto-encoder) on the msg_b200 kernels vs the golden from the
    unmodified reference: train-mode forward (batch statistics), masked-L1 loss (:160), every gradient, the running
    statistics update, and the eval-mode forward (running statistics)."""
    from multi_style_transfer_gan_b200.losses import l1
    g = golden("pretrain_c8_64.pt")
    G = _pretrain_model(g, "fp32").train()
    x = g["masked"].to(DEV).requires_grad_(True)
    real, mask = g["real"].to(DEV), g["mask"].to(DEV)
    y = G(x)
    assert_parity(y, g["y_train"], 1e-4, "pretrain y (train)")
    loss = l1(y * (1 - mask), real * (1 - mask))
    assert abs(float(loss) - float(g["loss"])) <= 1e-5
    loss.backward()
    assert_parity(x.grad, g["dx"], 2e-2, "pretrain dx")
    for k, p in G.named_parameters():
        ref = g["grads"][k]
        if float(ref.abs().max()) < 1e-7:      # conv biases in front of a BatchNorm: analytically zero, pure noise
            assert float(p.grad.abs().max()) < 1e-5, k
            continue
        assert_parity(p.grad, ref, 2e-2, f"pretrain grad {k}")
    sd = G.state_dict()
    for k, ref in g["running_after"].items():
        assert torch.allclose(sd[k].double().cpu(), ref.double(), rtol=1e-4, atol=1e-6), k
    G.eval()
    with torch.no_grad():
        y_eval = G(g["masked"].to(DEV))
    assert_parity(y_eval, g["y_eval"], 1e-4, "pretrain y (eval)")
    with pytest.raises(RuntimeError):
        G(torch.zeros(1, 3, 40, 64, device=DEV))


def test_pretrain_generator_bf16_and_step(golden):
    """bf16 (the reference trains under autocast, pretrain.py:158) tracks fp32, and the masked-L1 step with gradient
    clipping (:159-166) reduces the loss on a fixed batch."""
    from multi_style_transfer_gan_b200.enhanced_train import FusedAdam
    from multi_style_transfer_gan_b200.pretrain import Generator, pretrain_step
    g = golden("pretrain_c8_64.pt")
    G = _pretrain_model(g, "bf16").train()
    with torch.no_grad():
        y = G(g["masked"].to(DEV))
    _, l2 = parity_errors(y, g["y_train"])
    assert l2 <= 5e-2, l2
    torch.manual_seed(1)
    G = Generator(channels=64).to(DEV).set_precision("bf16").train()
    opt = FusedAdam(G.parameters(), lr=2e-4, betas=(0.5, 0.999))
    real = synth_images(4, 128, 128, seed=3).to(DEV)
    mask = (torch.rand(4, 1, 128, 128, device=DEV) > 0.3).float()
    losses = [pretrain_step(G, opt, real * mask, real, mask) for _ in range(12)]
    assert all(math.isfinite(v) for v in losses)
    assert losses[-1] < losses[0], losses


@pytest.mark.parametrize("c", [64, 32, 128])
def test_ring_inference_matches_the_unfused_engine_paths(c):
    """Inference at c = 64 (bf16) with the row-ring kernels and their fused InstanceNorm applies (csrc/msb_ring.cu, convt_ring.cu,
    down_ring.cu, out7_ring.cu) vs the same generator on the per-tap kernels + stand-alone apply launches, both measured against the
    fp32 engine (enhanced_generator.py:86-147): two bf16 evaluations of a 60-layer network differ from each other by as much as each
    differs from fp32, so the gate is that the ring path is no further from fp32 than the per-tap path; on a plane narrower than a
    strip and on one with several strips.  c = 32 / 128 put the MultiScaleBlock and transposed-conv rings on other stages (the fused
    down / output kernels are c = 64 only)."""
    from multi_style_transfer_gan_b200.enhanced_generator import EnhancedGenerator
    torch.manual_seed(11)
    G = EnhancedGenerator(channels=c, num_transformer_blocks=1).to(DEV).eval()
    eng = G._engine
    for H, W in ((64, 64), (32, 1040)) if c == 64 else ((48, 272),):
        x = torch.rand(2, 3, H, W, device=DEV) * 2 - 1
        with torch.no_grad():
            y32 = G.set_precision("fp32")(x).float().clone()
            G.set_precision("bf16")
            y_ring = G(x).float().clone()
            flags = {k: getattr(eng, k) for k in ("msb64_ring", "msb128_ring", "convT_ring", "out7_ring", "down_ring")}
            try:
                for k in flags:
                    setattr(eng, k, False)
                y_tap = G(x).float().clone()
            finally:
                for k, v in flags.items():
                    setattr(eng, k, v)
        assert torch.isfinite(y_ring).all()
        e_ring = float((y_ring - y32).norm() / y32.norm())
        e_tap = float((y_tap - y32).norm() / y32.norm())
        assert e_ring <= 1.25 * e_tap + 1e-3, (c, H, W, e_ring, e_tap)
