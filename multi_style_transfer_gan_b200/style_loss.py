"""VGG-19 feature Gram-matrix style loss on the msg_b200 kernels (north_star addition).

The reference contains no VGG / Gram / perceptual loss, so this follows the published
formulation (Gatys et al.; Johnson et al. normalisation) restated in oracle/restate.py:
    taps = relu1_1, relu2_1, relu3_1, relu4_1, relu5_1 of torchvision's vgg19().features[:30]
    G_l  = F_l F_l^T / (C_l H_l W_l)            F_l = feat.view(B, C, HW)
    L    = sum_l mean((G_l(y) - G_l(s))^2)
The trunk runs on the same conv kernels as the generator (3x3 convs with bias + ReLU fused in the
epilogue, 2x2 max-pool kernel); the Gram matrices, the loss reduction and dL/dF come from csrc/gram.cu.
Weights: pass a torchvision vgg19 state_dict ("features.N.weight/bias"); without one the trunk is
random-init under a seed (no network access for pretrained weights), which is what the benchmark and
the parity tests use.
"""
import torch

from . import ops
from .ops import ACT_RELU, ConvGeom

VGG19_CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512]
TAP_AFTER_CONV = (0, 2, 4, 8, 12)     # conv indices (0-based) whose ReLU output is tapped


def _edge_pad(dtype):
    return 4 if dtype == torch.float32 else 8


class VGG19Features:
    """Frozen trunk up to relu5_1.  weights: list of 13 (w [Cout,Cin,3,3], b [Cout]) fp32 tensors."""

    def __init__(self, device, weights=None, seed=0):
        self.device = torch.device(device)
        if weights is None:
            g = torch.Generator().manual_seed(seed)
            weights, cin = [], 3
            for c in VGG19_CFG:
                if c == "M":
                    continue
                std = (2.0 / (cin * 9)) ** 0.5          # kaiming-normal (torchvision's vgg init is fan_out; either is fine)
                weights.append((torch.randn(c, cin, 3, 3, generator=g) * std, torch.zeros(c)))
                cin = c
        self.weights = [(w.to(self.device).float().contiguous(), b.to(self.device).float().contiguous()) for w, b in weights]
        self._packed = {}

    @staticmethod
    def from_torchvision_state_dict(sd, device):
        idx = [0, 2, 5, 7, 10, 12, 14, 16, 19, 21, 23, 25, 28]
        return VGG19Features(device, [(sd[f"features.{i}.weight"], sd[f"features.{i}.bias"]) for i in idx])

    def _geom(self, i, dtype):
        w, _ = self.weights[i]
        cin = w.shape[1] if i else _edge_pad(dtype)
        return ConvGeom("conv", cin, w.shape[0], 3, 1, 1)

    def _pack(self, i, dtype, which):
        key = (i, dtype, which)
        if key not in self._packed:
            w, _ = self.weights[i]
            if i == 0:
                w = torch.nn.functional.pad(w, [0, 0, 0, 0, 0, _edge_pad(dtype) - 3]).contiguous()
            g = self._geom(i, dtype)
            self._packed[key] = g.pack_fwd(w, dtype) if which == "fwd" else g.pack_dgrad(w, dtype)
        return self._packed[key]

    def forward(self, x, dtype, save):
        """x: fp32 NCHW [B,3,H,W] (H, W multiples of 16).  Returns (taps [5 NHWC tensors], saved)."""
        a = ops.nchw_to_nhwc(x, dtype, _edge_pad(dtype))
        taps, tape, ci = [], [], 0
        for c in VGG19_CFG:
            if c == "M":
                y = ops.maxpool_fwd(a)
                tape.append(("pool", a))
                a = y
            else:
                y = self._geom(ci, dtype).forward(a, self._pack(ci, dtype, "fwd"), self.weights[ci][1], act=ACT_RELU)
                tape.append(("conv", ci, a.shape[1:3], y))
                a = y
                if ci in TAP_AFTER_CONV:
                    taps.append(a)
                ci += 1
        return taps, (tape if save else None)

    def backward(self, tape, dtaps, dtype):
        """dtaps: gradients of the five tap feature maps (NHWC, dtype).  Returns dx fp32 NCHW."""
        d = None
        tap_convs = list(TAP_AFTER_CONV)
        for entry in reversed(tape):
            if entry[0] == "pool":
                d = ops.maxpool_bwd(entry[1], d)
                continue
            _, ci, in_hw, y = entry
            if ci in tap_convs:
                g = dtaps[tap_convs.index(ci)]
                d = g if d is None else ops.add(d, g)
            dz = ops.act_bwd(y, d, ACT_RELU)
            d = self._geom(ci, dtype).dgrad(dz, self._pack(ci, dtype, "dgrad"), in_hw)
        return ops.nhwc_to_nchw(d, 3)


class _StyleLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, vgg, targets, dtype):
        need = ctx.needs_input_grad[0]
        taps, tape = vgg.forward(x.contiguous().float(), dtype, save=need)
        loss = torch.zeros(1, device=x.device, dtype=torch.float32)
        grams = []
        for f, tgt in zip(taps, targets):
            _, g = ops.gram_loss_fwd(f, tgt, 1.0, loss)
            grams.append(g)
        if need:
            ctx.vgg, ctx.dtype, ctx.tape, ctx.taps, ctx.grams, ctx.targets = vgg, dtype, tape, taps, grams, targets
        return loss.reshape(())

    @staticmethod
    def backward(ctx, gout):
        dtaps = [ops.gram_loss_bwd(f, g, t, 1.0) for f, g, t in zip(ctx.taps, ctx.grams, ctx.targets)]
        dx = ctx.vgg.backward(ctx.tape, dtaps, ctx.dtype)
        ctx.tape = ctx.taps = None
        return dx * gout, None, None, None


class GramStyleLoss:
    """loss = sum_l mean((G_l(vgg(y)) - G_l(vgg(style)))^2); differentiable w.r.t. y."""

    def __init__(self, vgg, precision="bf16"):
        self.vgg = vgg
        self.dtype = torch.bfloat16 if precision == "bf16" else torch.float32
        self.targets = None

    @torch.no_grad()
    def set_style(self, style_images):
        """style_images: fp32 NCHW [B,3,H,W] (one style per generated image, or B=1 broadcast)."""
        taps, _ = self.vgg.forward(style_images.contiguous().float(), self.dtype, save=False)
        self.targets = [ops.gram(f) for f in taps]
        return self

    def __call__(self, y):
        if self.targets is None:
            raise RuntimeError("GramStyleLoss: call set_style() first")
        tg = self.targets
        if tg[0].shape[0] == 1 and y.shape[0] > 1:
            tg = [t.expand(y.shape[0], -1, -1).contiguous() for t in tg]
        return _StyleLossFn.apply(y, self.vgg, tg, self.dtype)
