"""TEST INFRASTRUCTURE ONLY -- CPU restatement (plain PyTorch fp32) of the reference hot path.

This is the oracle the CUDA path is checked against.  It is a *functional* restatement that
works directly on ``state_dict`` tensors (no nn.Module tree), so it travels to the GPU box where
/root/reference does not exist.  It is pinned against the real reference by
``tests/test_init_matches_reference.py`` (container only) and against the committed golden vectors
in ``tests/golden/`` (everywhere); the goldens were produced by ``oracle/make_golden.py`` from
the unmodified reference.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this module.  The product package never does.

Parity status
  * generator / discriminator / train_step / blends: PINNED (reference imported, goldens).
  * StructuralTransformerBlock: source missing from the reference -> identity, parity UNPINNED.
  * blended-affine InstanceNorm: not in the reference -> parity UNPINNED (with gamma = 1, beta = 0 it
    reduces to the pinned InstanceNorm).
  * Gram/VGG style loss: not in the reference.  The VGG-19 trunk restatement is PINNED against the named
    dependency -- torchvision 0.26 ``vgg19(weights=None).features[:30]`` taps on shared seeded weights
    (tests/test_oracle_vgg_torchvision.py); the Gram / loss normalisation follows the published formulation
    (Gatys et al. / Johnson et al.: G = F F^T / (C H W), mean squared difference per layer) and has no reference to pin.

Every function cites the reference lines it follows (relative to /root/reference).
"""
import math

import torch
import torch.nn.functional as F

IN_EPS = 1e-5  # nn.InstanceNorm2d default, enhanced_generator.py:54,59,...,129
NORMALIZE_EPS = 1e-12  # F.normalize default, enhanced_generator.py:31-32
SN_EPS = 1e-12  # torch.nn.utils.spectral_norm default, enhanced_generator.py:269-271


# --------------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------------
def instance_norm(x, eps=IN_EPS):
    """nn.InstanceNorm2d(affine=False, track_running_stats=False): biased variance, eps in sqrt."""
    mu = x.mean(dim=(2, 3), keepdim=True)
    var = ((x - mu) ** 2).mean(dim=(2, 3), keepdim=True)
    return (x - mu) / torch.sqrt(var + eps)


def blended_affine_instance_norm(x, gammas, betas, w, eps=IN_EPS):
    """north_star extension (NOT in reference, parity unpinned): IN followed by the style blend
    gamma = sum_s w_s gamma_s, beta = sum_s w_s beta_s.  gammas/betas: [S, C], w: [S]."""
    g = (w[:, None] * gammas).sum(0)
    b = (w[:, None] * betas).sum(0)
    return instance_norm(x, eps) * g[None, :, None, None] + b[None, :, None, None]


def local_attention(x, qkv_w, qkv_b, proj_w, proj_b, ws=4):
    """enhanced_generator.py:13-47, restated as: 1x1 qkv on the full map -> per-window *channel*
    attention (C x C logits over the ws*ws pixels of one window, q and k L2-normalised over C per
    pixel, no scale, softmax over the last dim) -> 1x1 proj.  The reference's pad branch is
    broken for H % ws != 0 (:15-23, views with the un-padded H // ws) so such sizes raise."""
    B, C, H, W = x.shape
    if H % ws or W % ws:
        raise RuntimeError(f"LocalAttention: H, W must be multiples of {ws}, got {H}x{W}")
    qkv = F.conv2d(x, qkv_w, qkv_b)
    q, k, v = qkv.chunk(3, dim=1)

    def windows(t):  # [B,C,H,W] -> [B*nWh*nWw, C, ws*ws]
        t = t.reshape(B, C, H // ws, ws, W // ws, ws).permute(0, 2, 4, 1, 3, 5)
        return t.reshape(-1, C, ws * ws)

    q, k, v = windows(q), windows(k), windows(v)
    qn = q / q.norm(dim=1, keepdim=True).clamp_min(NORMALIZE_EPS)
    kn = k / k.norm(dim=1, keepdim=True).clamp_min(NORMALIZE_EPS)
    attn = torch.softmax(qn @ kn.transpose(1, 2), dim=-1)  # [nW, C, C]
    o = attn @ v  # [nW, C, ws*ws]
    o = o.reshape(B, H // ws, W // ws, C, ws, ws).permute(0, 3, 1, 4, 2, 5).reshape(B, C, H, W)
    return F.conv2d(o, proj_w, proj_b)


def multi_scale_block(x, sd, prefix):
    """enhanced_generator.py:49-84.  Branch INs + cat == one IN over the concatenated tensor."""
    outs = []
    for i, (pad, dil) in enumerate([(0, 1), (1, 1), (2, 2), (4, 4)], start=1):
        w = sd[f"{prefix}.branch{i}.0.weight"]
        b = sd[f"{prefix}.branch{i}.0.bias"]
        outs.append(F.conv2d(x, w, b, padding=pad, dilation=dil))
    cat = torch.relu(instance_norm(torch.cat(outs, dim=1)))
    f = F.conv2d(cat, sd[f"{prefix}.fusion.0.weight"], sd[f"{prefix}.fusion.0.bias"])
    return torch.relu(instance_norm(f)) + x  # ReLU before the residual add (:84)


def _stage(x, sd, name, transposed):
    """down1/down2 (:98-111) and up1/up2 (:120-133): (conv|convT)4x4 s2 p1 -> IN -> ReLU ->
    LocalAttention(ws=4) -> MultiScaleBlock."""
    w, b = sd[f"{name}.0.weight"], sd[f"{name}.0.bias"]
    if transposed:
        x = F.conv_transpose2d(x, w, b, stride=2, padding=1)
    else:
        x = F.conv2d(x, w, b, stride=2, padding=1)
    x = torch.relu(instance_norm(x))
    x = local_attention(x, sd[f"{name}.3.qkv.weight"], sd[f"{name}.3.qkv.bias"],
                        sd[f"{name}.3.proj.weight"], sd[f"{name}.3.proj.bias"], ws=4)
    return multi_scale_block(x, sd, f"{name}.4")


def generator_forward(sd, x, blocks=None):
    """EnhancedGenerator.forward, enhanced_generator.py:211-228 (the checkpoint branch :183-209
    is bit-identical).  ``blocks``: optional callables block(tokens, style, orig) standing in for
    the missing StructuralTransformerBlock; default = identity (style_encoder then is dead)."""
    if x.shape[2] % 16 or x.shape[3] % 16:
        raise RuntimeError("EnhancedGenerator: H and W must be multiples of 16")
    orig = x
    x = F.conv2d(x, sd["initial.0.weight"], sd["initial.0.bias"], padding=3)
    x = torch.relu(instance_norm(x))
    x = _stage(x, sd, "down1", False)
    x = _stage(x, sd, "down2", False)
    if blocks:
        style = torch.relu(F.linear(x.mean(dim=(2, 3)), sd["style_encoder.2.weight"],
                                    sd["style_encoder.2.bias"]))
        B, C, H, W = x.shape
        t = x.flatten(2).transpose(1, 2)
        for blk in blocks:
            t = blk(t, style, orig)
        x = t.transpose(1, 2).reshape(B, C, H, W)
    x = _stage(x, sd, "up1", True)
    x = _stage(x, sd, "up2", True)
    x = F.conv2d(x, sd["output.0.weight"], sd["output.0.bias"], padding=3)
    return torch.tanh(x)


def generator_state_dict_keys(num_transformer_blocks=0):
    """The 70 keys, in the reference's order (dumped from the reference class)."""
    keys = ["initial.0.weight", "initial.0.bias"]

    def stage(n):
        k = [f"{n}.0.weight", f"{n}.0.bias", f"{n}.3.qkv.weight", f"{n}.3.qkv.bias",
             f"{n}.3.proj.weight", f"{n}.3.proj.bias"]
        for i in range(1, 5):
            k += [f"{n}.4.branch{i}.0.weight", f"{n}.4.branch{i}.0.bias"]
        k += [f"{n}.4.fusion.0.weight", f"{n}.4.fusion.0.bias"]
        return k

    keys += stage("down1") + stage("down2") + stage("up1") + stage("up2")
    keys += ["output.0.weight", "output.0.bias", "style_encoder.2.weight", "style_encoder.2.bias"]
    return keys


# --------------------------------------------------------------------------------------------
# discriminator with old-style spectral norm
# --------------------------------------------------------------------------------------------
D_CONVS = ["main.0", "main.2", "main.5", "main.8", "batch_head.0", "structure_head.0",
           "structure_head.3"]


def spectral_norm_weight(w_orig, u, v, training):
    """torch.nn.utils.spectral_norm (old style) as applied at enhanced_generator.py:269-271:
    one power iteration per training-mode forward (u, v updated in place, no grad), then
    sigma = u^T W v and weight = weight_orig / sigma.  Returns (weight, u_new, v_new)."""
    wm = w_orig.reshape(w_orig.shape[0], -1)
    if training:
        with torch.no_grad():
            v = F.normalize(torch.mv(wm.t(), u), dim=0, eps=SN_EPS)
            u = F.normalize(torch.mv(wm, v), dim=0, eps=SN_EPS)
    sigma = torch.dot(u, torch.mv(wm, v))
    return w_orig / sigma, u, v


def discriminator_forward(sd, x, training=True):
    """EnhancedDiscriminator.forward, enhanced_generator.py:230-275.  Returns
    (score.squeeze(), struct_map, new_uv) with new_uv = {name: (u, v)} after this forward's power
    iteration (the reference mutates the buffers in place)."""
    new_uv = {}

    def conv(name, t, stride, pad):
        w, u, v = spectral_norm_weight(sd[f"{name}.weight_orig"], sd[f"{name}.weight_u"],
                                       sd[f"{name}.weight_v"], training)
        new_uv[name] = (u, v)
        return F.conv2d(t, w, sd[f"{name}.bias"], stride=stride, padding=pad)

    h = F.leaky_relu(conv("main.0", x, 2, 1), 0.2)
    h = F.leaky_relu(instance_norm(conv("main.2", h, 2, 1)), 0.2)
    h = F.leaky_relu(instance_norm(conv("main.5", h, 2, 1)), 0.2)
    h = F.leaky_relu(instance_norm(conv("main.8", h, 2, 1)), 0.2)
    score = conv("batch_head.0", h, 1, 1).mean(dim=(2, 3), keepdim=True).squeeze()
    s = F.leaky_relu(instance_norm(conv("structure_head.0", h, 1, 1)), 0.2)
    struct = conv("structure_head.3", s, 1, 1)
    return score, struct, new_uv


# --------------------------------------------------------------------------------------------
# training step
# --------------------------------------------------------------------------------------------
class AdamState:
    """torch.optim.Adam restated (enhanced_train.py:36-43): beta=(0.5,0.999), eps 1e-8, no decay."""

    def __init__(self, params, lr, betas=(0.5, 0.999), eps=1e-8):
        self.params, self.lr, self.betas, self.eps = params, lr, betas, eps
        self.m = [torch.zeros_like(p) for p in params]
        self.v = [torch.zeros_like(p) for p in params]
        self.t = 0

    def step(self, grads):
        self.t += 1
        b1, b2 = self.betas
        bc1 = 1 - b1 ** self.t
        bc2 = 1 - b2 ** self.t
        with torch.no_grad():
            for p, g, m, v in zip(self.params, grads, self.m, self.v):
                if g is None:
                    continue
                m.mul_(b1).add_(g, alpha=1 - b1)
                v.mul_(b2).addcmul_(g, g, value=1 - b2)
                denom = (v.sqrt() / math.sqrt(bc2)).add_(self.eps)
                p.addcdiv_(m, denom, value=-self.lr / bc1)


class OracleCycleGAN:
    """EnhancedCycleGAN.train_step restated (enhanced_train.py:59-131) on plain state dicts.
    On CPU the reference's autocast/GradScaler are disabled, so this is its exact fp32 math.
    Quirks kept: D grads from the generator phase leak into D .grad but are zeroed at the next
    step (:67, set_to_none) -- so they never reach d_optimizer; 10 D forwards per step each
    advance the spectral-norm power iteration in call order (:70-77, :98-99, :110-113)."""

    LAMBDA_CYCLE, LAMBDA_IDT, LAMBDA_STRUCT = 10.0, 2.0, 0.5  # enhanced_train.py:55-57

    def __init__(self, sd_G_AB, sd_G_BA, sd_D_A, sd_D_B):
        def own(sd):
            return {k: v.detach().clone() for k, v in sd.items()}

        self.G_AB, self.G_BA, self.D_A, self.D_B = own(sd_G_AB), own(sd_G_BA), own(sd_D_A), own(sd_D_B)
        for sd in (self.G_AB, self.G_BA):
            for v in sd.values():
                v.requires_grad_(True)
        for sd in (self.D_A, self.D_B):
            for k, v in sd.items():
                if k.endswith("weight_orig") or k.endswith("bias"):
                    v.requires_grad_(True)
        self.g_keys = list(self.G_AB.keys())
        self.d_keys = [k for k in self.D_A.keys() if k.endswith("weight_orig") or k.endswith("bias")]
        self.g_opt = AdamState([self.G_AB[k] for k in self.g_keys] + [self.G_BA[k] for k in self.g_keys], 5e-5)
        self.d_opt = AdamState([self.D_A[k] for k in self.d_keys] + [self.D_B[k] for k in self.d_keys], 2e-4)

    def _D(self, sd, x):
        score, struct, new_uv = discriminator_forward(sd, x, training=True)
        for name, (u, v) in new_uv.items():
            sd[f"{name}.weight_u"] = u.detach()
            sd[f"{name}.weight_v"] = v.detach()
        return score, struct

    def train_step(self, real_A, real_B):
        mse = lambda a, b: ((a - b) ** 2).mean()
        l1 = lambda a, b: (a - b).abs().mean()
        fake_B = generator_forward(self.G_AB, real_A)
        fake_A = generator_forward(self.G_BA, real_B)

        ra, _ = self._D(self.D_A, real_A)
        rb, _ = self._D(self.D_B, real_B)
        d_real = (mse(ra, torch.ones_like(ra)) + mse(rb, torch.ones_like(rb))) * 0.5
        fa, _ = self._D(self.D_A, fake_A.detach())
        fb, _ = self._D(self.D_B, fake_B.detach())
        d_fake = (mse(fa, torch.zeros_like(fa)) + mse(fb, torch.zeros_like(fb))) * 0.5
        d_loss = d_real + d_fake
        d_params = self.d_opt.params
        d_grads = torch.autograd.grad(d_loss, d_params, allow_unused=True)
        self.d_opt.step(d_grads)

        idt_A = generator_forward(self.G_BA, real_A)
        idt_B = generator_forward(self.G_AB, real_B)
        identity_loss = (l1(idt_A, real_A) + l1(idt_B, real_B)) * self.LAMBDA_IDT
        fa, _ = self._D(self.D_A, fake_A)
        fb, _ = self._D(self.D_B, fake_B)
        g_loss = mse(fa, torch.ones_like(fa)) + mse(fb, torch.ones_like(fb))
        recon_A = generator_forward(self.G_BA, fake_B)
        recon_B = generator_forward(self.G_AB, fake_A)
        cycle_loss = (l1(recon_A, real_A) + l1(recon_B, real_B)) * self.LAMBDA_CYCLE
        _, ras = self._D(self.D_A, real_A)
        _, fas = self._D(self.D_A, fake_A)
        _, rbs = self._D(self.D_B, real_B)
        _, fbs = self._D(self.D_B, fake_B)
        structure_loss = (l1(ras, fas) + l1(rbs, fbs)) * self.LAMBDA_STRUCT
        total = g_loss + cycle_loss + identity_loss + structure_loss
        g_grads = torch.autograd.grad(total, self.g_opt.params, allow_unused=True)
        self.g_opt.step(g_grads)
        self.last_g_grads = g_grads
        self.last_d_grads = d_grads
        return {"d_loss": d_loss.item(), "g_loss": g_loss.item(), "cycle_loss": cycle_loss.item(),
                "identity_loss": identity_loss.item(), "structure_loss": structure_loss.item()}


# --------------------------------------------------------------------------------------------
# multi-style blends (output space -- the only "style weights" the reference has)
# --------------------------------------------------------------------------------------------
def blend_outputs(ys, w, x=None, w_x=0.0, gain=1.0, clip=None):
    """out = gain * (sum_s w_s y_s + w_x * x), optionally clipped.
    advanced_transform.py:206-213 (weights [0.2,0.3,0.5], gain 1.1, clip [0,1] on the (y+1)/2
    images); direct_transform.py:155-165 (y*w + (x*2-1)*(1-w), x already in [-1,1] here)."""
    out = sum(wi * yi for wi, yi in zip(w, ys))
    if x is not None:
        out = out + w_x * x
    out = out * gain
    if clip is not None:
        out = out.clamp(*clip)
    return out


def to_uint8_image(y):
    """(y+1)/2 -> clamp(0,1) -> *255 -> uint8 (truncation), direct_transform.py:66-71."""
    return (((y + 1.0) / 2.0).clamp(0, 1) * 255).to(torch.uint8)


def letterbox_normalize(img_u8, H, W, off_y, off_x, fill=255):
    """batch_process_images.py:193-205 (and :270-291): paste the resized uint8 image [h,w,3] on a white HxW canvas at the
    centring offsets, then transforms.ToTensor() + Normalize((0.5,)*3, (0.5,)*3).  Returns (fp32 [3,H,W], uint8 canvas [H,W,3]).
    (The LANCZOS resize in front of it is PIL's, host side, out of scope.)"""
    h, w, _ = img_u8.shape
    canvas = torch.full((H, W, 3), fill, dtype=torch.uint8)
    canvas[off_y:off_y + h, off_x:off_x + w] = img_u8
    t = canvas.permute(2, 0, 1).to(torch.float32).div(255)          # ToTensor
    return (t - 0.5) / 0.5, canvas                                   # Normalize


def strength_blend_u8(orig_u8, styled_u8, strength):
    """batch_process_images.py:304-310 ('simple' mode; gan_login_gui.py:826): numpy float64 arithmetic, clip, truncating astype.
    orig_u8, styled_u8: uint8 [H,W,3] arrays / tensors."""
    import numpy as np
    o, s = np.asarray(orig_u8), np.asarray(styled_u8)
    r = o * (1 - strength) + s * strength
    return torch.from_numpy(np.clip(r, 0, 255).astype(np.uint8))


# --------------------------------------------------------------------------------------------
# pretrain.Generator (pretrain.py:60-97): 4x[Conv4x4 s2 (+BN) + LeakyReLU 0.2] -> 4x[ConvT4x4 s2 (+BN) + ReLU], tanh
# --------------------------------------------------------------------------------------------
PRETRAIN_ENC = ((0, None), (2, 3), (5, 6), (8, 9))      # (conv index, BatchNorm index) inside `encoder`
PRETRAIN_DEC = ((0, 1), (3, 4), (6, 7), (9, None))      # (convT index, BatchNorm index) inside `decoder`


def pretrain_generator_forward(sd, x, training=True, momentum=0.1, eps=1e-5):
    """Functional restatement of pretrain.Generator.forward (pretrain.py:93-96) on a state_dict.  In training mode
    BatchNorm uses batch statistics and the returned dict holds the UPDATED running statistics
    (torch.nn.BatchNorm2d semantics: biased variance to normalise, unbiased for the running estimate)."""
    new = {}

    def bn(h, pre):
        rm, rv = sd[pre + ".running_mean"].clone(), sd[pre + ".running_var"].clone()
        y = F.batch_norm(h, rm, rv, sd[pre + ".weight"], sd[pre + ".bias"], training, momentum, eps)
        if training:
            new[pre + ".running_mean"], new[pre + ".running_var"] = rm, rv
            new[pre + ".num_batches_tracked"] = sd[pre + ".num_batches_tracked"] + 1
        return y

    h = x
    for ci, bi in PRETRAIN_ENC:
        h = F.conv2d(h, sd[f"encoder.{ci}.weight"], sd[f"encoder.{ci}.bias"], stride=2, padding=1)
        if bi is not None:
            h = bn(h, f"encoder.{bi}")
        h = F.leaky_relu(h, 0.2)
    for ci, bi in PRETRAIN_DEC:
        h = F.conv_transpose2d(h, sd[f"decoder.{ci}.weight"], sd[f"decoder.{ci}.bias"], stride=2, padding=1)
        if bi is not None:
            h = torch.relu(bn(h, f"decoder.{bi}"))
    return torch.tanh(h), new


def pretrain_masked_l1(generated, real, mask):
    """pretrain.py:160: nn.L1Loss()(generated * (1 - masks), real * (1 - masks))."""
    return (generated * (1 - mask) - real * (1 - mask)).abs().mean()


# --------------------------------------------------------------------------------------------
# Gram / VGG style loss (north_star addition; NOT in the reference, parity unpinned)
# --------------------------------------------------------------------------------------------
VGG19_CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512]
VGG19_TAPS = (1, 6, 11, 20, 29)  # relu1_1, relu2_1, relu3_1, relu4_1, relu5_1 in features[:30]


def vgg19_features(weights, x):
    """torchvision vgg19().features[:30] restated functionally.  ``weights``: list of (w, b) for
    the 13 convs up to conv5_1.  Returns the five tap feature maps."""
    taps, wi, idx = [], 0, 0
    for c in VGG19_CFG:
        if c == "M":
            x = F.max_pool2d(x, 2, 2)
            idx += 1
        else:
            w, b = weights[wi]
            wi += 1
            x = torch.relu(F.conv2d(x, w, b, padding=1))
            idx += 2
            if idx - 1 in VGG19_TAPS:
                taps.append(x)
    return taps


def gram(feat):
    """G = F F^T / (C*H*W), F = feat.view(B, C, H*W)  (Johnson et al. normalisation)."""
    B, C, H, W = feat.shape
    f = feat.reshape(B, C, H * W)
    return f @ f.transpose(1, 2) / (C * H * W)


def style_loss(feats, target_feats):
    """L_style = sum_l mean((G_l(y) - G_l(s))^2)."""
    return sum(((gram(a) - gram(b)) ** 2).mean() for a, b in zip(feats, target_feats))
