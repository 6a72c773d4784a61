"""Forward / backward schedule of the EnhancedDiscriminator (enhanced_generator.py:230-275) on the
msg_b200 kernels, including the old-style spectral norm applied to all seven convs (:269-271):
one power iteration per training-mode forward (weight_u / weight_v updated in place), sigma =
u^T W v, weight = weight_orig / sigma -- folded into the weight-packing kernel.
"""
import torch
import torch.nn.functional as F

from . import ops
from .ops import ACT_LRELU, ACT_NONE, ConvGeom

D_CONVS = ("main.0", "main.2", "main.5", "main.8", "batch_head.0", "structure_head.0", "structure_head.3")


class DiscriminatorEngine:
    def __init__(self, channels):
        c = self.c = channels
        g = self.geom = {}
        g["main.0@4"] = ConvGeom("conv", 4, c, 4, 2, 1)     # image 3 -> 4 channels (fp32)
        g["main.0@8"] = ConvGeom("conv", 8, c, 4, 2, 1)     # image 3 -> 8 channels (bf16 / tcgen05)
        g["main.2"] = ConvGeom("conv", c, 2 * c, 4, 2, 1)
        g["main.5"] = ConvGeom("conv", 2 * c, 4 * c, 4, 2, 1)
        g["main.8"] = ConvGeom("conv", 4 * c, 8 * c, 4, 2, 1)
        g["batch_head.0"] = ConvGeom("conv", 8 * c, 1, 4, 1, 1)
        g["structure_head.0"] = ConvGeom("conv", 8 * c, 8 * c, 3, 1, 1)
        g["structure_head.3"] = ConvGeom("conv", 8 * c, 1, 4, 1, 1)

    @staticmethod
    def cin_pad(dtype):
        return 4 if dtype == torch.float32 else 8

    def _g(self, name, dtype):
        return self.geom[f"main.0@{self.cin_pad(dtype)}"] if name == "main.0" else self.geom[name]

    def _master(self, w, name, dtype):
        if name == "main.0":
            w = F.pad(w, [0, 0, 0, 0, 0, self.cin_pad(dtype) - w.shape[1]])
        return w.contiguous()

    def _sigma(self, P, name, training):
        """Runs the power iteration IN PLACE on the u/v buffers (as the reference's forward
        pre-hook does) and returns (sigma[1], u_snapshot, v_snapshot)."""
        w = P[f"{name}.weight_orig"]
        u, v = P[f"{name}.weight_u"], P[f"{name}.weight_v"]
        rows, cols = w.shape[0], w.numel() // w.shape[0]
        sigma = torch.empty(1, device=w.device, dtype=torch.float32)
        ops.spectral_norm(w.detach(), rows, cols, u, v, training, sigma)
        return sigma, u.clone(), v.clone()

    def _sigmas(self, P, training, save):
        """The seven power iterations of one forward (IN PLACE on the u / v buffers, as the reference's forward pre-hooks do) in ONE
        launch; returns {conv: (sigma[1], u snapshot, v snapshot)} -- the snapshots only when the backward needs them."""
        w0 = P[f"{D_CONVS[0]}.weight_orig"]
        sig = torch.empty(len(D_CONVS), device=w0.device, dtype=torch.float32)
        probs = []
        for i, n in enumerate(D_CONVS):
            w = P[f"{n}.weight_orig"].detach()
            probs.append((w, w.shape[0], w.numel() // w.shape[0], P[f"{n}.weight_u"], P[f"{n}.weight_v"], sig[i:i + 1]))
        ops.spectral_norm_batched(probs, training)
        out = {}
        for i, n in enumerate(D_CONVS):
            u, v = P[f"{n}.weight_u"], P[f"{n}.weight_v"]
            out[n] = (sig[i:i + 1], u.clone() if save else u, v.clone() if save else v)
        return out

    def forward(self, P, x, dtype, training, save):
        """x: fp32 NCHW [N,3,H,W] -> (score_map fp32 [N] (plane mean of the batch head),
        struct fp32 [N,1,h,w], saved)."""
        g = {n: self._g(n, dtype) for n in D_CONVS}
        N, Cx, H, W = x.shape
        if Cx != 3 or H % 16 or W % 16 or H < 32 or W < 32:
            raise RuntimeError(f"EnhancedDiscriminator: expected [N,3,H,W] with H,W multiples of 16 and >= 32, got {tuple(x.shape)}")
        sn = self._sigmas(P, training, save)

        def wp(n):
            return g[n].pack_fwd(self._master(P[f"{n}.weight_orig"].detach(), n, dtype), dtype, sn[n][0])

        def bias(n):
            return P[f"{n}.bias"].detach().contiguous()

        x0 = ops.nchw_to_nhwc(x, dtype, self.cin_pad(dtype))
        h1 = g["main.0"].forward(x0, wp("main.0"), bias("main.0"), act=ACT_LRELU)
        acts = {"x0": x0, "h1": h1}
        a = h1
        for n in ("main.2", "main.5", "main.8"):
            st = ops.new_stats(N, g[n].Cout, x.device)
            y = g[n].forward(a, wp(n), bias(n), stats=st)
            acts[f"in:{n}"] = a
            a = ops.instnorm_apply(y, st, ACT_LRELU)
            acts[f"y:{n}"], acts[f"st:{n}"] = y, st
        feat = a
        bh = g["batch_head.0"].forward(feat, wp("batch_head.0"), bias("batch_head.0"))
        score = ops.avgpool_fwd(bh).reshape(N)                      # AdaptiveAvgPool2d(1), :257
        st = ops.new_stats(N, 8 * self.c, x.device)
        ys = g["structure_head.0"].forward(feat, wp("structure_head.0"), bias("structure_head.0"), stats=st)
        s1 = ops.instnorm_apply(ys, st, ACT_LRELU)
        so = g["structure_head.3"].forward(s1, wp("structure_head.3"), bias("structure_head.3"))
        struct = so.float().reshape(N, 1, so.shape[1], so.shape[2]) if so.dtype != torch.float32 else so.reshape(N, 1, so.shape[1], so.shape[2])
        saved = None
        if save:
            acts.update(feat=feat, bh_shape=tuple(bh.shape), ys=ys, sts=st, s1=s1, so_shape=tuple(so.shape))
            saved = {"acts": acts, "sn": sn}
        return score, struct, saved

    def backward(self, P, saved, dscore, dstruct, dtype, need_dx, need_dw):
        """dscore fp32 [N], dstruct fp32 [N,1,h,w] (either may be None).  Returns (dx NCHW | None,
        grads {state_dict key: fp32})."""
        g, acts, sn = {n: self._g(n, dtype) for n in D_CONVS}, saved["acts"], saved["sn"]
        G = {}

        def conv_bwd(n, x, dy, need_dx=True):
            geom = g[n]
            w_orig = P[f"{n}.weight_orig"].detach()
            sigma, u, v = sn[n]
            if need_dw:
                m = self._master(w_orig, n, dtype)
                dw = torch.zeros_like(m)
                db = torch.zeros(geom.Cout, device=m.device, dtype=torch.float32)
                # (biases of the convs in front of an InstanceNorm: exactly zero gradient, see generator_engine._conv_bwd)
                geom.wgrad(x, dy, dw, None if n in ("main.2", "main.5", "main.8", "structure_head.0") else db)
                if n == "main.0":
                    dw = dw[:, :3].contiguous()
                dwo = torch.zeros_like(w_orig)
                rows, cols = w_orig.shape[0], w_orig.numel() // w_orig.shape[0]
                ops.spectral_norm_bwd(dw, w_orig.contiguous(), u, v, sigma, rows, cols, dwo)
                G[f"{n}.weight_orig"] = dwo
                G[f"{n}.bias"] = db
            if not need_dx:
                return None
            wpd = geom.pack_dgrad(self._master(w_orig, n, dtype), dtype, sigma)
            return geom.dgrad(dy, wpd, x.shape[1:3])

        feat = acts["feat"]
        dfeat = None
        if dstruct is not None:
            dso = dstruct.reshape(acts["so_shape"]).to(dtype).contiguous()
            ds1 = conv_bwd("structure_head.3", acts["s1"], dso)
            dys = ops.instnorm_bwd(acts["ys"], acts["sts"], ds1, ACT_LRELU)
            dfeat = conv_bwd("structure_head.0", feat, dys)
        if dscore is not None:
            dbh = ops.avgpool_bwd(dscore.reshape(-1, 1).float(), acts["bh_shape"], dtype)
            d2 = conv_bwd("batch_head.0", feat, dbh)
            dfeat = d2 if dfeat is None else ops.add(dfeat, d2)
        da = dfeat
        for n in ("main.8", "main.5", "main.2"):
            dy = ops.instnorm_bwd(acts[f"y:{n}"], acts[f"st:{n}"], da, ACT_LRELU)
            da = conv_bwd(n, acts[f"in:{n}"], dy)
        dh1 = ops.act_bwd(acts["h1"], da, ACT_LRELU)
        dx0 = conv_bwd("main.0", acts["x0"], dh1, need_dx=need_dx)
        dx = ops.nhwc_to_nchw(dx0, 3) if need_dx else None
        return dx, G
