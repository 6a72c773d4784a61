"""Drop-in for the reference's ``EnhancedCycleGAN`` (enhanced_train.py:13-152): same attributes
(G_AB, G_BA, D_A, D_B, g_optimizer, d_optimizer, device), same ``train_step(real_A, real_B) ->
dict`` with the five loss keys, same ``save_models`` checkpoint files.

Differences that are deliberate:
  * mixed precision is bf16 (no GradScaler needed) instead of fp16 autocast + GradScaler
    (enhanced_train.py:46,61,88); ``precision="fp32"`` gives the parity mode.
  * parameters of each optimizer live in ONE flat fp32 buffer (views), gradients likewise: one
    fused Adam launch and -- when torch.distributed is initialised -- one NCCL all-reduce per
    optimizer per step (data parallel over NVLink; IN statistics are per-sample so no sync-norm).
  * the reference leaks generator-phase gradients into D's .grad and discards them at the next
    zero_grad (:67); here they are simply not computed.
  * the data loop / dataset (enhanced_train.py:154-208) is out of scope: inputs are tensors.
  * optional extension (north_star / BASELINE config 4, not in the reference): ``style_loss=GramStyleLoss(...)``
    with ``lambda_style > 0`` adds lambda_style * (L_gram(fake_B | style real_B) + L_gram(fake_A | style real_A))
    to the generator objective and a sixth key ``style_loss`` to the returned dict.  Default: off, so the
    default step is the reference's.
"""
from pathlib import Path

import torch
import torch.distributed as dist

from . import ops
from .enhanced_generator import EnhancedDiscriminator, EnhancedGenerator
from .losses import l1, mse_to_const


def allreduce_flat_(flat_grad):
    """Data-parallel gradient exchange: ONE sum all-reduce over the flat gradient buffer (NCCL over
    NVLink on the GPUs; any initialised backend works).  Returns the scale (1/world) the fused Adam
    folds into its update.  No-op (returns 1.0) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
        return 1.0 / dist.get_world_size()
    return 1.0


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam(lr, betas=(0.5,0.999), eps=1e-8) semantics (enhanced_train.py:36-43) as one
    kernel launch over a flat parameter / gradient buffer.  Parameters are re-pointed to views of
    the flat buffer at construction."""

    def __init__(self, params, lr, betas=(0.5, 0.999), eps=1e-8):
        params = list(params)
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        dev = params[0].device
        n = sum(p.numel() for p in params)
        self.flat = torch.empty(n, device=dev, dtype=torch.float32)
        self.flat_grad = torch.zeros(n, device=dev, dtype=torch.float32)
        self.exp_avg = torch.zeros(n, device=dev, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(n, device=dev, dtype=torch.float32)
        off = 0
        for p in params:
            k = p.numel()
            self.flat[off:off + k].copy_(p.detach().reshape(-1))
            p.data = self.flat[off:off + k].view_as(p)
            p.grad = self.flat_grad[off:off + k].view_as(p)
            off += k
        self.step_count = 0
        self.step_dev = torch.zeros(1, device=dev, dtype=torch.int32)      # the same count on the device (graph replays)
        self._params = params

    def zero_grad(self, set_to_none=False):
        self.flat_grad.zero_()   # keep the views alive: autograd accumulates into them in place

    @torch.no_grad()
    def step(self, closure=None, grad_scale=1.0):
        self.step_count += 1
        g = self.param_groups[0]
        # the bias corrections come from a device-side step counter, so a captured train step replays correctly
        ops.adam_step_dev(self.flat, self.flat_grad, self.exp_avg, self.exp_avg_sq, g["lr"], g["betas"][0],
                          g["betas"][1], g["eps"], self.step_dev, grad_scale)
        self.bump_versions()

    def bump_versions(self):
        """The kernel wrote through raw pointers: bump the version counter of every parameter (a Parameter whose .data was
        re-pointed keeps its own counter, separate from the flat buffer's) so packed-weight caches and captured graphs
        keyed on (data_ptr, _version) see the update without an explicit invalidate call."""
        torch.autograd.graph.increment_version([self.flat] + self._params)


class EnhancedCycleGAN:
    def __init__(self, pretrained_path=None, channels=16, num_transformer_blocks=1, precision="bf16",
                 device=None, style_loss=None, lambda_style=0.0, use_graph=False, graph_warmup=2):
        if not torch.cuda.is_available():
            raise RuntimeError("EnhancedCycleGAN (msg_b200): a B200 GPU is required; there is no CPU path")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        c, nb = channels, num_transformer_blocks           # reference: 16 / 1 (enhanced_train.py:18-21)
        self.G_AB = EnhancedGenerator(channels=c, num_transformer_blocks=nb).to(self.device)
        self.G_BA = EnhancedGenerator(channels=c, num_transformer_blocks=nb).to(self.device)
        self.D_A = EnhancedDiscriminator(channels=c).to(self.device)
        self.D_B = EnhancedDiscriminator(channels=c).to(self.device)
        self.G_AB.gradient_checkpointing_enable()          # :24-25
        self.G_BA.gradient_checkpointing_enable()
        if pretrained_path and Path(pretrained_path).exists():
            ckpt = torch.load(pretrained_path, map_location=self.device)
            self.G_AB.load_state_dict(ckpt["model_state_dict"], strict=False)   # :32-33
            self.G_BA.load_state_dict(ckpt["model_state_dict"], strict=False)
        self.set_precision(precision)
        self.lambda_cycle, self.lambda_identity, self.lambda_structure = 10.0, 2.0, 0.5   # :55-57
        self.style_loss, self.lambda_style = style_loss, float(lambda_style)
        if self.lambda_style > 0 and style_loss is None:
            raise ValueError("lambda_style > 0 needs a style_loss (multi_style_transfer_gan_b200.style_loss.GramStyleLoss)")
        self._build_optimizers()
        # use_graph: after `graph_warmup` eager steps the whole step (6 G + 10 D forwards, both backwards and both Adam updates:
        # ~2000 launches) is captured ONCE into a CUDA graph and replayed -- same kernels, same order, same results; the step was
        # launch-bound (74 ms of kernels in a 93 ms step).  Data parallel: the two NCCL all-reduces stay eager, between four
        # graph segments that share one memory pool.  Inputs must keep their shape.
        self.use_graph, self.graph_warmup = bool(use_graph), int(graph_warmup)
        self._graph, self._graph_io, self._eager_steps, self.graph_error, self._capture = None, None, 0, None, None

    def _build_optimizers(self):
        gp = [p for m in (self.G_AB, self.G_BA) for k, p in m.named_parameters()]
        dp = [p for m in (self.D_A, self.D_B) for p in m.parameters()]
        self.g_optimizer = FusedAdam(gp, lr=5e-5, betas=(0.5, 0.999))
        self.d_optimizer = FusedAdam(dp, lr=2e-4, betas=(0.5, 0.999))
        self._broadcast_replica_state()
        for m in (self.G_AB, self.G_BA):
            m.invalidate_packed_weights()
            # the generators' backward accumulates straight into the flat gradient buffer (generator_engine.GradSink)
            m._grad_sinks = {k: p.grad for k, p in m.named_parameters()}

    def _broadcast_replica_state(self):
        """Data parallel: every rank must start from rank 0's parameters, Adam state and spectral-norm u / v buffers
        (randomly initialised, enhanced_generator.py:269-271) -- otherwise the replicas silently train different models
        (the gradient all-reduce alone does not keep them together)."""
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            return
        with torch.cuda.device(self.device):
            for opt in (self.g_optimizer, self.d_optimizer):
                for t in (opt.flat, opt.exp_avg, opt.exp_avg_sq):
                    dist.broadcast(t, src=0)
                opt.bump_versions()
            for m in (self.D_A, self.D_B):
                for b in m.buffers():
                    dist.broadcast(b, src=0)

    def set_precision(self, precision):
        self.precision = precision
        for m in (self.G_AB, self.G_BA, self.D_A, self.D_B):
            m.set_precision(precision)

    def load_state_dicts(self, G_AB=None, G_BA=None, D_A=None, D_B=None):
        """Loads weights IN PLACE (the flat optimizer buffers stay the parameter storage)."""
        for m, sd in ((self.G_AB, G_AB), (self.G_BA, G_BA), (self.D_A, D_A), (self.D_B, D_B)):
            if sd is not None:
                m.load_state_dict(sd, strict=True)
        for m in (self.G_AB, self.G_BA):
            m.invalidate_packed_weights()

    @staticmethod
    def _dp():
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    def _boundary(self, op):
        """While the step is being captured: end the current graph segment, remember what the replay has to do eagerly at this
        point, start the next segment (same memory pool, so tensors cross freely).  The NCCL collectives stay OUT of the graphs."""
        cap = self._capture
        cap["graphs"][-1].capture_end()
        cap["ops"].append(op)
        nxt = torch.cuda.CUDAGraph()
        nxt.capture_begin(pool=cap["graphs"][0].pool())
        cap["graphs"].append(nxt)

    def _sync_begin(self, opt):
        """Starts the sum all-reduce of the optimizer's flat gradient buffer (NCCL's own stream, after everything enqueued so far)
        and returns a token for _sync_end; independent work issued in between runs under the collective."""
        segment = self._dp() or getattr(self, "segment_at_syncs", False)     # (segment_at_syncs: one-GPU tests of the segmented replay)
        if self._capture is not None:
            if segment:
                self._boundary(("begin", opt))
            return None
        return dist.all_reduce(opt.flat_grad, op=dist.ReduceOp.SUM, async_op=True) if self._dp() else None

    def _sync_end(self, token):
        """Waits for the all-reduce started by _sync_begin; returns the 1/world scale the fused Adam folds into its update."""
        segment = self._dp() or getattr(self, "segment_at_syncs", False)
        if self._capture is not None:
            if segment:
                self._boundary(("wait", None))
        elif token is not None:
            token.wait()
        return 1.0 / dist.get_world_size() if self._dp() else 1.0

    def _sync_grads(self, opt):
        """One blocking NCCL sum all-reduce over the optimizer's flat gradient buffer; returns the 1/world scale."""
        if self._capture is not None:
            if self._dp() or getattr(self, "segment_at_syncs", False):
                self._boundary(("sync", opt))
        elif self._dp():
            dist.all_reduce(opt.flat_grad, op=dist.ReduceOp.SUM)
        return 1.0 / dist.get_world_size() if self._dp() else 1.0

    def train_step(self, real_A, real_B):
        """reference: enhanced_train.py:59-131 (order of the 6 G and 10 D forwards preserved, so the
        spectral-norm power iterations advance exactly as in the reference)."""
        with torch.cuda.device(self.device):
            if not self.use_graph:
                keys, vec = self._step_body(real_A.to(self.device, non_blocking=True), real_B.to(self.device, non_blocking=True))
                return dict(zip(keys, vec.tolist()))          # one sync instead of five .item()
            return self._train_step_graphed(real_A, real_B)

    def _train_step_graphed(self, real_A, real_B):
        from . import _lib
        io = self._graph_io
        if io is not None and (io["A"].shape != real_A.shape or io["B"].shape != real_B.shape):
            self._graph, self._graph_io, self._eager_steps = None, None, 0       # new geometry: warm up and capture again
            io = None
        if self._graph is None:
            if self._eager_steps < self.graph_warmup or self.graph_error is not None:
                self._eager_steps += 1
                keys, vec = self._step_body(real_A.to(self.device, non_blocking=True), real_B.to(self.device, non_blocking=True))
                return dict(zip(keys, vec.tolist()))
            io = {"A": torch.empty(real_A.shape, device=self.device, dtype=torch.float32),
                  "B": torch.empty(real_B.shape, device=self.device, dtype=torch.float32)}
            io["A"].copy_(real_A, non_blocking=True)
            io["B"].copy_(real_B, non_blocking=True)
            torch.cuda.synchronize()
            l0 = _lib.launches
            cap = {"graphs": [torch.cuda.CUDAGraph()], "ops": []}
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream())
            try:
                with torch.cuda.stream(side):
                    cap["graphs"][0].capture_begin()
                    self._capture = cap
                    try:
                        io["keys"], io["vec"] = self._step_body(io["A"], io["B"])
                    finally:
                        self._capture = None
                        cap["graphs"][-1].capture_end()
                torch.cuda.current_stream().wait_stream(side)
            except Exception as e:              # capture refused: stay eager, loudly
                self.graph_error = repr(e)
                torch.cuda.synchronize()
                for o in (self.g_optimizer, self.d_optimizer):
                    o.step_count = int(o.step_dev.item())
                import warnings
                warnings.warn(f"EnhancedCycleGAN: CUDA-graph capture of the train step failed ({e!r}); running eagerly")
                return self._train_step_graphed(real_A, real_B)
            io["launches"] = _lib.launches - l0
            _lib.launches = l0
            for o in (self.g_optimizer, self.d_optimizer):
                o.step_count -= 1               # the capture did not execute anything
            self._graph, self._graph_io = cap, io
        else:
            io["A"].copy_(real_A, non_blocking=True)
            io["B"].copy_(real_B, non_blocking=True)
        # replay: segment, [eager NCCL op, segment]*
        cap = self._graph
        pending = None
        for i, gr in enumerate(cap["graphs"]):
            if i > 0:
                kind, opt = cap["ops"][i - 1]
                if kind == "begin":
                    pending = dist.all_reduce(opt.flat_grad, op=dist.ReduceOp.SUM, async_op=True) if self._dp() else None
                elif kind == "sync":
                    if self._dp():
                        dist.all_reduce(opt.flat_grad, op=dist.ReduceOp.SUM)
                elif pending is not None:
                    pending.wait()
                    pending = None
            gr.replay()
        _lib.launches += io["launches"]          # the replay launched exactly the kernels the capture recorded
        for o in (self.g_optimizer, self.d_optimizer):
            o.step_count += 1
            o.bump_versions()
        return dict(zip(io["keys"], io["vec"].tolist()))

    def _step_body(self, real_A, real_B):
        """one train step on device tensors; returns (loss keys, stacked loss tensor)"""
        G_AB, G_BA, D_A, D_B = self.G_AB, self.G_BA, self.D_A, self.D_B
        fake_B = G_AB(real_A)
        fake_A = G_BA(real_B)

        # ---- discriminators (:66-85)
        self.d_optimizer.zero_grad()
        for m in (D_A, D_B):
            m.requires_grad_(True)
        real_A_score, _ = D_A(real_A)
        real_B_score, _ = D_B(real_B)
        d_real_loss = (mse_to_const(real_A_score, 1.0) + mse_to_const(real_B_score, 1.0)) * 0.5
        fake_A_score, _ = D_A(fake_A.detach())
        fake_B_score, _ = D_B(fake_B.detach())
        d_fake_loss = (mse_to_const(fake_A_score, 0.0) + mse_to_const(fake_B_score, 0.0)) * 0.5
        d_loss = d_real_loss + d_fake_loss
        d_loss.backward()
        # Data parallel: the all-reduce of the discriminator gradients runs UNDER the two identity forwards of the generator phase,
        # which do not touch the discriminators (same arithmetic as the reference's order -- D step first -- only the launch order
        # differs); the D update waits for it.
        d_sync = self._sync_begin(self.d_optimizer)

        # ---- generators (:87-123)
        self.g_optimizer.zero_grad()
        for m in (D_A, D_B):
            m.requires_grad_(False)        # no leaked D grads (see module docstring)
        idt_A = G_BA(real_A)
        idt_B = G_AB(real_B)
        identity_loss = (l1(idt_A, real_A) + l1(idt_B, real_B)) * self.lambda_identity
        self.d_optimizer.step(grad_scale=self._sync_end(d_sync))
        fake_A_score, _ = D_A(fake_A)
        fake_B_score, _ = D_B(fake_B)
        g_loss = mse_to_const(fake_A_score, 1.0) + mse_to_const(fake_B_score, 1.0)
        recon_A = G_BA(fake_B)
        recon_B = G_AB(fake_A)
        cycle_loss = (l1(recon_A, real_A) + l1(recon_B, real_B)) * self.lambda_cycle
        _, real_A_struct = D_A(real_A)
        _, fake_A_struct = D_A(fake_A)
        _, real_B_struct = D_B(real_B)
        _, fake_B_struct = D_B(fake_B)
        structure_loss = (l1(real_A_struct, fake_A_struct) + l1(real_B_struct, fake_B_struct)) * self.lambda_structure
        total_g_loss = g_loss + cycle_loss + identity_loss + structure_loss
        style_term = None
        if self.lambda_style > 0:
            # Gram style term on the translated images against the real images of the target domain (the targets
            # are captured by the autograd node, so re-pointing the style between the two calls is safe)
            sl = self.style_loss
            style_term = (sl.set_style(real_B)(fake_B) + sl.set_style(real_A)(fake_A)) * self.lambda_style
            total_g_loss = total_g_loss + style_term
        total_g_loss.backward()
        self.g_optimizer.step(grad_scale=self._sync_grads(self.g_optimizer))
        for m in (D_A, D_B):
            m.requires_grad_(True)
        for m in (G_AB, G_BA):
            m.invalidate_packed_weights()

        outs = [d_loss.detach(), g_loss.detach(), cycle_loss.detach(), identity_loss.detach(), structure_loss.detach()]
        keys = ["d_loss", "g_loss", "cycle_loss", "identity_loss", "structure_loss"]
        if style_term is not None:
            outs.append(style_term.detach())
            keys.append("style_loss")
        return keys, torch.stack(outs)

    def save_models(self, save_dir, epoch):
        """reference: enhanced_train.py:133-152 (same file names and dict keys)."""
        save_path = Path(save_dir)
        save_path.mkdir(parents=True, exist_ok=True)
        torch.save({"epoch": epoch, "G_AB_state_dict": self.G_AB.state_dict()}, save_path / f"G_AB_epoch_{epoch}.pth")
        torch.save({"epoch": epoch, "G_BA_state_dict": self.G_BA.state_dict()}, save_path / f"G_BA_epoch_{epoch}.pth")
        torch.save({"epoch": epoch, "D_A_state_dict": self.D_A.state_dict(), "D_B_state_dict": self.D_B.state_dict()},
                   save_path / f"discriminators_epoch_{epoch}.pth")
