#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config, on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload ...]

Workload at N=1 (configs[1]): batch stylisation, 64 synthetic 512x512 images per step, c=64 / 3-block
EnhancedGenerator, 3 styles (3 state dicts, seeds 0..2) blended in output space with weights
[0.2, 0.3, 0.5] (advanced_transform.py:206-213) -- i.e. THREE generator forwards per stylised image,
bf16.  A step = one pass over the batch.  `value` = stylised images/s with inputs resident in HBM;
`e2e` = the same through the public API (MultiStyleStylizer) from pinned HOST fp32 images to HOST
uint8 results, copies inside the timed region.  N>1: one process per GPU (torchrun), images sharded by
rank, no collective on the data path (weak scaling: 64 images per GPU per step).

--impl reference: the reference's own CPU implementation of the path (its PyTorch modules restated in
oracle/restate.py -- /root/reference is not on the GPU box) on the host cores, on a bounded sample
(1 image x 3 styles per step).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "stylised_images_per_sec_512x512_3styles"
UNIT = "images/s"
STYLE_W = [0.2, 0.3, 0.5]
GEN_GFLOP_512 = 198.844      # generator forward, c=64, per 512x512 image
CONV_GFLOP_512 = 181.664     # conv + convT part; LocalAttention bmm = 17.180
IN_BYTES_512_BF16 = 520e6    # InstanceNorm algorithmic bytes per image (1 read + 1 write)


def synth_images(B, H, W, seed=1234):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(B, 3, H, W, generator=g) * 2 - 1


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for l in self.lines:
            f = [t.strip() for t in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle restatement of the reference's PyTorch modules on host cores
# --------------------------------------------------------------------------------------------------
def cpu_reference_images_per_sec(steps, warmup, H=512, W=512, c=64, nb=3):
    from oracle import restate as R
    from multi_style_transfer_gan_b200.enhanced_generator import EnhancedGenerator
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    sds = []
    for seed in range(3):
        torch.manual_seed(seed)
        sds.append({k: v.detach() for k, v in EnhancedGenerator(c, nb).state_dict().items()})  # init only, CPU
    x = synth_images(1, H, W)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            ys = [R.generator_forward(sd, x) for sd in sds]
            R.to_uint8_image(R.blend_outputs(ys, STYLE_W))
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    per_step = sum(times) / len(times)
    return 1.0 / per_step, per_step, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    v, per_step, cores = cpu_reference_images_per_sec(args.steps, args.warmup)
    line = {"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": "batch stylisation 512x512, c=64/3-block generator, 3 styles blended [0.2,0.3,0.5]",
                       "sample": "1 image x 3 styles per step (of the 64-image batch)"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "1 image x 3 generator forwards + blend per step, fp32, torch CPU (oracle/restate.py)"},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
CONV_DRAM_BYTES_PER_LAUNCH = 675.61e6   # profiles/r1_launch_list_summary.md (final code of round 1)


def run_ours(args):
    import torch.distributed as dist
    from multi_style_transfer_gan_b200 import _lib, ops, profiler
    from multi_style_transfer_gan_b200.enhanced_generator import EnhancedGenerator
    from multi_style_transfer_gan_b200.stylize import MultiStyleStylizer

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    H = W = args.size
    B = args.batch            # images per GPU per step
    c, nb = args.channels, 3 if args.channels == 64 else 1
    gens = []
    for seed in range(3):
        torch.manual_seed(seed)
        gens.append(EnhancedGenerator(c, nb).to(dev))
    sty = MultiStyleStylizer(gens, precision=args.precision, micro_batch=args.micro_batch)
    # per-launch CUDA-event breakdown (roofline): the same step with one stream and no graph replay, because launches
    # inside a replayed graph cannot be bracketed by events and concurrent kernels of different styles share the SMs
    sty_serial = MultiStyleStylizer(gens, precision=args.precision, micro_batch=args.micro_batch, style_streams=False,
                                    use_graph=False)
    x_host = synth_images(B, H, W, seed=1234 + rank).pin_memory()
    x_dev = x_host.to(dev)
    out_dev = torch.empty((B, 3, H, W), device=dev, dtype=torch.float32)
    out_host = torch.empty((B, 3, H, W), dtype=torch.uint8).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup, prof=False):
        for _ in range(warmup):
            fn()
        barrier()
        l0 = _lib.launches
        if prof:
            profiler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        breakdown = profiler.stop() if prof else None
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps, _lib.launches - l0, breakdown

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # inputs (64 x 512^2 fp32 = 201 MB) + the activations of each micro-batch exceed the 126 MB L2,
    # so nothing is L2-resident between timed iterations (no explicit flush needed).
    ms_dev, launches, _ = timed(lambda: sty(x_dev, STYLE_W, out=out_dev), args.steps, args.warmup)
    ms_serial, _, breakdown = timed(lambda: sty_serial(x_dev, STYLE_W, out=out_dev), args.steps, 1, prof=True)
    ms_e2e, _, _ = timed(lambda: (sty(x_host, STYLE_W, out_uint8=True, out=out_host), torch.cuda.current_stream().synchronize()),
                         max(2, args.steps // 2), 1)
    clocks = sampler.stop() if rank == 0 else None

    value = world * B / (ms_dev * 1e-3)
    e2e = world * B / (ms_e2e * 1e-3)
    pk = peaks()
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": f"batch stylisation {H}x{W}, batch {B} per GPU, 3 style weights {STYLE_W}, "
                                   f"c={c}/{nb}-block EnhancedGenerator, reference-faithful output blend = 3 generator "
                                   f"forwards per image ({3 * GEN_GFLOP_512 * (H * W) / 512 ** 2:.1f} GFLOP/image)",
                       "global_batch": B * world, "micro_batch": args.micro_batch, "schedule": f"style_streams={int(sty.style_streams)}, cuda_graph={int(sty.use_graph)}", "parallelism": f"image-sharded x{world}, no collective",
                       "l2": "inputs + per-micro-batch activations > 126 MB L2; no flush needed",
                       "weights": "random init (seeds 0,1,2), fp32 master, bf16 packed"},
            "gpu_launches": launches,
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": x_host.numel() * 4, "d2h_bytes_per_step": out_host.numel(),
                    "ms_per_step": ms_e2e},
            "clocks": clocks}
    if breakdown and rank == 0:
        # roofline of the dominant kernel family (tcgen05 conv): algorithmic conv FLOPs of the step
        # divided by the summed CUDA-event time of its launches inside the timed region.
        scale = (H * W) / 512 ** 2
        conv_ms = sum(breakdown.get(k, {}).get("ms", 0.0) for k in ("msg_conv2d", "msg_conv_slab", "msg_conv_shift")) / args.steps
        conv_launches = sum(breakdown.get(k, {}).get("launches", 0) for k in ("msg_conv2d", "msg_conv_slab", "msg_conv_shift")) // args.steps
        flops = 3 * B * CONV_GFLOP_512 * scale * 1e9
        if conv_ms > 0:
            ach = flops / (conv_ms * 1e-3) / 1e12
            line["roofline"] = {"bound": "tensor", "kernel": "conv_tma_kernel + conv_slab_kernel + conv_shift_kernel (every conv / convT launch of the step)",
                                "achieved": ach, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                                "frac": ach / pk["bf16_tflops_sustained"],
                                # dram__bytes_read.sum + dram__bytes_write.sum per conv launch (average over the 138 conv launches
                                # of the ncu window in profiles/r1_launches_final.csv, 16-image micro-batch at 512^2)
                                "traffic": CONV_DRAM_BYTES_PER_LAUNCH if (H == 512 and args.micro_batch == 16 and c == 64) else None,
                                "peak_source": pk["source"] + " (sustained bf16; kernel timed inside a long step)",
                                "launches_per_step": conv_launches, "ms_per_step": conv_ms,
                                "measured_on": f"serialised pass of the same step (one stream, no graph replay, {ms_serial:.1f} ms per step): "
                                               "launches inside a replayed graph cannot be bracketed by events"}
        in_ms = sum(breakdown.get(k, {}).get("ms", 0.0) for k in ("msg_instnorm_apply", "msg_instnorm_stats")) / args.steps
        if in_ms > 0:
            in_launches = sum(breakdown.get(k, {}).get("launches", 0) for k in ("msg_instnorm_apply", "msg_instnorm_stats")) // args.steps
            per_fwd = in_launches // (3 * max(1, (B + args.micro_batch - 1) // args.micro_batch))
            if per_fwd <= 5:
                # 8 of the 13 IN applies of a forward are fused into their 1x1 consumers (no HBM pass at all); the
                # stand-alone launches left are the initial IN (1 read + 1 write of [c, H, W]) and the four
                # MultiScaleBlock outputs (read + residual read + write): bytes of exactly those launches (bf16)
                c = args.channels
                elems = c * H * W * 2 + 3 * (2 * (2 * c) * (H // 2) * (W // 2) + (4 * c) * (H // 4) * (W // 4) + c * H * W)
                in_bytes, note = elems * 2, "5 stand-alone apply launches per forward (8 fused into 1x1 convs)"
            else:
                in_bytes, note = IN_BYTES_512_BF16 * scale, "13 apply launches per forward, 1 read + 1 write each (SURVEY 8d)"
            gbs = 3 * B * in_bytes / (in_ms * 1e-3) / 1e9
            line["roofline_instnorm"] = {"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s",
                                         "frac": gbs / pk["hbm_gbs"], "ms_per_step": in_ms, "bytes": note}
        line["breakdown_ms_per_step"] = {k: round(v["ms"] / args.steps, 3) for k, v in sorted(breakdown.items(), key=lambda kv: -kv[1]["ms"])}
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            v, per_step, cores = cpu_reference_images_per_sec(1, 1)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "1 image x 3 generator forwards + blend, fp32, torch CPU restatement of the "
                                              "reference modules (oracle/restate.py), 1 warm-up + 1 timed"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_train(args):
    """Secondary workload (BASELINE.json configs[3]): EnhancedCycleGAN.train_step, batch 8 per GPU at
    256x256, bf16, data-parallel (one flat NCCL all-reduce per optimizer per step)."""
    import torch.distributed as dist
    from multi_style_transfer_gan_b200 import _lib, profiler
    from multi_style_transfer_gan_b200.enhanced_train import EnhancedCycleGAN
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)                       # identical init on every rank
    c = args.channels
    style = None
    if args.lambda_style > 0:                  # BASELINE config 4: + VGG-19 Gram style term (random-init trunk, seed 0)
        from multi_style_transfer_gan_b200.style_loss import GramStyleLoss, VGG19Features
        style = GramStyleLoss(VGG19Features(dev, seed=0), precision=args.precision)
    m = EnhancedCycleGAN(channels=c, num_transformer_blocks=3 if c == 64 else 1, precision=args.precision, device=dev,
                         style_loss=style, lambda_style=args.lambda_style)
    B, S = args.train_batch, args.train_size
    A = synth_images(B, S, S, seed=11 + rank).pin_memory()
    Bm = synth_images(B, S, S, seed=12 + rank).pin_memory()
    for _ in range(args.warmup):
        m.train_step(A, Bm)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = _lib.launches
    profiler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        losses = m.train_step(A, Bm)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    breakdown = profiler.stop()
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        line = {"metric": "train_steps_per_sec", "value": 1e3 / ms, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
                "config": {"workload": f"EnhancedCycleGAN.train_step (2 G + 2 D, LSGAN + cycle + identity + structure losses"
                                       + (f" + {args.lambda_style:g} x VGG-19 Gram style loss (random-init trunk)" if style else "")
                                       + f", fused Adam), c={c}, batch {B} per GPU at {S}x{S}", "global_batch": B * world,
                           "parallelism": f"data-parallel x{world}, 2 flat NCCL all-reduces per step"},
                "gpu_launches": _lib.launches - l0, "losses": losses,
                "e2e": {"value": 1e3 / ms, "unit": "steps/s", "h2d_bytes_per_step": 2 * A.numel() * 4, "d2h_bytes_per_step": 20},
                "breakdown_ms_per_step": {k: round(v["ms"] / args.steps, 3) for k, v in sorted(breakdown.items(), key=lambda kv: -kv[1]["ms"])}}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--channels", type=int, default=64)
    ap.add_argument("--micro-batch", type=int, default=16)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="stylise", choices=["stylise", "train"])
    ap.add_argument("--train-batch", type=int, default=8)
    ap.add_argument("--train-size", type=int, default=256)
    ap.add_argument("--lambda-style", type=float, default=0.0,
                    help="train workload: weight of the VGG-19 Gram style term (0 = the reference's train_step)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (the msg_b200 path has no CPU fallback); use --impl reference for the CPU arm")
        (run_train if args.workload == "train" else run_ours)(args)


if __name__ == "__main__":
    main()
