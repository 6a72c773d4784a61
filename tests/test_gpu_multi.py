"""Two-GPU data-parallel train step over NCCL (skipped on a single-GPU box): one data-parallel step on two half batches
equals one single-GPU step on the concatenated batch (IN statistics are per sample, the losses are batch means, the flat
gradient all-reduce + 1/world in the fused Adam gives the global-batch gradient), the two replicas stay bit-identical, and
the step also runs as a CUDA-graph replay (four graph segments, the two all-reduces eager between them, the first one under the identity forwards)
(enhanced_train.py:59-131)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q, init, A, B, use_graph, steps):
    import faulthandler
    import torch.distributed as dist
    faulthandler.dump_traceback_later(150, exit=True)       # a hung rank prints its stacks and dies instead of eating the box's time
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from multi_style_transfer_gan_b200.enhanced_train import EnhancedCycleGAN
    torch.manual_seed(100 + rank)            # different construction seeds: the rank-0 broadcast must make the replicas equal
    m = EnhancedCycleGAN(channels=8, num_transformer_blocks=1, precision="fp32", device=f"cuda:{rank}", use_graph=use_graph,
                         graph_warmup=1)
    if rank == 0:
        m.load_state_dicts(**init)
    m._broadcast_replica_state()
    per = A.shape[0] // world
    losses = [m.train_step(A[rank * per:(rank + 1) * per], B[rank * per:(rank + 1) * per]) for _ in range(steps)]
    q.put((rank, losses, m.g_optimizer.flat.cpu(), m.d_optimizer.flat.cpu(), m._graph is not None, m.graph_error))
    dist.barrier()
    dist.destroy_process_group()
    faulthandler.cancel_dump_traceback_later()


@pytest.mark.parametrize("use_graph", [False, True])
def test_two_gpu_data_parallel_step_matches_single_gpu(golden, use_graph):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    from multi_style_transfer_gan_b200.enhanced_train import EnhancedCycleGAN
    g = golden("train_step_c8_64.pt")
    torch.manual_seed(5)
    A = torch.rand(4, 3, 64, 64) * 2 - 1
    B = torch.rand(4, 3, 64, 64) * 2 - 1
    steps = 3 if use_graph else 1
    # single GPU, the whole batch
    ref = EnhancedCycleGAN(channels=8, num_transformer_blocks=1, precision="fp32", device="cuda:0")
    ref.load_state_dicts(**g["init"])
    g0, d0 = ref.g_optimizer.flat.clone().cpu(), ref.d_optimizer.flat.clone().cpu()
    ref_losses = [ref.train_step(A, B) for _ in range(steps)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q, g["init"], A, B, use_graph, steps)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted([q.get(timeout=200) for _ in range(2)], key=lambda t: t[0])
    for p in ps:
        p.join(120)
        assert p.exitcode == 0
    (_, l0, gf0, df0, graphed0, err0), (_, l1, gf1, df1, graphed1, err1) = res
    if use_graph:
        assert graphed0 and graphed1, (err0, err1)
    # replicas: bit-identical parameters after the steps (same gradients from the all-reduce, same update)
    assert torch.equal(gf0, gf1) and torch.equal(df0, df1)
    # step 1 losses: every loss is a mean over the local half batch, so the average of the two ranks is the global-batch loss
    for k in ref_losses[0]:
        avg = 0.5 * (l0[0][k] + l1[0][k])
        assert abs(avg - ref_losses[0][k]) <= 1e-4 * abs(ref_losses[0][k]) + 1e-6, (k, avg, ref_losses[0][k])
    if steps == 1:
        # the update: Adam's first step moves every element by ~lr * sign(gradient); the data-parallel gradient is the global-batch
        # gradient up to summation order, so only elements whose gradient is ~0 may move the other way
        for now, ref_flat, start, lr in ((gf0, ref.g_optimizer.flat.cpu(), g0, 5e-5), (df0, ref.d_optimizer.flat.cpu(), d0, 2e-4)):
            moved_ref, moved_dp = ref_flat - start, now - start
            live = moved_ref.abs() > 0.5 * lr
            flipped = (torch.sign(moved_ref[live]) != torch.sign(moved_dp[live])).float().mean()
            assert float(flipped) <= 0.02, float(flipped)
            assert float((moved_ref - moved_dp).abs().mean()) <= 0.1 * lr
