"""Host-side cProfile of EnhancedCycleGAN.train_step (where does the CPU time per launch go?)."""
import cProfile
import os
import pstats
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from multi_style_transfer_gan_b200.enhanced_train import EnhancedCycleGAN
    torch.manual_seed(0)
    m = EnhancedCycleGAN(channels=64, num_transformer_blocks=3, precision="bf16")
    A = torch.rand(8, 3, 256, 256) * 2 - 1
    B = torch.rand(8, 3, 256, 256) * 2 - 1
    for _ in range(3):
        m.train_step(A, B)
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(3):
        m.train_step(A, B)
    torch.cuda.synchronize()
    pr.disable()
    st = pstats.Stats(pr)
    st.sort_stats("tottime").print_stats(28)


if __name__ == "__main__":
    main()
