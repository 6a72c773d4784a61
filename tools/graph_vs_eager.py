import sys, os, torch
sys.path.insert(0, os.getcwd())
from multi_style_transfer_gan_b200.enhanced_train import EnhancedCycleGAN
g = torch.load("tests/golden/train_step_c8_64.pt", weights_only=False)
runs = []
for use_graph in (False, False, True, True):
    m = EnhancedCycleGAN(channels=8, num_transformer_blocks=1, precision="fp32", use_graph=use_graph, graph_warmup=1)
    m.load_state_dicts(**g["init"])
    runs.append([m.train_step(g["real_A"], g["real_B"]) for _ in range(5)])
for name, (i, j) in (("eager vs eager", (0, 1)), ("graph vs eager", (2, 0)), ("graph vs graph", (2, 3))):
    for step in range(5):
        print(name, step, " ".join(f"{k}:{abs(runs[i][step][k]-runs[j][step][k])/abs(runs[j][step][k]):.2e}" for k in runs[0][0]))
