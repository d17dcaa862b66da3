import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/shot-vae_b200")
import torch
from shot_vae_model.vae import VariationalAutoEncoder
from shotvae_b200.engine import TrainStep, default_hyper
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
torch.manual_seed(1)
m = VariationalAutoEncoder("wideresnet-28-10", 3, 0, (32, 32), True, 128, 10, 0.67, True).cuda().train()
ts = TrainStep(m, B, hyper=default_hyper("Cifar10"), use_graph=False, device_noise=True)
ts.set_epoch(100)
x = torch.rand(B, 3, 32, 32); y = torch.randint(0, 10, (B,))
for i in range(2):
    t = ts.step(x, y, x, y)
    torch.cuda.synchronize()
    print("step", i, {k: round(v, 3) for k, v in t.items()})
