"""Time the three parts of the step (forward+losses+decoder backward / encoder backward / SGD) with CUDA events,
each queued behind a device-side spin so that launch overhead is hidden."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "shot-vae_b200"))
import torch
from oracle import shotvae_oracle as O
from shot_vae_model.vae import VariationalAutoEncoder
from shotvae_b200.engine import TrainStep
B, nd = 128, 10
torch.manual_seed(1)
model = VariationalAutoEncoder("wideresnet-28-2", 3, 0, (32, 32), False, 128, nd, 0.67, True).cuda().train()
ts = TrainStep(model, B, hyper=O.default_hyper("Cifar10"), use_graph=False, device_noise=True,
               skip_dead_decoders=os.environ.get("SKIP_DEAD", "0") == "1")
ts.set_epoch(100)
il, ll, iu, lu = O.synthetic_batch(B, nd, 7)
ts.load_inputs(il, ll, iu, lu)
for _ in range(3):
    ts.run_resident()
torch.cuda.synchronize()
acc = [0.0, 0.0, 0.0]
N = 5
for _ in range(N):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    torch.cuda._sleep(80_000_000)
    ev[0].record(); ts._part0(); ev[1].record(); ts._part1(); ev[2].record(); ts._part2(); ev[3].record()
    torch.cuda.synchronize()
    for i in range(3):
        acc[i] += ev[i].elapsed_time(ev[i + 1])
print("part0 (fwd x2, losses, decoder bwd) %.3f ms | part1 (heads+encoder bwd) %.3f ms | part2 (SGD, BN running) %.3f ms | sum %.3f" %
      (acc[0] / N, acc[1] / N, acc[2] / N, sum(acc) / N))
