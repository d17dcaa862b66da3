"""Kernel timeline of CUDA-graph replays of the training step (torch.profiler / CUPTI activity records: kernel name, stream,
start, duration) -> gpurun_out/<tag>_timeline.json + a text summary: busy time per stream, time with no kernel running,
time with exactly one / several kernels running, the longest kernels and the gaps on the critical (main) stream.
usage: python tools/step_timeline.py [config] [tag]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "shot-vae_b200"))
import numpy as np
import torch
from torch.profiler import profile, ProfilerActivity
from shot_vae_model.vae import VariationalAutoEncoder
from shotvae_b200.engine import TrainStep, default_hyper

config = sys.argv[1] if len(sys.argv) > 1 else "c2"
tag = sys.argv[2] if len(sys.argv) > 2 else "tl"
CFG = {"c2": ("wideresnet-28-2", 10, "Cifar10", False), "c3": ("wideresnet-28-2", 100, "Cifar100", False),
       "c4": ("wideresnet-28-10", 10, "Cifar10", False), "c5": ("preactresnet18", 100, "Cifar100", True)}
net, nd, dataset, m2 = CFG[config]
B = 128
torch.manual_seed(1)
model = VariationalAutoEncoder(net, 3, 0, (32, 32), True, 128, nd, 0.67, True).cuda().train()
hyper = default_hyper(dataset, m2)
ts = TrainStep(model, B, hyper=hyper, m2=m2, use_graph=True, device_noise=True)
ts.set_epoch(100)
g = torch.Generator().manual_seed(1234)
batch = (torch.rand(B, 3, 32, 32, generator=g).pin_memory(), torch.randint(0, nd, (B,), generator=g).pin_memory(),
         torch.rand(B, 3, 32, 32, generator=g).pin_memory(), torch.randint(0, nd, (B,), generator=g).pin_memory())
np.random.seed(100)
for _ in range(6):
    ts.step(*batch)
torch.cuda.synchronize()
assert ts.graph is not None
NREP = 4
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(NREP):
        ts.run_resident()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
recs = sorted(((e.time_range.start, e.time_range.end, e.name, getattr(e, "device_resource_id", getattr(e, "thread", 0))) for e in evs
               if "memcpy" not in e.name.lower() and "memset" not in e.name.lower() or True), key=lambda r: r[0])
if not recs:
    print("no CUDA kernel records (CUPTI unavailable?)"); sys.exit(1)
# split into replays: NREP equal chunks by count
per = len(recs) // NREP
rep = recs[per * (NREP - 1):]            # the last replay
t0 = rep[0][0]
rows = [dict(t=round(s - t0, 2), d=round(e - s, 2), name=n.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0][:48], stream=int(st)) for s, e, n, st in rep]
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "%s_timeline.json" % tag), "w"))
span = max(r["t"] + r["d"] for r in rows)
print("replay: %d records, span %.1f us" % (len(rows), span))
streams = sorted(set(r["stream"] for r in rows))
for st in streams:
    rs = [r for r in rows if r["stream"] == st]
    print("stream %d: %d kernels, busy %.1f us, first %.1f last end %.1f" % (st, len(rs), sum(r["d"] for r in rs), rs[0]["t"], max(r["t"] + r["d"] for r in rs)))
# concurrency profile
pts = sorted([(r["t"], 1) for r in rows] + [(r["t"] + r["d"], -1) for r in rows])
lvl, last, acc = 0, 0.0, {}
for t, dlt in pts:
    acc[lvl] = acc.get(lvl, 0.0) + (t - last)
    lvl += dlt; last = t
print("time by number of kernels in flight:", {k: round(v, 1) for k, v in sorted(acc.items())})
# per kernel-name totals
tot = {}
for r in rows:
    k = r["name"]
    a = tot.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += r["d"]
print("top kernels by total time:")
for k, (n, d) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:25]:
    print("  %-60s x%3d %8.1f us  avg %6.1f" % (k, n, d, d / n))
# coarse timeline: 100-us buckets, busy time per stream
print("timeline (per 100 us: busy us per stream %s)" % streams)
nb = int(span // 100) + 1
for b in range(nb):
    lo, hi = b * 100.0, (b + 1) * 100.0
    line = []
    for st in streams:
        busy = sum(max(0.0, min(hi, r["t"] + r["d"]) - max(lo, r["t"])) for r in rows if r["stream"] == st)
        line.append("%5.0f" % busy)
    print("  %5d: %s" % (lo, " ".join(line)))

# kernel sequence of the busiest stream in the backward window (the critical chain)
if os.environ.get("TL_DUMP", "1") == "1":
    for st in streams:
        rs = [r for r in rows if r["stream"] == st]
        print("---- stream %d" % st)
        prev_end = None
        for r in rs:
            gap = 0.0 if prev_end is None else r["t"] - prev_end
            print("  t=%7.1f gap=%6.1f d=%6.1f %s" % (r["t"], gap, r["d"], r["name"]))
            prev_end = r["t"] + r["d"]
