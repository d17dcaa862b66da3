"""bench.py -- SHOT-VAE WRN-28-2 training-step throughput on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2|c3|c4|c5]
  (N > 1: launched by torch.distributed.run, one rank per GPU over NCCL.)

A "step" is one pass of the hot path -- the loop body of main_shot_vae.train (:281-366): 4 network
forwards, 2 backwards, SGD -- over one (labelled 128, unlabelled 128) synthetic batch pair per GPU.
Prints ONE JSON line (rank 0).  `value` is images/s with the inputs already resident in HBM
(CUDA-graph replay of the whole step); `e2e` is the same metric through the public call
`TrainStep.step(host tensors)`: pinned-host -> device copies of the batch + the host RNG draws, the
step, and a device -> host read of the loss terms, every step.  `roofline` describes the dominant
kernel family (the implicit-GEMM convolution kernel), timed with CUDA events around each launch in
an instrumented eager replay of the same launch sequence.  `cpu_baseline` is the CPU oracle port
(oracle/shotvae_oracle.py, torch FP32 ATen kernels = what the reference executes on a CPU) timed
on this box's host cores on a bounded sample.  `--impl reference` times that CPU path alone.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "shot-vae_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np
import torch

METRIC = "SHOT-VAE WRN-28-2 train images/s"
CONFIGS = {
    "c2": dict(net="wideresnet-28-2", nd=10, dataset="Cifar10", m2=False, br=True,
               workload="C2: SHOT-VAE WideResNet-28-2, Cifar10-shaped synthetic 3x32x32, nd=10, --br, batch 128 labelled + 128 unlabelled per GPU"),
    "c3": dict(net="wideresnet-28-2", nd=100, dataset="Cifar100", m2=False, br=True,
               workload="C3: SHOT-VAE WideResNet-28-2, Cifar100-shaped synthetic, nd=100, batch 128+128 per GPU"),
    "c4": dict(net="wideresnet-28-10", nd=10, dataset="Cifar10", m2=False, br=True,
               workload="C4: SHOT-VAE WideResNet-28-10, Cifar10-shaped synthetic, nd=10, batch 128+128 per GPU"),
    "c5": dict(net="preactresnet18", nd=100, dataset="Cifar100", m2=True, br=False,
               workload="C5: M2-VAE PreActResNet18, Cifar100-shaped synthetic, nd=100, batch 128+128 per GPU"),
}
EPOCH = 100        # schedules evaluated at a fixed epoch where every loss term is active
BATCH = 128


# DRAM bytes per launch of the sv_igemm_fprop family (C2, N=1), from the ncu pass summarised in
# profiles/r01_launches_n1_summary.md: 14.02 MB against 19.7 MB algorithmic (outputs and re-read inputs stay in L2)
NCU_TRAFFIC_BYTES_PER_LAUNCH = 14.02e6


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed regions run"""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def summary(self, windows):
        if self.proc is not None:
            self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for t, line in self.lines:
            if not any(a <= t <= b for a, b in windows):
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); mx = max(mx, float(f[1]))
            except Exception:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_reference_arm(cfg, steps, warmup, threads=None):
    """the reference's CPU path (oracle port: same ATen FP32 ops the reference issues) on the host cores"""
    from oracle import shotvae_oracle as O
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    hyper = O.default_hyper(cfg["dataset"], cfg["m2"])
    hyper["br"] = cfg["br"]
    st = O.init_state(cfg["net"], cfg["nd"])
    il, ll, iu, lu = O.synthetic_batch(BATCH, cfg["nd"], 0)
    fn = O.m2_step if cfg["m2"] else O.shot_step
    mom = {}
    torch.manual_seed(0); np.random.seed(0)
    ts = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        fn(st, cfg["net"], cfg["nd"], il, ll, iu, lu, EPOCH, hyper, O.LiveDraws())
        O.sgd_step(st, mom, hyper["lr"], hyper["momentum"], hyper["wd"])
        if i >= warmup:
            ts.append(time.perf_counter() - t0)
    tot = float(sum(ts))
    return dict(value=BATCH * steps / tot, unit="images/s", cores=threads, kind="port",
                sample="%d full training steps (batch 128+128, FP32) of the CPU oracle after %d warm-up, %.1f s" % (steps, warmup, tot),
                ms_per_step=1e3 * tot / steps)


_REAL_STDOUT = None


def _guard_stdout():
    """stdout carries exactly ONE JSON line: everything else that writes to fd 1 (NCCL's version banner, library
    chatter from C code) is sent to stderr for the lifetime of the process"""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def _emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _guard_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", default="c2")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--cpu-steps", type=int, default=12)
    ap.add_argument("--dump-kernels", default="")
    a = ap.parse_args()
    cfg = CONFIGS[a.config]
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    base = dict(metric=METRIC if a.config in ("c2", "c3") else cfg["workload"].split(",")[0] + " train images/s",
                unit="images/s", n_gpus=world, higher_is_better=True, scaling="weak", vs_baseline=None, data="synthetic",
                config={"workload": cfg["workload"], "schedule_epoch": EPOCH})

    if a.impl == "reference":
        if rank != 0:
            return
        steps, warm = max(1, min(a.steps, 20)), max(1, min(a.warmup, 2))
        r = cpu_reference_arm(cfg, steps, warm)
        line = dict(base, impl="reference", value=r["value"], steps=steps, warmup=warm, ms_per_step=r["ms_per_step"], dtype="f32",
                    n_gpus=0 if world == 1 else world, gpu_launches=0,
                    cpu_baseline=dict(value=r["value"], unit="images/s", cores=r["cores"], kind=r["kind"], sample=r["sample"]),
                    e2e=dict(value=r["value"], unit="images/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        line["n_gpus"] = world
        _emit(line)
        return

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (libshotvae has no CPU path)"
    torch.cuda.set_device(local)
    import torch.distributed as dist
    if world > 1:
        os.environ["NCCL_DEBUG"] = os.environ.get("SHOTVAE_NCCL_DEBUG", "WARN")   # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from shot_vae_model.vae import VariationalAutoEncoder
    from shotvae_b200.engine import TrainStep, default_hyper
    from shotvae_b200 import _abi          # noqa: F401  (loads libshotvae.so: fails loudly here if the extension is missing)
    from shotvae_b200.ddp import GradReducer

    log = lambda m: print("[bench rank %d] %s" % (rank, m), file=sys.stderr, flush=True)
    log("process group up" if world > 1 else "single GPU")
    torch.manual_seed(1)
    model = VariationalAutoEncoder(cfg["net"], 3, 0, (32, 32), True, 128, cfg["nd"], 0.67, True).cuda().train()
    hyper = default_hyper(cfg["dataset"], cfg["m2"])
    hyper["br"] = cfg["br"]
    model._ensure_bound()
    reducer = GradReducer(model._net) if world > 1 else None
    use_graph = not a.no_graph
    ts = TrainStep(model, BATCH, hyper=hyper, m2=cfg["m2"], use_graph=use_graph, device_noise=True, reducer=reducer)
    ts.set_epoch(EPOCH)
    # synthetic host batches (pinned), a small rotating pool; per-rank seed
    g = torch.Generator().manual_seed(1234 + rank)
    pool = [(torch.rand(BATCH, 3, 32, 32, generator=g).pin_memory(), torch.randint(0, cfg["nd"], (BATCH,), generator=g).pin_memory(),
             torch.rand(BATCH, 3, 32, 32, generator=g).pin_memory(), torch.randint(0, cfg["nd"], (BATCH,), generator=g).pin_memory())
            for _ in range(4)]
    np.random.seed(100 + rank)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    graph_ok = use_graph
    try:
        for i in range(3):                      # allocate buffers, then capture the CUDA graph
            ts.step(*pool[i % len(pool)])
            log("warm step %d done" % i)
    except Exception as e:                      # capture refused (e.g. a collective that cannot be captured)
        if not use_graph:
            raise
        torch.cuda.synchronize()
        graph_ok = False
        ts.use_graph, ts.graph = False, None
        base["config"]["graph_fallback"] = repr(e)[:200]
        ts.step(*pool[0])
    W = max(a.warmup, 3)
    for i in range(W):
        ts.step(*pool[i % len(pool)])
    log("warm-up done")
    sampler = ClockSampler(local) if rank == 0 else None
    windows = []
    # ---- timed region 1: inputs resident in HBM --------------------------------------------------
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    w0 = time.time()
    ev0.record()
    for _ in range(a.steps):
        ts.run_resident()
    ev1.record()
    sync_all()
    windows.append((w0, time.time()))
    t_res = ev0.elapsed_time(ev1) / 1e3
    log("resident region done")
    # ---- timed region 2: end to end through TrainStep.step(host tensors) -----------------------------
    sync_all()
    w0 = time.time()
    ev0.record()
    last = None
    for i in range(a.steps):
        last = ts.step(*pool[i % len(pool)])
    ev1.record()
    sync_all()
    windows.append((w0, time.time()))
    t_e2e = ev0.elapsed_time(ev1) / 1e3
    log("e2e region done")
    if world > 1:
        tt = torch.tensor([t_res, t_e2e], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_res, t_e2e = float(tt[0]), float(tt[1])
    clocks = sampler.summary(windows) if sampler else None
    launches = (ts.launches_per_step or 0)

    # ---- roofline of the dominant kernel family: instrumented eager replay ---------------------------
    roof = None
    if rank == 0:
        net = model._net
        saved = (ts.use_graph, ts.graph)
        saved_reducer, ts.reducer = ts.reducer, None      # rank-0-only replay: no collective (the other ranks are not in it)
        saved_side, net.side = net.side, None             # one stream: every launch is timed alone, not next to another stream's
        ts.use_graph = False
        ts.run_resident()
        net.timing = []
        for _ in range(3):
            # queue the step behind a ~25 ms device-side spin so that every launch (and its bracketing events)
            # is already enqueued when it runs: the event deltas are then GPU durations, not CPU launch gaps
            torch.cuda._sleep(50_000_000)
            ts.run_resident()
            torch.cuda.synchronize()
        recs = net.timing
        net.timing = None
        ts.use_graph, ts.graph = saved
        ts.reducer = saved_reducer
        net.side = saved_side
        agg = {}
        for kind, key, flops, e0, e1, nbytes in recs:
            d = agg.setdefault(kind, dict(ms=0.0, flops=0.0, n=0, bytes=0.0))
            d["ms"] += e0.elapsed_time(e1); d["flops"] += flops; d["n"] += 1; d["bytes"] += nbytes
        pk, how = peaks()
        f = agg.get("igemm_fprop", dict(ms=1e-9, flops=0.0, n=1, bytes=0.0))
        hbm_peak = pk.get("hbm_gbs")
        hbm_ach = f["bytes"] / (f["ms"] * 1e-3) / 1e9
        ach = f["flops"] / (f["ms"] * 1e-3) / 1e12
        peak = pk.get("bf16_tflops_sustained", pk.get("bf16_tflops"))
        step_flops = sum(d["flops"] for d in agg.values()) / 3
        roof = dict(bound="tensor", kernel="sv_igemm_fprop family (conv / convT fprop + dgrad: tcgen05 halo-tile and per-tap TMA kernels, "
                                           "mma.sync for strided / C=16 shapes)",
                    achieved=ach, peak=peak, unit="TFLOP/s", frac=ach / peak, traffic=NCU_TRAFFIC_BYTES_PER_LAUNCH,
                    traffic_source="dram__bytes_read.sum + dram__bytes_write.sum averaged over the family's launches of one step, "
                                   "profiles/r01_launches_n1_summary.md (ncu, C2, N=1); algorithmic bytes per launch = "
                                   "hbm.algorithmic_mb_per_step / launches_per_step", peak_source=how + ", sustained bf16",
                    launches_per_step=f["n"] // 3, avg_launch_us=1e3 * f["ms"] / max(f["n"], 1),
                    algorithmic_gflop_per_step=step_flops / 1e9,
                    hbm=dict(achieved=hbm_ach, peak=hbm_peak, unit="GB/s", frac=(hbm_ach / hbm_peak if hbm_peak else None),
                             algorithmic_mb_per_step=f["bytes"] / 3 / 1e6,
                             note="same launches against the HBM roof: the 32-channel (32x32) layers sit below the ridge "
                                  "(~144 FLOP/B vs ~216), the 64/128-channel layers above it"),
                    step_share={k: dict(ms_per_step=d["ms"] / 3, gbps=d["bytes"] / (d["ms"] * 1e-3) / 1e9 if d["ms"] > 0 else 0.0, tflops=(d["flops"] / (d["ms"] * 1e-3) / 1e12 if d["ms"] > 0 else 0.0),
                                        launches=d["n"] // 3) for k, d in agg.items()},
                    eager_instrumented_ms_per_step=None,
                    measured="CUDA events around every launch of the kernel in an instrumented eager replay of the step "
                             "(3 steps, each queued behind a device-side spin so launches run back to back)")
        if a.dump_kernels:
            per = {}
            for kind, key, flops, e0, e1, nbytes in recs:
                d = per.setdefault(kind + ":" + key, dict(ms=0.0, flops=flops, bytes=nbytes, n=0))
                d["ms"] += e0.elapsed_time(e1); d["n"] += 1
            for d in per.values():
                d["us"] = 1e3 * d["ms"] / d["n"]; d["tflops"] = d["flops"] / (d["us"] * 1e-6) / 1e12 if d["us"] > 0 else 0
                d["gbps"] = d["bytes"] / (d["us"] * 1e-6) / 1e9 if d["us"] > 0 else 0
            os.makedirs(os.path.dirname(os.path.abspath(a.dump_kernels)), exist_ok=True)
            json.dump(per, open(a.dump_kernels, "w"), indent=1, sort_keys=True)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None
    if world == 1 and a.cpu_steps > 0:
        r = cpu_reference_arm(cfg, a.cpu_steps, 1)
        cpu = dict(value=r["value"], unit="images/s", cores=r["cores"], kind=r["kind"], sample=r["sample"])
    imgs = world * BATCH * a.steps
    line = dict(base, impl="ours", value=imgs / t_res, steps=a.steps, warmup=W, ms_per_step=1e3 * t_res / a.steps, dtype="bf16",
                clocks=clocks, gpu_launches=launches * a.steps,
                e2e=dict(value=imgs / t_e2e, unit="images/s", h2d_bytes_per_step=ts.h2d_bytes(), d2h_bytes_per_step=64,
                         ms_per_step=1e3 * t_e2e / a.steps),
                roofline=roof, cpu_baseline=cpu)
    line["config"].update(parallelism="dp%d" % world, passes_per_step=2 if cfg["m2"] else 4, cuda_graph=bool(graph_ok and ts.graph is not None),
                          launches_per_step=launches, noise="device RNG (torch CUDA generator); lambda / pairing drawn on the host as in the reference",
                          l2="no explicit flush: one step streams ~1.3 GB of saved activations + 150 MB of parameter/optimizer state, > 126 MB L2",
                          last_terms={k: round(v, 4) for k, v in (last or {}).items()},
                          allreduce_bytes_per_step=(reducer.bytes_per_step if reducer else 0))
    _emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
