"""bench.py -- SHOT-VAE WRN-28-2 training-step throughput on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config c2|c3|c4|c5]
  (N > 1: launched by torch.distributed.run, one rank per GPU over NCCL.)

A "step" is one pass of the hot path -- the loop body of main_shot_vae.train (:281-366): 4 network
forwards, 2 backwards, SGD -- over one (labelled 128, unlabelled 128) synthetic batch pair per GPU.
Prints ONE JSON line (rank 0).  `value` is images/s with the inputs already resident in HBM
(CUDA-graph replay of the whole step); `e2e` is the same metric through the public call
`TrainStep.step_async(host tensors)`: every step the batch + the host RNG draws are copied pinned-host -> device
(on a copy stream, beside the previous step), the step runs, and its loss terms are read device -> host (returned
by the NEXT call: the host is never more than one step ahead; `--e2e-sync` times `TrainStep.step`, the same work
with a host synchronisation per step: `e2e.sync_ms_per_step`).  `roofline` describes the dominant
kernel family (the implicit-GEMM convolution kernel), timed with CUDA events around each launch in
an instrumented eager replay of the same launch sequence.  `cpu_baseline` / `--impl reference` time the UNMODIFIED
reference's own `train()` (baseline/_ref, driven by baseline/ref_harness.py in a subprocess: FP32 ATen kernels on this
box's host cores, all of them) on a bounded sample of the same workload; when baseline/_ref is absent they fall back to
the CPU oracle port (kind "port").  `gpu_reference` is the same unmodified reference on ONE GPU through torch's eager
CUDA path (cuDNN TF32) -- the kernel-library baseline the hand-written kernels are measured against.
`python bench.py --kernels` times the bandwidth-bound kernels (ELBO loss+gradient, mixup, SGD, augmentation) at
B = 16 384 (HBM-resident) and at B = 128 (launch floor) and prints their achieved GB/s against the measured HBM peak.
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


# roofline.traffic (DRAM bytes per launch) needs ncu and is therefore not measured inside this run: the bench line carries
# null and names the launch-list summary generated with ncu from the same commit
TRAFFIC_SOURCE = "profiles/r02_launches_n1_summary.md"


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


def run_ref_harness(config, device, steps, warmup, gpu_index="0", timeout=1500):
    """the UNMODIFIED reference's own train() (baseline/_ref) in a subprocess -> the harness's JSON dict, or None"""
    harness = os.path.join(ROOT, "baseline", "ref_harness.py")
    if not os.path.exists(os.path.join(ROOT, "baseline", "_ref", "main_shot_vae.py")):
        return None
    env = dict(os.environ)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT", "CUDA_VISIBLE_DEVICES"):
        env.pop(k, None)
    try:
        r = subprocess.run([sys.executable, harness, "--device", device, "--config", config, "--steps", str(steps), "--warmup", str(warmup),
                            "--batch", str(BATCH), "--epoch", str(EPOCH), "--gpu-index", str(gpu_index)], capture_output=True, text=True,
                           timeout=timeout, env=env)
        lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
        out = json.loads(lines[-1]) if lines else None
        if out is None or "unavailable" in out:
            print("[bench] reference harness: %s" % (r.stderr[-500:] if out is None else out), file=sys.stderr)
            return None
        return out
    except Exception as e:
        print("[bench] reference harness failed: %r" % (e,), file=sys.stderr)
        return None


def cpu_reference_arm(config, cfg, steps, warmup):
    """The reference's CPU implementation of the step on the host cores: its own train() through the harness (kind
    "reference"), or -- only when baseline/_ref is absent -- the CPU oracle port (same ATen FP32 ops)."""
    h = run_ref_harness(config, "cpu", steps, warmup)
    if h is not None:
        return dict(value=h["images_per_s"], unit="images/s", cores=h["cores"], kind="reference", ms_per_step=h["ms_per_step"],
                    sample="%d optimizer steps (batch 128 labelled + 128 unlabelled, FP32) of the unmodified reference's %s on list loaders "
                           "after %d warm-up steps, %.1f s, torch %s CPU" % (steps, h["entry"], warmup, h["seconds"], h["torch"]))
    from oracle import shotvae_oracle as O
    threads = os.cpu_count()
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


def bandwidth_kernels_leg():
    """`--kernels`: the bandwidth-bound kernels of the step alone, CUDA events over back-to-back launches on rotating
    buffers.  B = 16 384 streams far more than the 126 MB L2 (the HBM roofline); B = 128 is the size the step runs them at
    (time against the launch floor)."""
    from shotvae_b200 import _abi
    from shotvae_b200._abi import lib, check, ptr
    pk, how = peaks()
    peak = pk["hbm_gbs"]
    dev = torch.device("cuda")
    st = _abi.stream()
    out = {}

    def timed(name, B, nbytes, make, launch, reps=30, nbuf=4):
        bufs = [make() for _ in range(nbuf)]
        for b in bufs:
            launch(b)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            launch(bufs[i % nbuf])
        e1.record()
        torch.cuda.synchronize()
        us = 1e3 * e0.elapsed_time(e1) / reps
        out.setdefault(name, {})["B=%d" % B] = dict(us=us, algorithmic_mb=nbytes / 1e6, gbps=nbytes / (us * 1e-6) / 1e9,
                                                    frac_of_hbm_peak=nbytes / (us * 1e-6) / 1e9 / peak)

    D, nd, ch, hw = 128, 10, 3, 1024
    for B in (16384, 128):
        n_img = B * ch * hw
        coef = torch.tensor([1e-3] * 16, device=dev)
        # ELBO reconstruction term + gradient: read x, x_hat (fp32), write d x_hat (bf16 NHWC padded to 16 channels, the layout the
        # decoder backward consumes) -- SURVEY 8(d): 3 * B * 3072 * 4 bytes with an fp32 gradient; here the gradient is 16 ch bf16
        def mk_rec():
            return dict(x=torch.rand(B, ch, 32, 32, device=dev), xh=torch.randn(B, 32, 32, ch, device=dev), t=torch.zeros(4, device=dev),
                        g=torch.empty(B, 32, 32, 16, dtype=torch.bfloat16, device=dev))
        timed("sv_elbo_rec_fwd_bwd", B, n_img * 4 * 2 + B * hw * 16 * 2, mk_rec,
              lambda b: check(lib.sv_elbo_rec_fwd_bwd(ptr(b["x"]), ptr(b["xh"]), 1, B, ch, hw, 1, 1.0, ptr(coef), ptr(b["t"]), ptr(b["g"]), 16, None, st)))
        def mk_recf():
            return dict(x=torch.rand(B, ch, 32, 32, device=dev), xh=torch.randn(B, ch, 32, 32, device=dev), t=torch.zeros(4, device=dev),
                        g=torch.empty(B, ch, 32, 32, device=dev))
        timed("sv_elbo_rec_fwd_bwd (fp32 NCHW gradient, the drop-in criterion)", B, n_img * 4 * 3, mk_recf,
              lambda b: check(lib.sv_elbo_rec_fwd_bwd(ptr(b["x"]), ptr(b["xh"]), 0, B, ch, hw, 1, 1.0, None, ptr(b["t"]), None, 0, ptr(b["g"]), st)))
        # mixup / label smoothing: read image + gathered partner, write the mixed image (fp32) + small latents
        def mk_mix():
            return dict(x=torch.rand(B, ch, 32, 32, device=dev), mu=torch.randn(B, D, device=dev), ls=torch.randn(B, D, device=dev) * 0.1,
                        la=torch.log_softmax(torch.randn(B, nd, device=dev), 1), idx=torch.randperm(B, device=dev), lam=torch.tensor([0.3, 0.7], device=dev),
                        o=torch.empty(B, ch, 32, 32, device=dev), omu=torch.empty(B, D, device=dev), osg=torch.empty(B, D, device=dev),
                        oal=torch.empty(B, nd, device=dev))
        timed("sv_mixup_lerp", B, n_img * 4 * 3 + B * (2 * D + nd) * 4 * 3, mk_mix,
              lambda b: check(lib.sv_mixup_lerp(ptr(b["x"]), ptr(b["mu"]), ptr(b["ls"]), ptr(b["la"]), ptr(b["idx"]), ptr(b["lam"]), B, ch, hw, D, nd,
                                                ptr(b["o"]), None, 0, ptr(b["omu"]), ptr(b["osg"]), ptr(b["oal"]), st)))
        # device augmentation: read uint8 images, write fp32 NCHW
        def mk_aug():
            return dict(d=torch.randint(0, 256, (B, 32, 32, 3), dtype=torch.uint8, device=dev), idx=torch.randperm(B, device=dev),
                        p=torch.cat([torch.randint(0, 9, (B, 2), device=dev), torch.randint(0, 2, (B, 1), device=dev)], 1).to(torch.int32),
                        o=torch.empty(B, ch, 32, 32, device=dev))
        timed("sv_augment_batch", B, n_img * (1 + 4), mk_aug,
              lambda b: check(lib.sv_augment_batch(ptr(b["d"]), ptr(b["idx"]), ptr(b["p"]), B, 3, 32, 32, 4, 32, 32, 1, ptr(b["o"]), st)))
    # SGD over the flat arena: read p, g, m; write p, m, g (zeroed)
    for n, tag in ((12790346, "C2 arena (12.79 M parameters)"), (47933770, "C4 arena (47.93 M parameters)")):
        na = (n + 3) // 4 * 4
        hyper = torch.tensor([0.1, 0.9, 5e-4, 1.0, 0.0, 0, 0, 0], device=dev)
        def mk_sgd():
            return dict(p=torch.randn(na, device=dev), g=torch.randn(na, device=dev), m=torch.randn(na, device=dev))
        timed("sv_sgd_step " + tag, n, n * 4 * 6, mk_sgd, lambda b: check(lib.sv_sgd_step(ptr(b["p"]), ptr(b["g"]), ptr(b["m"]), ptr(hyper), n, st)),
              nbuf=3)
    return dict(impl="ours", leg="bandwidth kernels", unit="GB/s", hbm_peak_gbs=peak, peak_source=how, kernels=out,
                note="algorithmic bytes / CUDA-event time over 30 back-to-back launches on rotating buffers; B = 16384 operands exceed the L2, "
                     "B = 128 is the in-step size (latency / launch bound)")


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
    ap.add_argument("--e2e-sync", action="store_true", help="e2e through TrainStep.step (host sync per step) instead of step_async")
    ap.add_argument("--cpu-steps", type=int, default=12)
    ap.add_argument("--dump-kernels", default="")
    ap.add_argument("--kernels", action="store_true", help="time the bandwidth-bound kernels alone (B = 16384 and B = 128)")
    ap.add_argument("--gpu-reference-steps", type=int, default=20, help="steps of the reference's eager CUDA path timed beside ours (0: skip)")
    a = ap.parse_args()
    cfg = CONFIGS[a.config]
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    # `config` is identical in both arms (the driver compares them); everything measured or run-specific goes to `details`
    base = dict(metric=METRIC if a.config in ("c2", "c3") else cfg["workload"].split(",")[0] + " train images/s",
                unit="images/s", n_gpus=world, higher_is_better=True, scaling="weak", vs_baseline=None, data="synthetic",
                config={"workload": cfg["workload"], "schedule_epoch": EPOCH, "batch_per_gpu": "%d labelled + %d unlabelled" % (BATCH, BATCH),
                        "passes_per_step": 2 if cfg["m2"] else 4,
                        "l2": "no explicit flush: one step streams ~1.3 GB of saved activations + 150 MB of parameter/optimizer state, > 126 MB L2"})

    if a.kernels:
        assert torch.cuda.is_available(), "bench.py --kernels needs a CUDA device"
        _emit(bandwidth_kernels_leg())
        return

    if a.impl == "reference":
        if rank != 0:
            return
        # the driver's K and W are honoured; only configurations whose CPU step takes tens of seconds are bounded
        cap = {"c4": 4}.get(a.config, 10 ** 6)
        steps, warm = max(1, min(a.steps, cap)), max(0, min(a.warmup, cap))
        r = cpu_reference_arm(a.config, cfg, steps, warm)
        line = dict(base, impl="reference", value=r["value"], steps=steps, warmup=warm, ms_per_step=r["ms_per_step"], dtype="f32",
                    n_gpus=world, gpu_launches=0,
                    cpu_baseline=dict(value=r["value"], unit="images/s", cores=r["cores"], kind=r["kind"], sample=r["sample"]),
                    e2e=dict(value=r["value"], unit="images/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        _emit(line)
        return

    if a.impl == "gpu_reference":
        # the unmodified reference on ONE GPU through torch's eager CUDA path (its own train(), cuDNN TF32 convolutions)
        if rank != 0:
            return
        h = run_ref_harness(a.config, "cuda", a.steps, max(a.warmup, 3), gpu_index=str(local))
        _emit(dict(base, impl="gpu_reference", **({"unavailable": "baseline/_ref not present"} if h is None else
                   dict(value=h["images_per_s"], steps=a.steps, warmup=max(a.warmup, 3), ms_per_step=h["ms_per_step"], dtype="tf32",
                        details=h))))
        return

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (libshotvae has no CPU path)"
    torch.cuda.set_device(local)
    import torch.distributed as dist
    if world > 1:
        # NCCL's log (communicator ranks, NVLS / ring choice) goes to stderr: fd 1 is guarded, stdout stays the one JSON line
        # (forced: a box-level NCCL_DEBUG=WARN would hide the "comm ... nranks N" lines the driver checks; SHOTVAE_NCCL_DEBUG overrides)
        os.environ["NCCL_DEBUG"] = os.environ.get("SHOTVAE_NCCL_DEBUG", "INFO")
        os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
        # NCCL on at most 4 SMs; the backward's persistent grids then use the other 144 (ddp.GradReducer.cta_limit).  The
        # exchange is 51 MB (C2) / 192 MB (C4) per step hidden behind >= 2 ms of backward: bandwidth is not what it needs.
        # SHOTVAE_NCCL_MAX_CTAS=0 restores NCCL's own choice (24 channels for NVLS all-reduce on 8 GPUs) and full grids.
        if os.environ.get("SHOTVAE_NCCL_MAX_CTAS", "4") != "0":
            os.environ.setdefault("NCCL_MAX_CTAS", os.environ.get("SHOTVAE_NCCL_MAX_CTAS", "4"))
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

    for i in range(3):                          # two eager steps allocate every buffer, the third captures the CUDA graph
        ts.step(*pool[i % len(pool)])           # (a capture failure propagates: there is no eager fallback for the timed run)
        log("warm step %d done" % i)
    assert (ts.graph is not None) == use_graph, "CUDA graph capture did not happen"
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
    # ---- timed region 2: end to end through the public call, host tensors in, loss terms out, every step ----------
    def e2e_region(pipelined):
        sync_all()
        w0 = time.time()
        ev0.record()
        last = None
        for i in range(a.steps):
            if pipelined:
                last = ts.step_async(*pool[i % len(pool)]) or last
            else:
                last = ts.step(*pool[i % len(pool)])
        if pipelined:
            last = ts.drain()            # the last step's terms: inside the timed region
        ev1.record()
        sync_all()
        windows.append((w0, time.time()))
        return ev0.elapsed_time(ev1) / 1e3, last

    t_e2e_sync, last = e2e_region(False)             # TrainStep.step: copy, step, read, host sync -- every step
    t_e2e = t_e2e_sync
    if not a.e2e_sync:
        for i in range(3):                           # builds the staging slots / copy stream outside the timed region
            ts.step_async(*pool[i % len(pool)])
        ts.drain()
        t_e2e, last = e2e_region(True)               # TrainStep.step_async: the same, copies beside the previous step
    log("e2e region done")
    if world > 1:
        tt = torch.tensor([t_res, t_e2e, t_e2e_sync], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_res, t_e2e, t_e2e_sync = float(tt[0]), float(tt[1]), float(tt[2])
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
                    achieved=ach, peak=peak, unit="TFLOP/s", frac=ach / peak, traffic=None,
                    traffic_source="not measured in-run (needs ncu): dram__bytes_read.sum + dram__bytes_write.sum per launch of this family is "
                                   "tabulated in " + TRAFFIC_SOURCE + " (ncu launch list of this command at this commit); algorithmic "
                                   "bytes per launch = hbm.algorithmic_mb_per_step / launches_per_step", peak_source=how + ", sustained bf16",
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
    cpu = gpu_ref = None
    if world == 1 and a.cpu_steps > 0:
        r = cpu_reference_arm(a.config, cfg, a.cpu_steps if a.config != "c4" else min(a.cpu_steps, 2), 1)
        cpu = dict(value=r["value"], unit="images/s", cores=r["cores"], kind=r["kind"], sample=r["sample"])
    if world == 1 and a.gpu_reference_steps > 0:
        # the unmodified reference on this GPU through torch's eager CUDA path, after our own timed regions (separate process:
        # its `lib` / `shot_vae_model` packages share their names with the drop-in ones)
        h = run_ref_harness(a.config, "cuda", a.gpu_reference_steps, 5, gpu_index=str(local))
        if h is not None:
            gpu_ref = dict(value=h["images_per_s"], unit="images/s", ms_per_step=h["ms_per_step"], steps=h["steps"], warmup=h["warmup"],
                           kind="unmodified reference (%s), torch %s eager CUDA, cuDNN %s, %s" % (h["entry"], h["torch"], h.get("cudnn"), h.get("precision")),
                           gpu=h.get("gpu"))
    imgs = world * BATCH * a.steps
    line = dict(base, impl="ours", value=imgs / t_res, steps=a.steps, warmup=W, ms_per_step=1e3 * t_res / a.steps, dtype="bf16",
                clocks=clocks, gpu_launches=launches * a.steps,
                e2e=dict(value=imgs / t_e2e, unit="images/s",
                         h2d_bytes_per_step=ts.h2d_bytes() if a.e2e_sync else ts.h2d_bytes_async(), d2h_bytes_per_step=64,
                         ms_per_step=1e3 * t_e2e / a.steps,
                         api="TrainStep.step (host sync per step)" if a.e2e_sync else
                             "TrainStep.step_async (batch k+1 copied on a copy stream beside step k; terms of step k returned by call k+1; drain() at the end)",
                         sync_ms_per_step=1e3 * t_e2e_sync / a.steps),
                roofline=roof, cpu_baseline=cpu, gpu_reference=gpu_ref)
    line["details"] = dict(parallelism="dp%d" % world, cuda_graph=bool(ts.graph is not None), launches_per_step=launches,
                           noise="device RNG (torch CUDA generator); lambda / pairing drawn on the host as in the reference",
                           last_terms={k: round(v, 4) for k, v in (last or {}).items()},
                           allreduce_bytes_per_step=(reducer.bytes_per_step if reducer else 0))
    _emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
