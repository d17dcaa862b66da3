"""Micro-benchmark of the implicit-GEMM kernels (run on the GPU box): per shape and per kernel
implementation, time per launch (CUDA events, L2-cold rotation over several buffer sets) and the
max abs difference against the mma.sync kernel."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "shot-vae_b200"))
import torch
from shotvae_b200 import _abi
from shotvae_b200._abi import lib, check, ptr, taps_array, IgemmArgs
from shotvae_b200.plan import conv_taps

SHAPES = [  # name, NB, H, C, N, k
    ("block1 3x3 32->32 @32x32", 256, 32, 32, 32, 3),
    ("block2 3x3 64->64 @16x16", 256, 16, 64, 64, 3),
    ("block3 3x3 128->128 @8x8", 256, 8, 128, 128, 3),
    ("u0.conv1 3x3 16->32 @32x32", 256, 32, 16, 32, 3),
    ("sc 1x1 16->32 @32x32", 256, 32, 16, 32, 1),
]


def run(shape, impl, with_epi, nset=6, iters=30):
    name, NB, H, Cc, N, k = shape
    taps = conv_taps(k, k // 2)
    torch.manual_seed(0)
    W = (torch.randn(len(taps), N, Cc, device="cuda") * 0.1).to(torch.bfloat16)
    if impl == 3:    # 8-channel-plane layout [T][C/8][N][8]
        W = W.view(len(taps), N, Cc // 8, 8).permute(0, 2, 1, 3).contiguous()
    As = [torch.randn(NB, H, H, Cc, device="cuda").to(torch.bfloat16) for _ in range(nset)]
    Rs = [torch.randn(NB, H, H, N, device="cuda").to(torch.bfloat16) for _ in range(nset)]
    Os = [torch.empty(NB, H, H, N, device="cuda", dtype=torch.bfloat16) for _ in range(nset)]
    stats = torch.zeros(2, 2, N, device="cuda")
    a = IgemmArgs()
    a.Wt = ptr(W)
    a.NB, a.H, a.W, a.C, a.OH, a.OW, a.N, a.T = NB, H, H, Cc, H, H, N, len(taps)
    a.in_stride, a.out_stride, a.out_off_y, a.out_off_x, a.OHf, a.OWf = 1, 1, 0, 0, H, H
    a.n_valid, a.group_images = 0, NB // 2
    a.dy, a.dx = taps_array([t[1] for t in taps]), taps_array([t[2] for t in taps])
    a.impl = impl
    a.w_layout = 1 if impl == 3 else 0
    a.A, a.out_bf16 = ptr(As[0]), ptr(Os[0])
    if with_epi:
        a.residual, a.stats = ptr(Rs[0]), ptr(stats)
    if not lib.sv_igemm_fprop_supports(C.byref(a), impl):
        return None, None

    def launch(i):
        st = _abi.stream()
        a.A, a.out_bf16 = ptr(As[i % nset]), ptr(Os[i % nset])
        if with_epi:
            a.residual = ptr(Rs[i % nset])
        check(lib.sv_igemm_fprop(C.byref(a), st))
    for i in range(5):
        launch(i)
    torch.cuda.synchronize()
    # GPU time, not Python launch rate: capture `iters` launches into a CUDA graph and time the replay
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters):
            launch(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    stats.zero_()
    launch(0)
    torch.cuda.synchronize()
    return us, (Os[0].float().clone(), stats.clone())


def main():
    out = {}
    for shape in SHAPES:
        for epi in (False, True):
            ref = None
            for impl in (1, 2, 3):
                try:
                    us, res = run(shape, impl, epi)
                except Exception as e:
                    print("%-30s epi=%d impl=%d ERROR %s" % (shape[0], epi, impl, str(e)[:150]))
                    continue
                if us is None:
                    continue
                if impl == 1:
                    ref = res
                err = float((res[0] - ref[0]).abs().max()) if ref is not None else -1
                serr = float((res[1] - ref[1]).abs().max() / ref[1].abs().max().clamp_min(1e-9)) if (ref is not None and epi) else 0.0
                name, NB, H, Cc, N, k = shape
                gflop = 2.0 * NB * H * H * N * Cc * k * k / 1e9
                mb = NB * H * H * (Cc + N * (2 if epi else 1)) * 2 / 1e6
                print("%-30s epi=%d impl=%d %8.1f us  %7.1f TFLOP/s  %6.0f GB/s(unique)  maxdiff=%.3g statsdiff=%.2g" %
                      (name, epi, impl, us, gflop / us * 1e-3, mb / us * 1e-3 * 1e3, err, serr))
                out["%s|epi%d|impl%d" % (name, epi, impl)] = dict(us=us, tflops=gflop / us * 1e-3, maxdiff=err, statsdiff=serr)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "igemm_bench.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
