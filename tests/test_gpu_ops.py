"""GPU parity tests, kernel level: every libshotvae entry point against the CPU oracle / plain torch
FP32 on the same seeded inputs.  Tolerances: index and integer results bit-exact; FP32 kernels 1e-5
relative; bf16-operand GEMMs are compared on bf16-rounded inputs with FP32 accumulation, so only
the accumulation order differs (tolerance 2e-3 of the output RMS, bf16 output rounding included)."""
import ctypes as C
import glob
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sv():
    from shotvae_b200 import _abi
    return _abi


def dev(t):
    return t.cuda()


def rel_rms(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / max(b.norm(), 1e-30))


def bf(x):
    return x.to(torch.bfloat16).float()


def nhwc(x):           # NCHW fp32 -> NHWC bf16 cuda
    return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda()


def from_nhwc(x):      # NHWC cuda -> NCHW fp32 cpu
    return x.float().cpu().permute(0, 3, 1, 2).contiguous()


def pack(sv, w, N, Cc, taps, n_real, c_real, sn, sc, st, impl=1):
    """packed bf16 weights; the halo-tile kernel (impl 3) takes the 8-channel-plane layout"""
    from shotvae_b200._abi import lib, check, ptr, taps_array
    dst = torch.zeros(len(taps), N, Cc, dtype=torch.bfloat16, device="cuda")
    wd = w.contiguous().cuda()
    check(lib.sv_pack_weight(ptr(wd), ptr(dst), N, Cc, len(taps), n_real, c_real, sn, sc, st, taps_array([t[0] for t in taps]),
                             1 if impl == 3 else 0, sv.stream()))
    return dst


def igemm(sv, A, Wt, taps, NB, H, W, Cc, OH, OW, N, in_stride=1, out=None, outf=None, res=None, bias=None, stats=None,
          out_stride=1, off=(0, 0), OHf=None, OWf=None, n_valid=0, group_images=None, impl=1, batch=None):
    from shotvae_b200._abi import lib, check, ptr, taps_array, IgemmArgs
    a = IgemmArgs()
    a.A, a.Wt, a.out_bf16, a.out_f32, a.residual, a.bias, a.stats = ptr(A), ptr(Wt), ptr(out), ptr(outf), ptr(res), ptr(bias), ptr(stats)
    a.NB, a.H, a.W, a.C, a.OH, a.OW, a.N, a.T = NB, H, W, Cc, OH, OW, N, len(taps)
    a.in_stride, a.out_stride, a.out_off_y, a.out_off_x = in_stride, out_stride, off[0], off[1]
    a.OHf, a.OWf = OHf or OH * out_stride, OWf or OW * out_stride
    a.n_valid, a.group_images = n_valid, group_images or NB
    a.dy, a.dx = taps_array([t[1] for t in taps]), taps_array([t[2] for t in taps])
    a.impl = impl
    a.w_layout = 1 if impl == 3 else 0
    if impl > 1 and not lib.sv_igemm_fprop_supports(C.byref(a), impl):
        pytest.skip("shape not covered by tcgen05 kernel %d (runs on another kernel)" % impl)
    if batch is not None:
        batch.append((a, A, Wt))           # (operands are kept alive with the argument block)
        return
    check(lib.sv_igemm_fprop(C.byref(a), sv.stream()))


IMPLS = [1, 2, 3]   # 1 = mma.sync, 2 = tcgen05 + per-tap TMA, 3 = tcgen05 halo-tile kernel


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("cin,cout,H,stride,k,NB", [
    (32, 32, 32, 1, 3, 4), (16, 32, 32, 1, 3, 3), (32, 64, 32, 2, 3, 4), (64, 64, 16, 1, 3, 8), (64, 128, 16, 2, 3, 8),
    (128, 128, 8, 1, 3, 16), (16, 32, 32, 1, 1, 2), (32, 64, 32, 2, 1, 4), (160, 160, 8, 1, 3, 2), (16, 16, 32, 1, 3, 5),
    (160, 160, 32, 1, 3, 2), (320, 320, 16, 1, 3, 4), (640, 640, 8, 1, 3, 4), (16, 160, 32, 1, 3, 2), (160, 320, 16, 1, 1, 4),
    # stride 2 on the TMA kernel (tensor-map element strides): first conv / projection shortcut of a resolution block
    (160, 320, 32, 2, 3, 2), (320, 640, 16, 2, 3, 8), (160, 320, 32, 2, 1, 4), (64, 128, 32, 2, 3, 4), (256, 512, 8, 2, 3, 32)])
def test_conv_fprop_matches_torch(sv, impl, cin, cout, H, stride, k, NB):
    from shotvae_b200.plan import conv_taps
    torch.manual_seed(cin * 1000 + cout + H + stride)
    x, w = bf(torch.randn(NB, cin, H, H)), bf(torch.randn(cout, cin, k, k) * 0.1)
    bias, resid = torch.randn(cout), bf(torch.randn(NB, cout, H // stride, H // stride))
    want = F.conv2d(x, w, bias, stride, k // 2) + resid
    taps = conv_taps(k, k // 2)
    Wt = pack(sv, w, cout, cin, taps, cout, cin, cin * k * k, k * k, 1, impl)
    Ho = H // stride
    out = torch.empty(NB, Ho, Ho, cout, dtype=torch.bfloat16, device="cuda")
    G = 1 if NB % 2 else 2
    stats = torch.zeros(G, 2, cout, device="cuda")
    igemm(sv, nhwc(x), Wt, taps, NB, H, H, cin, Ho, Ho, cout, in_stride=stride, out=out, res=nhwc(resid), bias=bias.cuda(),
          stats=stats, group_images=NB // G, impl=impl)
    got = from_nhwc(out)
    assert rel_rms(got, want) < 4e-3
    gq = got.view(G, NB // G, cout, -1)
    assert rel_rms(stats[:, 0].cpu(), gq.sum(dim=(1, 3))) < 2e-3
    assert rel_rms(stats[:, 1].cpu(), (gq * gq).sum(dim=(1, 3))) < 1e-3


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("cin,cout,Hin,NB", [(1024, 512, 1, 8), (512, 256, 2, 4), (128, 64, 8, 4), (64, 3, 16, 4)])
def test_convT_fprop_phases_match_torch(sv, impl, cin, cout, Hin, NB):
    from shotvae_b200.plan import dgrad_phase_taps, live_taps, pad16
    torch.manual_seed(cin + cout)
    x, w = bf(torch.randn(NB, cin, Hin, Hin)), bf(torch.randn(cin, cout, 4, 4) * 0.05)
    want = F.conv_transpose2d(x, w, None, 2, 1)
    cp, Ho = pad16(cout), 2 * Hin
    out = torch.zeros(NB, Ho, Ho, cp, dtype=torch.bfloat16, device="cuda")
    outf = torch.zeros(NB, Ho, Ho, cout, dtype=torch.float32, device="cuda")
    A = nhwc(x)
    for (py, px), taps in dgrad_phase_taps(4, 2, 1).items():
        taps = live_taps(taps, Hin, Hin, Hin, Hin, 1)
        Wt = pack(sv, w, cp, cin, taps, cout, cin, 16, cout * 16, 1, impl)
        igemm(sv, A, Wt, taps, NB, Hin, Hin, cin, Hin, Hin, cp, out=out, outf=outf, out_stride=2, off=(py, px), OHf=Ho, OWf=Ho,
              n_valid=cout, impl=impl)
    assert rel_rms(from_nhwc(out)[:, :cout], want) < 4e-3
    assert rel_rms(from_nhwc(outf), want) < 1e-3


# ---- the benchmark regime: NB = 256 images per forward launch ([P1|P3] / [P2|P4], two pass groups) and NB = 512 per
# backward launch (four pass groups), every stride-1 conv shape of WRN-28-2.  Each persistent CTA then runs
# 8 ... 31 tiles, so the TMEM accumulator ring (8 deep), the shared-memory stage ring and the cross-tile BatchNorm
# statistics registers all wrap -- teacher-forced against torch FP32 on the same bf16-rounded operands.
WRN282_SHAPES = [(16, 32, 32, 3), (32, 32, 32, 3), (16, 32, 32, 1), (64, 64, 16, 3), (128, 128, 8, 3)]


@pytest.mark.parametrize("NB,G", [(256, 2), (512, 4)])
@pytest.mark.parametrize("cin,cout,H,k", WRN282_SHAPES)
def test_conv_fprop_benchmark_batch_matches_torch(sv, cin, cout, H, k, NB, G):
    from shotvae_b200.plan import conv_taps
    from shotvae_b200._abi import lib, IgemmArgs
    torch.manual_seed(cin * 100 + cout + H + NB)
    x, w = bf(torch.randn(NB, cin, H, H)), bf(torch.randn(cout, cin, k, k) * 0.1)
    resid = bf(torch.randn(NB, cout, H, H))
    torch.set_num_threads(max(1, (os.cpu_count() or 8)))
    want = F.conv2d(x, w, None, 1, k // 2) + resid
    taps = conv_taps(k, k // 2)
    # the kernel the engine picks for this shape (halo-tile layout when it covers it, else auto selection)
    probe = IgemmArgs()
    probe.A = probe.Wt = probe.out_bf16 = 4096          # (alignment checks only; never dereferenced)
    probe.NB, probe.H, probe.W, probe.C, probe.OH, probe.OW, probe.N, probe.T = NB, H, H, cin, H, H, cout, len(taps)
    probe.in_stride, probe.out_stride, probe.OHf, probe.OWf, probe.group_images, probe.w_layout = 1, 1, H, H, NB // G, 1
    from shotvae_b200._abi import taps_array
    probe.dy, probe.dx = taps_array([t[1] for t in taps]), taps_array([t[2] for t in taps])
    impl = 3 if lib.sv_igemm_fprop_supports(C.byref(probe), 3) and not (cin >= 128 and cout >= 128) else 0
    Wt = pack(sv, w, cout, cin, taps, cout, cin, cin * k * k, k * k, 1, impl)
    out = torch.empty(NB, H, H, cout, dtype=torch.bfloat16, device="cuda")
    stats = torch.zeros(G, 2, cout, device="cuda")
    igemm(sv, nhwc(x), Wt, taps, NB, H, H, cin, H, H, cout, out=out, res=nhwc(resid), stats=stats, group_images=NB // G, impl=impl)
    got = from_nhwc(out)
    assert rel_rms(got, want) < 4e-3
    gq = got.view(G, NB // G, cout, -1)
    assert rel_rms(stats[:, 0].cpu(), gq.sum(dim=(1, 3))) < 2e-3
    assert rel_rms(stats[:, 1].cpu(), (gq * gq).sum(dim=(1, 3))) < 1e-3
    # second launch into the same buffers: the rings must come back to the same state (bit-identical output)
    out2 = torch.empty_like(out)
    igemm(sv, nhwc(x), Wt, taps, NB, H, H, cin, H, H, cout, out=out2, res=nhwc(resid), impl=impl)
    assert torch.equal(out, out2)


# wide layers on the TMA kernel at sizes that fill the machine (WRN-28-10): the widest channel tile (160 / 256 columns) and, with
# >= 2 x 74 tile pairs, the 2-CTA cluster variant whose weight tiles are fetched half by each CTA and multicast to both
@pytest.mark.parametrize("cin,cout,H,k,NB,G", [(320, 320, 16, 3, 80, 2), (640, 640, 8, 3, 160, 4), (160, 160, 32, 3, 20, 2), (256, 256, 8, 3, 304, 1),
                                               (160, 320, 16, 1, 80, 1),
                                               (640, 640, 8, 3, 150, 1), (320, 160, 16, 3, 160, 4)])     # (odd tile count: phantom tile of a CTA pair)
def test_conv_fprop_wide_tiles_match_torch(sv, cin, cout, H, k, NB, G):
    from shotvae_b200.plan import conv_taps
    torch.manual_seed(cin + cout + H + NB)
    x, w = bf(torch.randn(NB, cin, H, H)), bf(torch.randn(cout, cin, k, k) * 0.05)
    resid = bf(torch.randn(NB, cout, H, H))
    torch.set_num_threads(max(1, (os.cpu_count() or 8)))
    want = F.conv2d(x, w, None, 1, k // 2) + resid
    taps = conv_taps(k, k // 2)
    Wt = pack(sv, w, cout, cin, taps, cout, cin, cin * k * k, k * k, 1, 2)
    out = torch.empty(NB, H, H, cout, dtype=torch.bfloat16, device="cuda")
    stats = torch.zeros(G, 2, cout, device="cuda")
    igemm(sv, nhwc(x), Wt, taps, NB, H, H, cin, H, H, cout, out=out, res=nhwc(resid), stats=stats, group_images=NB // G, impl=2)
    got = from_nhwc(out)
    assert rel_rms(got, want) < 4e-3
    gq = got.view(G, NB // G, cout, -1)
    assert rel_rms(stats[:, 0].cpu(), gq.sum(dim=(1, 3))) < 2e-3
    assert rel_rms(stats[:, 1].cpu(), (gq * gq).sum(dim=(1, 3))) < 1e-3
    out2 = torch.empty_like(out)
    igemm(sv, nhwc(x), Wt, taps, NB, H, H, cin, H, H, cout, out=out2, res=nhwc(resid), impl=2)
    assert torch.equal(out, out2)


@pytest.mark.parametrize("cin,cout,H,k", WRN282_SHAPES)
def test_conv_wgrad_benchmark_batch_matches_autograd(sv, cin, cout, H, k):
    from shotvae_b200.plan import conv_taps
    NB = 512
    torch.manual_seed(17 + cin + cout + H)
    x = bf(torch.randn(NB, cin, H, H))
    w = torch.randn(cout, cin, k, k, requires_grad=True)
    g = bf(torch.randn(NB, cout, H, H) * 0.1)
    torch.set_num_threads(max(1, (os.cpu_count() or 8)))
    F.conv2d(x, w, None, 1, k // 2).backward(g)
    grad = torch.zeros(cout, cin, k, k, device="cuda")
    wgrad(sv, nhwc(x), nhwc(g).view(-1, cout), conv_taps(k, k // 2), NB, H, H, cin, H, H, cout, 1, grad, cout, cin,
          cin * k * k, k * k, 1, 1, 2)
    assert rel_rms(grad.cpu(), w.grad) < 2e-3


@pytest.mark.parametrize("cin,cout,H,k", [(32, 32, 32, 3), (64, 64, 16, 3), (128, 128, 8, 3)])
def test_cta_limit_leaves_results_unchanged(sv, cin, cout, H, k):
    """sv_set_cta_limit (the data-parallel backward leaves NCCL's SMs alone, ddp.GradReducer.cta_limit): the persistent conv
    and weight-gradient grids launched on fewer CTAs than SMs give the same convolution bit for bit (tile striding only)
    and the same weight gradient up to the summation order over the per-CTA partial sums"""
    from shotvae_b200.plan import conv_taps
    from shotvae_b200._abi import lib
    NB = 256
    torch.manual_seed(5 + cin + H)
    x, w = bf(torch.randn(NB, cin, H, H)), bf(torch.randn(cout, cin, k, k) * 0.1)
    g = bf(torch.randn(NB, cout, H, H) * 0.1)
    taps = conv_taps(k, k // 2)
    impl = 3 if cin < 128 else 0
    Wt = pack(sv, w, cout, cin, taps, cout, cin, cin * k * k, k * k, 1, impl)
    xa, ga = nhwc(x), nhwc(g).view(-1, cout)
    outs, grads = [], []
    try:
        for limit in (0, 144, 37):
            assert lib.sv_set_cta_limit(limit) in (0, 144, 37)
            out = torch.empty(NB, H, H, cout, dtype=torch.bfloat16, device="cuda")
            igemm(sv, xa, Wt, taps, NB, H, H, cin, H, H, cout, out=out, impl=impl)
            grad = torch.zeros(cout, cin, k, k, device="cuda")
            wgrad(sv, xa, ga, taps, NB, H, H, cin, H, H, cout, 1, grad, cout, cin, cin * k * k, k * k, 1, 1, 2)
            torch.cuda.synchronize()
            outs.append(out)
            grads.append(grad)
    finally:
        lib.sv_set_cta_limit(0)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    assert rel_rms(grads[1], grads[0]) < 1e-5 and rel_rms(grads[2], grads[0]) < 1e-5
    assert rel_rms(from_nhwc(outs[0]), F.conv2d(x, w, None, 1, k // 2)) < 4e-3


@pytest.mark.parametrize("impl", [0, 2])
@pytest.mark.parametrize("cin,cout,Hin,NB", [(1024, 512, 1, 16), (512, 256, 2, 8), (256, 128, 4, 8), (128, 64, 8, 4)])
def test_convT_phases_in_one_grid_equal_separate_launches(sv, impl, cin, cout, Hin, NB):
    """sv_igemm_fprop_batch: the four output-parity phases in one grid (per-tap tcgen05 kernel) or one after the
    other (any other kernel) give exactly what four sv_igemm_fprop calls give, statistics included"""
    from shotvae_b200._abi import lib, check, IgemmArgs
    from shotvae_b200.plan import dgrad_phase_taps, live_taps
    torch.manual_seed(cin)
    x, w = bf(torch.randn(NB, cin, Hin, Hin)), bf(torch.randn(cin, cout, 4, 4) * 0.05)
    Ho, A = 2 * Hin, nhwc(x)
    res = []
    for batched in (False, True):
        out = torch.zeros(NB, Ho, Ho, cout, dtype=torch.bfloat16, device="cuda")
        stats = torch.zeros(2, 2, cout, device="cuda")
        batch = [] if batched else None
        for (py, px), taps in dgrad_phase_taps(4, 2, 1).items():
            taps = live_taps(taps, Hin, Hin, Hin, Hin, 1)
            Wt = pack(sv, w, cout, cin, taps, cout, cin, 16, cout * 16, 1, 1)
            igemm(sv, A, Wt, taps, NB, Hin, Hin, cin, Hin, Hin, cout, out=out, stats=stats, out_stride=2, off=(py, px), OHf=Ho, OWf=Ho,
                  group_images=NB // 2, impl=impl, batch=batch)
        if batched:
            arr = (IgemmArgs * len(batch))(*[b[0] for b in batch])
            check(lib.sv_igemm_fprop_batch(arr, len(batch), sv.stream()))
        torch.cuda.synchronize()
        res.append((out.float().cpu(), stats.cpu()))
    assert torch.equal(res[0][0], res[1][0])
    assert rel_rms(res[1][1], res[0][1]) < 1e-5
    assert rel_rms(from_nhwc(res[1][0].cuda().to(torch.bfloat16)), F.conv_transpose2d(x, w, None, 2, 1)) < 4e-3


@pytest.mark.parametrize("cin,cout,c_real,Hin,NB", [(64, 16, 3, 16, 16), (128, 64, 64, 8, 4)])
def test_convT_phases_in_one_mma_grid_equal_separate_launches(sv, cin, cout, c_real, Hin, NB):
    """sv_igemm_fprop_batch on the mma.sync kernel (the last decoder layer, 64 -> 3 channels, fp32 output with n_valid):
    the four output-parity phases in ONE grid (blockIdx.z = phase) equal four separate launches bit for bit"""
    from shotvae_b200._abi import lib, check, IgemmArgs
    from shotvae_b200.plan import dgrad_phase_taps, live_taps
    torch.manual_seed(cin + cout)
    x = bf(torch.randn(NB, cin, Hin, Hin))
    w = torch.zeros(cin, cout, 4, 4)
    w[:, :c_real] = bf(torch.randn(cin, c_real, 4, 4) * 0.05)
    Ho, A = 2 * Hin, nhwc(x)
    res = []
    for batched in (False, True):
        outf = torch.zeros(NB, Ho, Ho, c_real, dtype=torch.float32, device="cuda")
        batch = [] if batched else None
        keep = []
        for (py, px), taps in dgrad_phase_taps(4, 2, 1).items():
            taps = live_taps(taps, Hin, Hin, Hin, Hin, 1)
            Wt = pack(sv, w, cout, cin, taps, c_real, cin, 16, cout * 16, 1, 1)
            keep.append(Wt)
            igemm(sv, A, Wt, taps, NB, Hin, Hin, cin, Hin, Hin, cout, outf=outf, out_stride=2, off=(py, px), OHf=Ho, OWf=Ho,
                  n_valid=c_real, impl=1, batch=batch)
        if batched:
            n0 = lib.sv_launch_count()
            arr = (IgemmArgs * len(batch))(*[b[0] for b in batch])
            check(lib.sv_igemm_fprop_batch(arr, len(batch), sv.stream()))
            assert lib.sv_launch_count() - n0 == 1
        torch.cuda.synchronize()
        res.append(outf.cpu())
    assert torch.equal(res[0], res[1])
    want = F.conv_transpose2d(x, w, None, 2, 1)[:, :c_real]
    assert rel_rms(res[1].permute(0, 3, 1, 2), want) < 4e-3


@pytest.mark.parametrize("impl", IMPLS)
@pytest.mark.parametrize("cin,cout,H,stride,k,NB", [(32, 32, 16, 1, 3, 4), (32, 64, 32, 2, 3, 4), (32, 64, 32, 2, 1, 4), (16, 32, 8, 1, 1, 4)])
def test_conv_dgrad_phases_match_autograd(sv, impl, cin, cout, H, stride, k, NB):
    from shotvae_b200.plan import dgrad_phase_taps, live_taps
    torch.manual_seed(7 + cin + stride)
    x = torch.randn(NB, cin, H, H, requires_grad=True)
    w = bf(torch.randn(cout, cin, k, k) * 0.1)
    Ho = H // stride
    g = bf(torch.randn(NB, cout, Ho, Ho))
    F.conv2d(x, w, None, stride, k // 2).backward(g)
    gin = torch.zeros(NB, H, H, cin, dtype=torch.bfloat16, device="cuda")
    G_ = nhwc(g)
    for (py, px), taps in dgrad_phase_taps(k, stride, k // 2).items():
        taps = live_taps(taps, Ho, Ho, Ho, Ho, 1)
        if not taps:
            continue
        Wt = pack(sv, w, cin, cout, taps, cin, cout, k * k, cin * k * k, 1, impl)
        igemm(sv, G_, Wt, taps, NB, Ho, Ho, cout, Ho, Ho, cin, out=gin, out_stride=stride, off=(py, px), OHf=H, OWf=H, impl=impl)
    assert rel_rms(from_nhwc(gin), x.grad) < 4e-3


@pytest.mark.parametrize("impl", [2, 3])
@pytest.mark.parametrize("cin,cout,H,k,NB,G,slope", [(32, 32, 32, 3, 64, 2, 0.01), (64, 64, 16, 3, 128, 4, 0.01), (128, 128, 8, 3, 256, 4, 0.0),
                                                     (32, 16, 32, 1, 40, 1, 1.0), (32, 32, 32, 3, 512, 4, 0.01), (160, 160, 8, 3, 16, 2, 0.01),
                                                     (640, 640, 8, 3, 64, 4, 0.01), (320, 320, 16, 3, 32, 4, 0.01), (160, 160, 32, 3, 16, 2, 0.01),
                                                     (512, 512, 4, 3, 64, 2, 0.0)])
def test_dgrad_epilogue_accumulates_bn_backward_statistics(sv, impl, cin, cout, H, k, NB, G, slope):
    """input-gradient conv with the fused BatchNorm-backward statistics epilogue == the same conv followed by
    sv_bn_bwd_reduce on its (bf16) output; (cin -> cout is the FORWARD conv: the launch maps cout gradients to cin)"""
    from shotvae_b200._abi import lib, check, ptr, taps_array, IgemmArgs
    from shotvae_b200.plan import dgrad_phase_taps
    torch.manual_seed(cin + cout + H + NB)
    w = bf(torch.randn(cout, cin, k, k) * 0.1)
    g_out = nhwc(bf(torch.randn(NB, cout, H, H)))
    y = nhwc(bf(torch.randn(NB, cin, H, H) * 1.5 + 0.3))
    scale, shift = (torch.rand(G, cin) + 0.5).cuda(), (torch.randn(G, cin) * 0.3).cuda()
    mean, var = (torch.randn(G, cin) * 0.2 + 0.3).cuda(), (torch.rand(G, cin) + 0.5).cuda()
    taps = dgrad_phase_taps(k, 1, k // 2)[(0, 0)]
    Wt = pack(sv, w, cin, cout, taps, cin, cout, k * k, cin * k * k, 1, impl)
    a = IgemmArgs()
    out = torch.empty(NB, H, H, cin, dtype=torch.bfloat16, device="cuda")
    stats = torch.zeros(2, G, cin, device="cuda")
    a.A, a.Wt, a.out_bf16, a.stats = ptr(g_out), ptr(Wt), ptr(out), ptr(stats)
    a.NB, a.H, a.W, a.C, a.OH, a.OW, a.N, a.T = NB, H, H, cout, H, H, cin, len(taps)
    a.in_stride, a.out_stride, a.OHf, a.OWf, a.group_images = 1, 1, H, H, NB // G
    a.dy, a.dx = taps_array([t[1] for t in taps]), taps_array([t[2] for t in taps])
    a.impl, a.w_layout = impl, (1 if impl == 3 else 0)
    a.bn_y, a.bn_scale, a.bn_shift, a.bn_mean, a.bn_var, a.bn_slope, a.bn_eps = ptr(y), ptr(scale), ptr(shift), ptr(mean), ptr(var), slope, 1e-5
    if not lib.sv_igemm_fprop_supports(C.byref(a), impl):
        pytest.skip("shape not covered by tcgen05 kernel %d" % impl)
    check(lib.sv_igemm_fprop(C.byref(a), sv.stream()))
    # reference: plain launch + the separate reduction pass over its output
    out2 = torch.empty_like(out)
    a.bn_y, a.stats, a.out_bf16 = None, None, ptr(out2)
    check(lib.sv_igemm_fprop(C.byref(a), sv.stream()))
    assert torch.equal(out, out2)
    dg, db = torch.zeros(G, cin, device="cuda"), torch.zeros(G, cin, device="cuda")
    check(lib.sv_bn_bwd_reduce(ptr(out2), None, ptr(y), ptr(scale), ptr(shift), ptr(mean), ptr(var), 1e-5, slope, (NB // G) * H * H, H * H, G, cin,
                               ptr(dg), ptr(db), sv.stream()))
    assert rel_rms(stats[0], db) < 1e-4 and rel_rms(stats[1], dg) < 1e-4
    # and against torch on the same bf16 values
    go = out2.float().view(G, -1, cin)
    yy = y.float().view(G, -1, cin)
    pre = yy * scale[:, None] + shift[:, None]
    dz = torch.where(pre > 0, go, go * slope)
    xh = (yy - mean[:, None]) * torch.rsqrt(var[:, None] + 1e-5)
    assert rel_rms(stats[0], dz.sum(1)) < 1e-3 and rel_rms(stats[1], (dz * xh).sum(1)) < 1e-3


def wgrad(sv, A, Gr, taps, NB, H, W, Cc, OH, OW, N, in_stride, grad, n_real, c_real, sn, sc, st, splits, impl=1):
    from shotvae_b200._abi import lib, check, ptr, taps_array, WgradArgs
    T = len(taps)
    a = WgradArgs()
    a.A, a.Gr = ptr(A), ptr(Gr)
    a.NB, a.H, a.W, a.C, a.OH, a.OW, a.N, a.T, a.in_stride, a.splits = NB, H, W, Cc, OH, OW, N, T, in_stride, splits
    a.dy, a.dx = taps_array([t[1] for t in taps]), taps_array([t[2] for t in taps])
    a.impl = impl
    if impl == 2:
        a.partial = ptr(grad)          # any non-null pointer: the query only inspects the geometry
        splits = lib.sv_igemm_wgrad_splits(C.byref(a))
        if splits <= 0:
            pytest.skip("shape not covered by the tcgen05 wgrad kernel")
        a.splits = splits
    ws = torch.full((splits * N * T * Cc,), float("nan"), device="cuda")
    a.partial = ptr(ws)
    check(lib.sv_igemm_wgrad(C.byref(a), sv.stream()))
    check(lib.sv_wgrad_reduce(ptr(ws), ptr(grad), splits, N, Cc, T, n_real, c_real, sn, sc, st, taps_array([t[0] for t in taps]),
                              sv.stream()))


@pytest.mark.parametrize("impl", [1, 2])
@pytest.mark.parametrize("cin,cout,H,stride,k,NB,splits", [(32, 32, 16, 1, 3, 4, 3), (32, 64, 32, 2, 3, 4, 5), (16, 32, 16, 1, 1, 4, 1),
                                                           (128, 128, 8, 1, 3, 8, 2), (160, 160, 8, 1, 3, 2, 1), (16, 16, 32, 1, 3, 2, 7),
                                                           (32, 32, 32, 1, 3, 6, 1), (64, 64, 16, 1, 3, 9, 1), (16, 32, 32, 1, 3, 3, 1),
                                                           (64, 128, 8, 1, 1, 5, 1),
                                                           # wide layers (WRN-28-10, PreActResNet18): TMA-fed tcgen05 kernel
                                                           (160, 160, 32, 1, 3, 4, 1), (320, 320, 16, 1, 3, 8, 1), (640, 640, 8, 1, 3, 16, 1),
                                                           (160, 160, 8, 1, 3, 32, 1), (256, 256, 8, 1, 3, 8, 1), (512, 512, 4, 1, 3, 32, 1),
                                                           (64, 160, 16, 1, 1, 4, 1), (160, 320, 16, 1, 1, 2, 1),
                                                           # stride 2 on the TMA kernel (tensor-map element strides)
                                                           (64, 128, 16, 2, 3, 8, 1), (160, 320, 32, 2, 3, 4, 1), (320, 640, 16, 2, 1, 8, 1),
                                                           (128, 256, 16, 2, 3, 2, 1),
                                                           # halo-tile kernel, output channels in power-of-two slices (128 + 32, 64 + 32)
                                                           (16, 160, 32, 1, 3, 4, 1), (16, 160, 32, 1, 1, 4, 1), (32, 96, 16, 1, 3, 4, 1)])
def test_conv_wgrad_matches_autograd(sv, impl, cin, cout, H, stride, k, NB, splits):
    from shotvae_b200.plan import conv_taps
    torch.manual_seed(11 + cin + stride + k)
    x = bf(torch.randn(NB, cin, H, H))
    w = torch.randn(cout, cin, k, k, requires_grad=True)
    Ho = H // stride
    g = bf(torch.randn(NB, cout, Ho, Ho))
    F.conv2d(x, w, None, stride, k // 2).backward(g)
    grad = torch.ones(cout, cin, k, k, device="cuda")      # the kernel accumulates (+=)
    wgrad(sv, nhwc(x), nhwc(g).view(-1, cout), conv_taps(k, k // 2), NB, H, H, cin, Ho, Ho, cout, stride, grad, cout, cin,
          cin * k * k, k * k, 1, splits, impl)
    assert rel_rms(grad.cpu() - 1.0, w.grad) < 2e-3


@pytest.mark.parametrize("impl", [1, 2])
@pytest.mark.parametrize("cin,cout,Hin,NB", [(1024, 512, 1, 8), (128, 64, 8, 4), (64, 3, 16, 4), (1024, 512, 1, 128), (512, 256, 2, 32),
                                             (256, 128, 4, 16)])
def test_convT_wgrad_and_dgrad_match_autograd(sv, impl, cin, cout, Hin, NB):
    from shotvae_b200.plan import conv_taps, live_taps, pad16
    torch.manual_seed(13 + cin)
    x = bf(torch.randn(NB, cin, Hin, Hin)).requires_grad_(True)
    w = bf(torch.randn(cin, cout, 4, 4) * 0.05).requires_grad_(True)
    Ho, cp = 2 * Hin, pad16(cout)
    g = bf(torch.randn(NB, cout, Ho, Ho))
    F.conv_transpose2d(x, w, None, 2, 1).backward(g)
    gp = torch.zeros(NB, cp, Ho, Ho)
    gp[:, :cout] = g
    Gd = nhwc(gp)
    taps = live_taps(conv_taps(4, 1), Hin, Hin, Ho, Ho, 2)
    grad = torch.zeros(cin, cout, 4, 4, device="cuda")
    # (impl 2: the TMA-fed tcgen05 kernels with element-strided activation boxes; each half skips when its shape is not covered)
    ran = 0
    a = _wgrad_covered(sv, Gd, nhwc(x.detach()).view(-1, cin), taps, NB, Ho, Ho, cp, Hin, Hin, cin, 2, impl)
    if a:
        wgrad(sv, Gd, nhwc(x.detach()).view(-1, cin), taps, NB, Ho, Ho, cp, Hin, Hin, cin, 2, grad, cin, cout, cout * 16, 16, 1, 2, impl)
        assert rel_rms(grad.cpu(), w.grad) < 2e-3
        ran += 1
    Wt = pack(sv, w.detach(), cin, cp, taps, cin, cout, cout * 16, 16, 1)
    gin = torch.zeros(NB, Hin, Hin, cin, dtype=torch.bfloat16, device="cuda")
    if impl == 1 or _igemm_covered(sv, Gd, Wt, taps, NB, Ho, Ho, cp, Hin, Hin, cin, 2, gin, impl):
        igemm(sv, Gd, Wt, taps, NB, Ho, Ho, cp, Hin, Hin, cin, in_stride=2, out=gin, impl=impl)
        assert rel_rms(from_nhwc(gin), x.grad) < 4e-3
        ran += 1
    if not ran:
        pytest.skip("shape not covered by the tcgen05 kernels")


def _wgrad_covered(sv, A, Gr, taps, NB, H, W, Cc, OH, OW, N, in_stride, impl):
    from shotvae_b200._abi import lib, ptr, taps_array, WgradArgs
    if impl != 2:
        return True
    a = WgradArgs()
    a.A, a.Gr, a.partial = ptr(A), ptr(Gr), ptr(A)
    a.NB, a.H, a.W, a.C, a.OH, a.OW, a.N, a.T, a.in_stride, a.splits = NB, H, W, Cc, OH, OW, N, len(taps), in_stride, 1
    a.dy, a.dx = taps_array([t[1] for t in taps]), taps_array([t[2] for t in taps])
    a.impl = 2
    return lib.sv_igemm_wgrad_splits(C.byref(a)) > 0


def _igemm_covered(sv, A, Wt, taps, NB, H, W, Cc, OH, OW, N, in_stride, out, impl):
    from shotvae_b200._abi import lib, ptr, taps_array, IgemmArgs
    a = IgemmArgs()
    a.A, a.Wt, a.out_bf16 = ptr(A), ptr(Wt), ptr(out)
    a.NB, a.H, a.W, a.C, a.OH, a.OW, a.N, a.T = NB, H, W, Cc, OH, OW, N, len(taps)
    a.in_stride, a.out_stride, a.OHf, a.OWf, a.group_images = in_stride, 1, OH, OW, NB
    a.dy, a.dx = taps_array([t[1] for t in taps]), taps_array([t[2] for t in taps])
    a.impl = impl
    return bool(lib.sv_igemm_fprop_supports(C.byref(a), impl))


# ------------------------------------------------------------------------------------------ BatchNorm
@pytest.mark.parametrize("Cc,HW,B,G,slope", [(32, 64, 8, 2, 0.01), (160, 16, 4, 1, 0.0), (1024, 1, 16, 2, 0.0), (16, 1024, 4, 1, 1.0)])
def test_bn_act_forward_backward(sv, Cc, HW, B, G, slope):
    from shotvae_b200._abi import lib, check, ptr, BnBwdTerm
    torch.manual_seed(Cc + HW)
    NB = G * B
    y = bf(torch.randn(NB, HW, Cc) * 2 + 0.5)
    gamma, beta = torch.rand(Cc) + 0.5, torch.randn(Cc) * 0.1
    g_a = bf(torch.randn(NB, HW, Cc))
    addend = bf(torch.randn(NB, HW, Cc))
    # torch reference, per group
    wants, gys, dgs, dbs = [], [], [], []
    gam_ref = gamma.clone().requires_grad_(True)
    bet_ref = beta.clone().requires_grad_(True)
    for g in range(G):
        yy = y[g * B:(g + 1) * B].reshape(-1, Cc).clone().requires_grad_(True)
        o = F.batch_norm(yy, None, None, gam_ref, bet_ref, True, 0.1, 1e-5)
        o = torch.where(o > 0, o, o * slope)
        wants.append(o.detach())
        o.backward(g_a[g * B:(g + 1) * B].reshape(-1, Cc))
        gys.append(yy.grad + addend[g * B:(g + 1) * B].reshape(-1, Cc))
    st = sv.stream()
    yd = y.to(torch.bfloat16).cuda()
    stats = torch.stack([torch.stack([y[g * B:(g + 1) * B].sum(dim=(0, 1)), (y[g * B:(g + 1) * B] ** 2).sum(dim=(0, 1))]) for g in range(G)]).cuda()
    mean, var, scale, shift = (torch.empty(G, Cc, device="cuda") for _ in range(4))
    gamma_d, beta_d, addend_d = gamma.cuda(), beta.cuda(), addend.to(torch.bfloat16).cuda()   # keep alive: ptr() of a temporary dangles
    check(lib.sv_bn_finalize(ptr(stats), ptr(gamma_d), ptr(beta_d), float(B * HW), 1e-5, G, Cc, Cc, ptr(mean), ptr(var),
                             ptr(scale), ptr(shift), st))
    a = torch.empty_like(yd)
    check(lib.sv_bn_act_fwd(ptr(yd), ptr(a), ptr(scale), ptr(shift), slope, B * HW, G, Cc, st))
    assert rel_rms(a.float().cpu().view(-1, Cc), torch.cat(wants)) < 4e-3
    dg, db = torch.zeros(G, Cc, device="cuda"), torch.zeros(G, Cc, device="cuda")
    gad = g_a.to(torch.bfloat16).cuda()
    check(lib.sv_bn_bwd_reduce(ptr(gad), None, ptr(yd), ptr(scale), ptr(shift), ptr(mean), ptr(var), 1e-5, slope, B * HW, HW, G, Cc,
                               ptr(dg), ptr(db), st))
    assert rel_rms(dg.sum(0).cpu(), gam_ref.grad) < 1e-3
    assert rel_rms(db.sum(0).cpu(), bet_ref.grad) < 1e-3
    gg, gb = torch.zeros(Cc, device="cuda"), torch.zeros(Cc, device="cuda")
    term = (BnBwdTerm * 1)()
    term[0].g_a, term[0].scale, term[0].shift, term[0].mean, term[0].var = ptr(gad), ptr(scale), ptr(shift), ptr(mean), ptr(var)
    term[0].dgamma, term[0].dbeta, term[0].grad_gamma, term[0].grad_beta = ptr(dg), ptr(db), ptr(gg), ptr(gb)
    term[0].slope, term[0].c_real = slope, Cc
    gy = torch.empty_like(yd)
    check(lib.sv_bn_bwd_apply(term, 1, ptr(yd), ptr(addend_d), ptr(gy), 1e-5, B * HW, HW, G, Cc, st))
    assert rel_rms(gy.float().cpu().view(-1, Cc), torch.cat(gys)) < 6e-3
    assert rel_rms(gg.cpu(), gam_ref.grad) < 1e-3 and rel_rms(gb.cpu(), bet_ref.grad) < 1e-3


# ------------------------------------------------------------------------------ loss / sample / mixup
def test_vae_criterion_matches_oracle(sv):
    from oracle import shotvae_oracle as O
    from lib.criterion import VAECriterion, ClsCriterion
    torch.manual_seed(3)
    B, nd = 32, 10
    for bce in (True, False):
        x = torch.rand(B, 3, 32, 32)
        xr = (torch.randn(B, 3, 32, 32) * 2).requires_grad_(True)
        mu, ls = torch.randn(B, 128).requires_grad_(True), (torch.randn(B, 128) * 0.3).requires_grad_(True)
        la = F.log_softmax(torch.randn(B, nd), 1).requires_grad_(True)
        want = O.vae_criterion(x, xr, mu, ls, la, nd, 1.0, bce)
        (want[0] * 0.3 + want[1] * 1.7 - want[2] * 2.0).backward()
        d = [t.detach().cuda().requires_grad_(True) for t in (xr, mu, ls, la)]
        got = VAECriterion(nd, 1, bce).cuda()(x.cuda(), *d)
        (got[0] * 0.3 + got[1] * 1.7 - got[2] * 2.0).backward()
        for gv, wv in zip(got, want):
            assert abs(float(gv) - float(wv)) <= 1e-5 * abs(float(wv))
        for gt, wt in zip(d, (xr, mu, ls, la)):
            assert rel_rms(gt.grad, wt.grad) < 1e-5
    # NHWC-backed reconstruction view (what the VAE returns internally)
    xr2 = xr.detach().permute(0, 2, 3, 1).contiguous().cuda().permute(0, 3, 1, 2).requires_grad_(True)
    got = VAECriterion(nd, 1, False).cuda()(x.cuda(), xr2, d[1], d[2], d[3])
    assert abs(float(got[0]) - float(want[0])) <= 1e-5 * abs(float(want[0]))
    pred = F.log_softmax(torch.randn(B, nd), 1).requires_grad_(True)
    tgt = torch.rand(B, nd)
    w = O.cls_criterion(pred, tgt)
    w.backward()
    pd = pred.detach().cuda().requires_grad_(True)
    gv = ClsCriterion()(pd, tgt.cuda())
    gv.backward()
    assert abs(float(gv) - float(w)) < 1e-5 * abs(float(w)) and rel_rms(pd.grad, pred.grad) < 1e-5


def test_sample_forward_backward_matches_oracle(sv):
    from oracle import shotvae_oracle as O
    from shotvae_b200._abi import lib, check, ptr
    torch.manual_seed(5)
    B, D, nd, T = 16, 128, 10, 0.67
    for mode in (0, 1, 2):
        mu, ls = torch.randn(B, D).requires_grad_(True), (torch.randn(B, D) * 0.3).requires_grad_(True)
        la = F.log_softmax(torch.randn(B, nd), 1).requires_grad_(True)
        eps, unif = torch.randn(B, D), torch.rand(B, nd)
        lab, lab2, lam = torch.randint(0, nd, (B,)), torch.randint(0, nd, (B,)), 0.3
        log = [("randn", eps)] + ([("rand", unif)] if mode == 2 else [])
        want = O.sample_latent(mu, ls, la, O.ReplayDraws(log), T, None if mode == 2 else lab, mode == 1, lab2, lam).view(B, -1)
        gl = torch.randn(B, D + nd)
        want.backward(gl)
        lat = torch.empty(B, D + nd, device="cuda")
        lam_dev = torch.tensor([lam, 1 - lam], dtype=torch.float32).cuda()
        md, lsd, lad = mu.detach().cuda(), ls.detach().cuda(), la.detach().cuda()
        epd, und, labd, lab2d, gld = eps.cuda(), unif.cuda(), lab.cuda(), lab2.cuda(), gl.cuda()
        check(lib.sv_sample_fwd(ptr(md), ptr(lsd), ptr(lad), ptr(epd), ptr(und), ptr(labd), ptr(lab2d), ptr(lam_dev), mode, T,
                                B, D, nd, ptr(lat), D + nd, sv.stream()))
        assert rel_rms(lat, want.detach()) < 1e-5
        g_mu, g_ls, g_la = (torch.zeros_like(t) for t in (md, lsd, lad))
        check(lib.sv_sample_bwd(ptr(gld), D + nd, ptr(lsd), ptr(epd), ptr(lat), mode, T, B, D, nd, ptr(g_mu), ptr(g_ls), ptr(g_la),
                                1, sv.stream()))
        assert rel_rms(g_mu, mu.grad) < 1e-5 and rel_rms(g_ls, ls.grad) < 1e-5
        if mode == 2:
            assert rel_rms(g_la, la.grad) < 1e-4


def test_mixup_and_label_smoothing_match_oracle(sv):
    from oracle import shotvae_oracle as O
    import lib.utils.mixup as MX
    B, nd = 24, 10
    torch.manual_seed(9)
    img, mu, ls = torch.rand(B, 3, 32, 32), torch.randn(B, 128), torch.randn(B, 128) * 0.3
    la, lab = F.log_softmax(torch.randn(B, nd), 1), torch.randint(0, nd, (B,))
    for fn in ("mixup", "smooth"):
        torch.manual_seed(21); np.random.seed(21)
        d = O.LiveDraws()
        want = O.mixup_vae_data(img, mu, ls, la, d) if fn == "mixup" else O.label_smoothing(img, mu, ls, la, d, 0.1, lab)
        torch.manual_seed(21); np.random.seed(21)
        args = [t.cuda() for t in (img, mu, ls, la)]
        got = MX.mixup_vae_data(*args) if fn == "mixup" else MX.label_smoothing(*args, epsilon=0.1, disc_label=lab.cuda())
        for i in range(4):
            assert torch.equal(got[i].cpu(), want[i]) or rel_rms(got[i], want[i]) < 1e-6, (fn, i)
        if fn == "mixup":
            assert got[4] == want[4]
        else:
            assert torch.equal(got[4].cpu(), want[4]) and got[5] == want[5]


def test_om_pairing_bit_exact_vs_reference_golden(sv):
    import lib.utils.mixup as MX
    from oracle import shotvae_oracle as O
    path = os.path.join(os.path.dirname(__file__), "golden", "c2_wrn28x2_nd10_b32_e400_om.json")
    g = json.load(open(path))
    mu, ls = torch.tensor(g["om_mu"]), torch.tensor(g["om_ls"])
    idx, kl = MX.optimal_match_index(mu.cuda(), ls.cuda(), return_matrix=True)
    assert idx.cpu().tolist() == g["om_index"]                       # the reference's own loop (mixup.py:11-18)
    ref = O.pairwise_kl_matrix(mu, ls)
    assert float((kl.cpu() - ref).abs().max()) < 0.25 * g["om_min_gap_2nd_3rd"]
    assert torch.all(torch.diagonal(kl.cpu()) == 0)
    # larger seeded cases against the oracle: N(0,1) latents and near-duplicate latents
    for B, seed, scale in ((128, 1, 1.0), (128, 2, 0.05), (200, 3, 0.3), (2, 4, 1.0)):
        gen = torch.Generator().manual_seed(seed)
        mu, ls = torch.randn(B, 128, generator=gen) * scale, torch.randn(B, 128, generator=gen) * 0.1 * scale
        want = O.optimal_match_index(mu, ls)
        got = MX.optimal_match_index(mu.cuda(), ls.cuda())
        assert got.cpu().tolist() == want.tolist(), (B, seed)


def test_posterior_and_inference_kl_match_oracle(sv):
    from oracle import shotvae_oracle as O
    from shotvae_b200._abi import lib, check, ptr
    torch.manual_seed(17)
    B, D, nd, lam = 32, 128, 10, 0.35
    la = F.log_softmax(torch.randn(B, nd), 1).requires_grad_(True)
    mu, ls = torch.randn(B, D).requires_grad_(True), (torch.randn(B, D) * 0.3).requires_grad_(True)
    mu_t, sig_t = torch.randn(B, D), torch.rand(B, D) + 0.5
    la_, lb_ = torch.randint(0, nd, (B,)), torch.randint(0, nd, (B,))
    oh = lambda y: torch.zeros(B, nd).scatter_(1, y.view(-1, 1), 1)
    disc = lam * O.cls_criterion(la, oh(la_)) + (1 - lam) * O.cls_criterion(la, oh(lb_))
    cont = (F.mse_loss(mu, mu_t, reduction="sum") + F.mse_loss(torch.exp(ls), sig_t, reduction="sum")) / B
    cd, cc = 1.3, 0.07
    (cd * disc + cc * cont).backward()
    terms = torch.zeros(2, device="cuda")
    g_la, g_mu, g_ls = torch.zeros(B, nd, device="cuda"), torch.zeros(B, D, device="cuda"), torch.zeros(B, D, device="cuda")
    coef = torch.tensor([cd, cc], device="cuda")
    lam_dev = torch.tensor([lam, 1 - lam], dtype=torch.float32).cuda()
    d = [t.detach().cuda() for t in (la, la_, lb_, mu, ls, mu_t, sig_t)]
    check(lib.sv_posterior_fwd_bwd(ptr(d[0]), None, ptr(d[1]), ptr(d[2]), ptr(lam_dev), ptr(d[3]), ptr(d[4]), ptr(d[5]), ptr(d[6]),
                                   ptr(coef), B, D, nd, ptr(terms), ptr(g_la), ptr(g_mu), ptr(g_ls), 0, sv.stream()))
    assert abs(float(terms[0]) - float(disc)) < 1e-5 * abs(float(disc)) and abs(float(terms[1]) - float(cont)) < 1e-5 * float(cont)
    assert rel_rms(g_la, la.grad) < 1e-5 and rel_rms(g_mu, mu.grad) < 1e-5 and rel_rms(g_ls, ls.grad) < 1e-5
    lab = torch.randint(0, nd, (B,))
    lsu = torch.zeros(B, nd).scatter_(1, lab.view(-1, 1), 1 - 0.001 - 0.001 / (nd - 1)) + 0.001 / (nd - 1)
    au = torch.exp(la.detach())
    want = float(torch.sum(au * la.detach() - au * torch.log(lsu)) / B)
    out = torch.zeros(1, device="cuda")
    labd = lab.cuda()
    check(lib.sv_inference_kl(ptr(d[0]), ptr(labd), B, nd, ptr(out), sv.stream()))
    assert abs(float(out) - want) < 1e-5 * abs(want)


def test_sgd_matches_oracle(sv):
    from shotvae_b200._abi import lib, check, ptr
    torch.manual_seed(23)
    n = 100003
    p, g, m = torch.randn(n + 1), torch.randn(n + 1), torch.randn(n + 1)
    for first in (1.0, 0.0):
        d = g * 0.5 + 5e-4 * p
        mm = d if first else 0.9 * m + d
        want_p = p - 0.1 * mm
        pd, gd, md = p.cuda(), g.cuda(), m.cuda()
        hyper = torch.tensor([0.1, 0.9, 5e-4, 0.5, first], device="cuda")
        check(lib.sv_sgd_step(ptr(pd), ptr(gd), ptr(md), ptr(hyper), n, sv.stream()))
        assert rel_rms(pd[:n], want_p[:n]) < 1e-6 and rel_rms(md[:n], mm[:n]) < 1e-6
        assert float(gd[:n].abs().max()) == 0.0 and float(gd[n]) == float(g[n])


def test_linear_kernels_match_torch(sv):
    from shotvae_b200._abi import lib, check, ptr
    torch.manual_seed(29)
    for B, N, K, kn in ((32, 128, 128, 0), (24, 10, 640, 0), (16, 1024, 138, 1)):
        x = torch.randn(B, K)
        W = (torch.randn(K, N) if kn else torch.randn(N, K)).requires_grad_(True)
        bias = torch.randn(N).requires_grad_(True)
        xr = x.clone().requires_grad_(True)
        want = xr @ (W if kn else W.t()) + bias
        g = torch.randn(B, N)
        want.backward(g)
        out = torch.empty(B, N, device="cuda")
        st = sv.stream()
        xd, Wd, bd, gd = x.cuda(), W.detach().cuda(), bias.detach().cuda(), g.cuda()
        check(lib.sv_linear_fwd(ptr(xd), K, ptr(Wd), N if kn else K, kn, ptr(bd), ptr(out), None, N, None, 0, B, N, K, st))
        assert rel_rms(out, want.detach()) < 1e-5
        gx = torch.zeros(B, K, device="cuda")
        check(lib.sv_linear_bwd_input(ptr(gd), None, N, ptr(Wd), N if kn else K, kn, ptr(gx), K, 0, B, N, K, st))
        assert rel_rms(gx, xr.grad) < 1e-5
        dW, db = torch.zeros_like(W.detach()).cuda(), torch.zeros(N, device="cuda")
        check(lib.sv_linear_bwd_weight(ptr(gd), None, N, ptr(xd), K, ptr(dW), N if kn else K, kn, ptr(db), B, N, K, st))
        assert rel_rms(dW, W.grad) < 1e-5 and rel_rms(db, bias.grad) < 1e-5


def test_linear_split_reductions_match_torch(sv):
    """shapes that take the split paths: the decoder stem's input gradient (reduction over 1024 outputs split over gridDim.z,
    bf16 and fp32 gradients, with and without accumulation) and batch-sliced weight gradients (B = 512)"""
    from shotvae_b200._abi import lib, check, ptr
    torch.manual_seed(31)
    B, N, K = 256, 1024, 138
    W = torch.randn(K, N)                      # ConvTranspose2d k=1 stem: [latent][c0] = W(n,k) with w_kn = 1
    g = torch.randn(B, N)
    st = sv.stream()
    Wd = W.cuda()
    for dt in (torch.float32, torch.bfloat16):
        gd = g.to(dt).cuda()
        want = gd.float().cpu() @ W.t()
        gx = torch.full((B, K), 7.0, device="cuda")          # stale contents must be overwritten
        a32, a16 = (ptr(gd), None) if dt == torch.float32 else (None, ptr(gd))
        check(lib.sv_linear_bwd_input(a32, a16, N, ptr(Wd), N, 1, ptr(gx), K, 0, B, N, K, st))
        assert rel_rms(gx, want) < 1e-5
        check(lib.sv_linear_bwd_input(a32, a16, N, ptr(Wd), N, 1, ptr(gx), K, 1, B, N, K, st))
        assert rel_rms(gx, 2 * want) < 1e-5
    for Bw, Nw, Kw, kn in ((512, 128, 128, 0), (512, 10, 128, 0), (256, 1024, 138, 1), (200, 100, 128, 0)):
        x, gw = torch.randn(Bw, Kw), torch.randn(Bw, Nw)
        want_w = (x.t() @ gw) if kn else (gw.t() @ x)
        dW, db = torch.ones_like(want_w).cuda(), torch.ones(Nw, device="cuda")     # += semantics
        gwd, xd = gw.cuda(), x.cuda()
        check(lib.sv_linear_bwd_weight(ptr(gwd), None, Nw, ptr(xd), Kw, ptr(dW), Nw if kn else Kw, kn, ptr(db), Bw, Nw, Kw, st))
        assert rel_rms(dW, want_w + 1) < 1e-5 and rel_rms(db, gw.sum(0) + 1) < 1e-5


@pytest.mark.parametrize("B,nd", [(512, 10), (200, 100), (24, 10)])
def test_heads_launches_match_torch(sv, B, nd):
    """sv_heads_fwd / sv_heads_bwd_input / sv_heads_bwd_weight (the three inference heads of vae.py:113-129 in one launch each)
    against torch FP32; the forward is the arithmetic of sv_linear_fwd, so the two must agree bit for bit"""
    from shotvae_b200._abi import lib, check, ptr, Heads
    torch.manual_seed(37)
    K, Ns = 128, (128, 128, nd)
    x = torch.randn(B, K)
    Ws = [torch.randn(n, K) * 0.1 for n in Ns]
    bs = [torch.randn(n) for n in Ns]
    gs = [torch.randn(B, n) for n in Ns]
    st = sv.stream()
    xd = x.cuda()
    Wd, bd, gd = [w.cuda() for w in Ws], [b.cuda() for b in bs], [g.cuda() for g in gs]
    outs = [torch.empty(B, n, device="cuda") for n in Ns]
    dW, db = [torch.ones(n, K, device="cuda") for n in Ns], [torch.ones(n, device="cuda") for n in Ns]
    h = Heads()
    h.n = 3
    for i in range(3):
        h.W[i], h.bias[i], h.out[i], h.g[i], h.dW[i], h.dbias[i], h.N[i] = ptr(Wd[i]), ptr(bd[i]), ptr(outs[i]), ptr(gd[i]), ptr(dW[i]), ptr(db[i]), Ns[i]
    check(lib.sv_heads_fwd(ptr(xd), K, C.byref(h), B, K, st))
    for i in range(3):
        assert rel_rms(outs[i], x @ Ws[i].t() + bs[i]) < 1e-5
        single = torch.empty(B, Ns[i], device="cuda")
        check(lib.sv_linear_fwd(ptr(xd), K, ptr(Wd[i]), K, 0, ptr(bd[i]), ptr(single), None, Ns[i], None, 0, B, Ns[i], K, st))
        assert torch.equal(single, outs[i])
    gx = torch.full((B, K), 3.0, device="cuda")
    check(lib.sv_heads_bwd_input(C.byref(h), ptr(gx), K, B, K, st))
    assert rel_rms(gx, sum(g @ w for g, w in zip(gs, Ws))) < 1e-5
    check(lib.sv_heads_bwd_weight(C.byref(h), ptr(xd), K, B, K, st))
    for i in range(3):
        assert rel_rms(dW[i], gs[i].t() @ x + 1) < 1e-5 and rel_rms(db[i], gs[i].sum(0) + 1) < 1e-5


def test_wgrad_reduce_batched_equals_single(sv):
    """sv_wgrad_reduce_batched over several weight tensors (more than one launch's worth of descriptors, ragged n_real / c_real,
    permuted taps, accumulate semantics) against sv_wgrad_reduce one tensor at a time"""
    from shotvae_b200._abi import lib, check, ptr, taps_array, ReduceDesc
    torch.manual_seed(41)
    st = sv.stream()
    shapes = [(148, 32, 32, 9, 32, 32), (20, 16, 16, 9, 16, 3), (51, 128, 128, 9, 128, 128), (7, 48, 64, 4, 40, 64), (1, 16, 16, 1, 16, 16)] * 9
    descs, want, got = [], [], []
    keep = []
    for i, (splits, N, Cc, T, n_real, c_real) in enumerate(shapes):
        part = torch.randn(splits, N, T * Cc, device="cuda")
        tap = list(reversed(range(T)))
        g0 = torch.randn(n_real, c_real, T, device="cuda")        # grad[n][c][tap]: sn = c_real*T, sc = T, st = 1
        a, b = g0.clone(), g0.clone()
        check(lib.sv_wgrad_reduce(ptr(part), ptr(a), splits, N, Cc, T, n_real, c_real, c_real * T, T, 1, taps_array(tap), st))
        d = ReduceDesc()
        d.partial, d.grad, d.sn, d.sc, d.st = ptr(part), ptr(b), c_real * T, T, 1
        d.splits, d.N, d.C, d.T, d.n_real, d.c_real = splits, N, Cc, T, n_real, c_real
        for k, v in enumerate(tap):
            d.tap_index[k] = v
        descs.append(d)
        want.append(a)
        got.append(b)
        keep.append(part)
        ref = g0.cpu() + part.sum(0).cpu().view(N, T, Cc)[:n_real, :, :c_real].permute(0, 2, 1).flip(2)
        assert rel_rms(a, ref) < 1e-5
    arr = (ReduceDesc * len(descs))(*descs)
    check(lib.sv_wgrad_reduce_batched(arr, len(descs), st))
    for a, b in zip(want, got):
        assert rel_rms(b, a) < 1e-6


def test_cpu_tensors_are_refused(sv):
    from lib.criterion import VAECriterion
    with pytest.raises(sv.ShotVaeError):
        VAECriterion(10, 1, True)(torch.rand(2, 3, 32, 32), torch.rand(2, 3, 32, 32), torch.rand(2, 128), torch.rand(2, 128),
                                  torch.rand(2, 10))
