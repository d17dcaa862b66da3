"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol include/shotvae.h
declares, the ctypes mirrors match the C structs, the drop-in modules own the reference's
state_dict (keys, shapes, default initialisation) and refuse CPU tensors, and the tap tables that
drive the implicit-GEMM kernels reproduce torch's conv / conv_transpose arithmetic."""
import os
import re

import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from shotvae_b200 import _abi
    hdr = open(os.path.join(ROOT, "include", "shotvae.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)     # prose in comments may mention function names
    declared = set(re.findall(r"\b(sv_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(_abi.lib, name), "libshotvae.so does not export %s" % name
    assert declared == set(_abi.EXPORTS), declared ^ set(_abi.EXPORTS)
    assert _abi.lib.sv_abi_version() == 1


def test_struct_mirrors_match():
    import ctypes as C
    from shotvae_b200 import _abi
    assert _abi.lib.sv_sizeof_igemm_args() == C.sizeof(_abi.IgemmArgs)
    assert _abi.lib.sv_sizeof_wgrad_args() == C.sizeof(_abi.WgradArgs)
    assert _abi.lib.sv_sizeof_bn_bwd_term() == C.sizeof(_abi.BnBwdTerm)


@pytest.mark.parametrize("net,nd", [("wideresnet-28-2", 10), ("wideresnet-28-2", 100), ("preactresnet18", 100), ("wideresnet-28-10", 10)])
def test_dropin_state_dict_equals_reference_init(net, nd):
    from shot_vae_model.vae import VariationalAutoEncoder
    from oracle import shotvae_oracle as O
    torch.manual_seed(1)
    m = VariationalAutoEncoder(net, 3, 0, (32, 32), True, 128, nd, 0.67, True)
    sd, st = m.state_dict(), O.init_state(net, nd)          # O.init_state is pinned to the reference by the goldens
    assert list(sd.keys()) == list(st.keys())
    assert all(torch.equal(sd[k], st[k]) for k in sd)
    dp = {k.replace("wideblock1.", "wideblock1.module."): v for k, v in sd.items()}
    m.load_state_dict(dp)                                   # nn.DataParallel key form is accepted


def test_unsupported_configurations_raise_like_the_reference():
    from shot_vae_model.vae import VariationalAutoEncoder
    with pytest.raises(NotImplementedError):
        VariationalAutoEncoder("vgg16", 3, 0, (32, 32), True, 128, 10, 0.67, True)       # vae.py:106
    with pytest.raises(NotImplementedError):
        VariationalAutoEncoder("wideresnet-28-2", 3, 0, (160, 160), True, 128, 10, 0.67, False)
    with pytest.raises(AssertionError):
        VariationalAutoEncoder("wideresnet-27-2", 3, 0, (32, 32), True, 128, 10, 0.67, True)  # wideresnet.py:72


def test_cpu_tensors_fail_loudly():
    from shot_vae_model.vae import VariationalAutoEncoder
    from shotvae_b200._abi import ShotVaeError
    m = VariationalAutoEncoder("wideresnet-10-1", 3, 0, (32, 32), True, 128, 10, 0.67, True)
    with pytest.raises(ShotVaeError):
        m(torch.rand(2, 3, 32, 32))


def _gather_conv(x, w_tap, taps, OH, OW, in_stride, out_stride, off, OHf, OWf, out=None):
    """pure-torch model of sv_igemm_fprop's tap semantics (NCHW for convenience)"""
    NB, Cc, H, W = x.shape
    N = w_tap.shape[1]
    out = torch.zeros(NB, N, OHf, OWf, dtype=x.dtype) if out is None else out
    for t, (_, dy, dx) in enumerate(taps):
        for oh in range(OH):
            ih = oh * in_stride + dy
            if not 0 <= ih < H:
                continue
            for ow in range(OW):
                iw = ow * in_stride + dx
                if 0 <= iw < W:
                    out[:, :, oh * out_stride + off[0], ow * out_stride + off[1]] += x[:, :, ih, iw] @ w_tap[t].t()
    return out


@pytest.mark.parametrize("k,s,pad,H", [(3, 1, 1, 6), (3, 2, 1, 8), (1, 2, 0, 8), (4, 2, 1, 4)])
def test_dgrad_phase_tap_tables_reproduce_conv_input_gradient(k, s, pad, H):
    from shotvae_b200.plan import dgrad_phase_taps, live_taps
    torch.manual_seed(k * 10 + s)
    cin, cout, NB = 3, 4, 2
    x = torch.randn(NB, cin, H, H, dtype=torch.double, requires_grad=True)
    w = torch.randn(cout, cin, k, k, dtype=torch.double)
    y = F.conv2d(x, w, None, s, pad)
    Ho = y.shape[-1]
    g = torch.randn_like(y)
    y.backward(g)
    got = torch.zeros(NB, cin, H, H, dtype=torch.double)
    for (py, px), taps in dgrad_phase_taps(k, s, pad).items():
        taps = live_taps(taps, Ho, Ho, Ho, Ho, 1)
        if not taps:
            continue
        w_tap = torch.stack([w[:, :, t[0] // k, t[0] % k].t() for t in taps])       # [T][n=ci][c=co]
        nph = len(range(py, H, s))
        _gather_conv(g, w_tap, taps, nph, nph, 1, s, (py, px), H, H, out=got)
    assert torch.allclose(got, x.grad, atol=1e-10)


@pytest.mark.parametrize("Hin", [1, 2, 4])
def test_convT_phase_tables_reproduce_conv_transpose(Hin):
    from shotvae_b200.plan import dgrad_phase_taps, live_taps, conv_taps
    torch.manual_seed(Hin)
    cin, cout, NB = 5, 3, 2
    x = torch.randn(NB, cin, Hin, Hin, dtype=torch.double, requires_grad=True)
    w = torch.randn(cin, cout, 4, 4, dtype=torch.double)
    want = F.conv_transpose2d(x, w, None, 2, 1)
    got = torch.zeros_like(want)
    n_live = 0
    for (py, px), taps in dgrad_phase_taps(4, 2, 1).items():
        taps = live_taps(taps, Hin, Hin, Hin, Hin, 1)
        n_live += len(taps)
        w_tap = torch.stack([w[:, :, t[0] // 4, t[0] % 4].t() for t in taps])       # [T][n=co][c=ci]
        _gather_conv(x.detach(), w_tap, taps, Hin, Hin, 1, 2, (py, px), 2 * Hin, 2 * Hin, out=got)
    assert torch.allclose(got, want, atol=1e-10)
    if Hin == 1:
        assert n_live == 4          # 12 of the 16 taps of the 1x1 -> 2x2 layer never touch the output
    # input gradient of the transposed conv == strided conv with the same taps
    g = torch.randn_like(want)
    want.backward(g)
    taps = live_taps(conv_taps(4, 1), Hin, Hin, 2 * Hin, 2 * Hin, 2)
    w_tap = torch.stack([w[:, :, t[0] // 4, t[0] % 4] for t in taps])               # [T][n=ci][c=co]
    gin = _gather_conv(g, w_tap, taps, Hin, Hin, 2, 1, (0, 0), Hin, Hin)
    assert torch.allclose(gin, x.grad, atol=1e-10)


def test_product_package_never_imports_the_oracle():
    """the oracle is test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use it"""
    import re
    pkg = os.path.join(ROOT, "shot-vae_b200")
    bad = []
    for dirpath, _, files in os.walk(pkg):
        if os.sep + "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M) or "shotvae_oracle" in txt:
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad
