"""A few launches of the block-1 input-gradient conv (3x3, 32 -> 32 channels, 32x32, NB = 512, 4 pass groups) for ncu captures:
python tools/halo_bnb_one.py [mode]   mode 0 = plain, 1 = fused BatchNorm-backward statistics, 2 = forward-style residual + statistics"""
import ctypes as C
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "shot-vae_b200"))
import torch
from shotvae_b200 import _abi
from shotvae_b200._abi import lib, check, ptr, taps_array, IgemmArgs
from shotvae_b200.plan import conv_taps

mode = int(sys.argv[1]) if len(sys.argv) > 1 else 1
NB, H, Cc, G = 512, 32, 32, 4
taps = conv_taps(3, 1)
torch.manual_seed(0)
w = torch.randn(Cc, Cc, 3, 3, device="cuda") * 0.1
Wt = torch.zeros(9, Cc, Cc, dtype=torch.bfloat16, device="cuda")
check(lib.sv_pack_weight(ptr(w), ptr(Wt), Cc, Cc, 9, Cc, Cc, Cc * 9, 9, 1, taps_array([t[0] for t in taps]), 1, _abi.stream()))
mk = lambda: torch.randn(NB, H, H, Cc, device="cuda").to(torch.bfloat16)
sets = [(mk(), mk(), torch.empty(NB, H, H, Cc, dtype=torch.bfloat16, device="cuda")) for _ in range(3)]
coef = [torch.rand(G, Cc, device="cuda") + 0.5 for _ in range(4)]
stats = torch.zeros(2, G, Cc, device="cuda")
a = IgemmArgs()
a.Wt = ptr(Wt)
a.NB, a.H, a.W, a.C, a.OH, a.OW, a.N, a.T = NB, H, H, Cc, H, H, Cc, 9
a.in_stride, a.out_stride, a.OHf, a.OWf, a.group_images = 1, 1, H, H, NB // G
a.dy, a.dx = taps_array([t[1] for t in taps]), taps_array([t[2] for t in taps])
a.impl, a.w_layout = 3, 1
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for it in range(7):
    A, Y, out = sets[it % 3]
    a.A, a.out_bf16 = ptr(A), ptr(out)
    a.stats = ptr(stats) if mode else None
    a.residual = ptr(Y) if mode == 2 else None
    if mode == 1:
        a.bn_y, a.bn_scale, a.bn_shift, a.bn_mean, a.bn_var, a.bn_slope, a.bn_eps = ptr(Y), ptr(coef[0]), ptr(coef[1]), ptr(coef[2]), ptr(coef[3]), 0.01, 1e-5
    if it == 3:
        ev[0].record()
    check(lib.sv_igemm_fprop(C.byref(a), _abi.stream()))
ev[1].record()
torch.cuda.synchronize()
print("mode", mode, "us/launch", 1e3 * ev[0].elapsed_time(ev[1]) / 4)
