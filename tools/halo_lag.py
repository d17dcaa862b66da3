"""commit->epilogue lag and per-tile cadence for ablation variants (set SHOTVAE_HALO_ABLATE before running)"""
import ctypes as C, os, sys
os.environ["SHOTVAE_HALO_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "shot-vae_b200")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import igemm_bench as ib
from shotvae_b200._abi import lib
for epi in (False, True):
    us, _ = ib.run(ib.SHAPES[0], 3, epi, nset=2, iters=3)
    buf = (C.c_longlong * (4 * 64 * 2 + 240 + 8))()
    lib.sv_debug_halo_trace(buf)
    t = torch.tensor(list(buf)[:512]).view(4, 64, 2)
    lag = [int(t[3, i, 0] - t[1, i, 1]) for i in range(1, 14)]
    issue = [int(t[1, i, 1] - t[1, i, 0]) for i in range(1, 14)]
    epi_t = [int(t[3, i, 1] - t[3, i, 0]) for i in range(1, 14)]
    cad = int(t[1, 13, 0] - t[1, 1, 0]) // 12
    print("ablate=%s epi=%d us=%.1f cadence=%d | mma issue %s | commit->epilogue lag %s | epilogue %s" %
          (os.environ.get("SHOTVAE_HALO_ABLATE", "0"), epi, us, cad, issue[:6], lag[:8], epi_t[:6]))
