"""Dump the per-tile role timeline of CTA 0 of the halo kernel (SHOTVAE_HALO_TRACE=1)."""
import ctypes as C, os, sys
os.environ["SHOTVAE_HALO_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "shot-vae_b200")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import igemm_bench as ib
from shotvae_b200._abi import lib
for shape, epi in ((ib.SHAPES[0], False), (ib.SHAPES[0], True), (ib.SHAPES[2], True)):
    us, _ = ib.run(shape, 3, epi, nset=2, iters=3)
    buf = (C.c_longlong * (4 * 64 * 2 + 240 + 8))()
    lib.sv_debug_halo_trace(buf)
    marks = list(buf)[752:760]
    m = torch.tensor(list(buf)[512:752]).view(3, 80)
    t = torch.tensor(list(buf)[:512]).view(4, 64, 2)
    t0 = int(t[0, 0, 0])
    print(shape[0], "us/launch", round(us, 1))
    print("marks (cycles after kernel entry): setup_done", marks[1] - marks[0], "weights_ready", marks[2] - marks[0], "first_prod_issue", t0 - marks[0], "thread0_at_final_sync", marks[3] - marks[0], "after_final_sync", marks[4] - marks[0])
    print("tile  prod_issue prod_publish | mma_accwait_begin mma_accwait_end mma_begin mma_end | epi_begin epi_end   (cycles since first producer issue)")
    for i in range(18):
        r = [int(t[0, i, 0]) - t0, int(t[0, i, 1]) - t0, int(t[2, i, 0]) - t0, int(t[2, i, 1]) - t0, int(t[1, i, 0]) - t0, int(t[1, i, 1]) - t0,
             int(t[3, i, 0]) - t0, int(t[3, i, 1]) - t0]
        print("%3d  %9d %9d | %9d %9d %9d %9d | %9d %9d" % tuple([i] + r))
    for k in range(3):
        b0 = int(t[1, k + 2, 0])
        print("tile", k + 2, "per-MMA issue stamps (cycles after mma_begin):", [int(v) - b0 for v in m[k, :20]], "commit1", int(m[k, 78]) - b0, "commit2", int(m[k, 79]) - b0)
    e0 = int(t[3, 3, 0])
    print("epilogue tile 3 stamps rel. to epi_begin [chunk: before_ld, after_ld, after_store, after_stats]:", [int(v) - e0 for v in m[0, 40:56]], "epi_end", int(t[3, 3, 1]) - e0)
