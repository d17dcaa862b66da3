"""Run the block-1 halo conv a few times (for ncu captures): python tools/halo_one.py [epi] [NB]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "shot-vae_b200")); sys.path.insert(0, os.path.join(ROOT, "tools"))
import igemm_bench as ib
epi = bool(int(sys.argv[1])) if len(sys.argv) > 1 else False
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 512
shape = ("block1 3x3 32->32 @32x32", nb, 32, 32, 32, 3)
us, _ = ib.run(shape, 3, epi, nset=3, iters=4)
print("us/launch", us)
