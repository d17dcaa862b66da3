"""smoke(): one tiny SHOT-VAE training iteration on cuda:0 through the drop-in modules (libshotvae
kernels), checked against the CPU oracle.  Used by __graft_entry__.smoke()."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F


def smoke(verbose=True):
    root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    if root not in sys.path:
        sys.path.insert(0, root)
    from oracle import shotvae_oracle as O           # checker only
    from shot_vae_model.vae import VariationalAutoEncoder
    from lib.criterion import VAECriterion, ClsCriterion
    from lib.utils.mixup import mixup_vae_data
    from shotvae_b200 import _abi
    assert torch.cuda.is_available(), "smoke() needs a CUDA device; libshotvae has no CPU path"
    torch.cuda.set_device(0)
    net, nd, B, epoch = "wideresnet-10-1", 10, 8, 100
    hyper = O.default_hyper("Cifar10")
    s = O.schedules(hyper, epoch)
    st = O.init_state(net, nd)
    il, ll, iu, lu = O.synthetic_batch(B, nd, 3)
    ost = O.clone_state(st)
    torch.manual_seed(2); np.random.seed(2)
    want = O.shot_step(ost, net, nd, il, ll, iu, lu, epoch, hyper, O.LiveDraws())
    model = VariationalAutoEncoder(net, 3, 0, (32, 32), True, 128, nd, 0.67, True)
    model.load_state_dict(st)
    model = model.cuda().train()
    crit, cls = VAECriterion(nd, 1, True).cuda(), ClsCriterion()
    n0 = _abi.launch_count()
    torch.manual_seed(2); np.random.seed(2)
    sys.path.insert(0, os.path.join(root, "tests"))
    from test_gpu_step import shot_loop_body
    got = shot_loop_body(model, crit, cls, il.cuda(), ll.cuda(), iu.cuda(), lu.cuda(), s, nd, False, hyper["epsilon"])
    torch.cuda.synchronize()
    launches = _abi.launch_count() - n0
    for k in ("rec_l", "klc_l", "rec_u", "klc_u"):
        r = abs(got[k] - want[k]) / abs(want[k])
        if verbose:
            print("smoke %-6s cuda %.6f oracle %.6f rel %.2e" % (k, got[k], want[k], r))
        assert r < 1e-3, (k, got[k], want[k])
    assert launches > 100, "libshotvae kernels did not run (%d launches)" % launches
    if verbose:
        print("smoke OK: %d libshotvae kernel launches, tcgen05 path built: %d" % (launches, _abi.lib.sv_has_tcgen05()))
