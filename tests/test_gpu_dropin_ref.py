"""The drop-in boundary exercised by the reference itself: the UNMODIFIED `main_shot_vae.train` / `main_M2_vae.train`
(baseline/_ref, a verbatim copy of the reference that travels to the GPU box) run two optimizer steps on list loaders
with `shot-vae_b200/` providing every package they import, and the logged `Train/KL_Inference` plus the post-training
state are compared with the CPU oracle driven by the same seeds (reference: main_shot_vae.py:261-383,
main_M2_vae.py:242-323).  Skipped when baseline/_ref is absent."""
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest
import torch

from test_gpu_step import grad_errors, rel

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "main_shot_vae.py")), reason="baseline/_ref (reference copy) not present")
@pytest.mark.parametrize("script,net,nd,dataset", [("main_shot_vae", "wideresnet-28-2", 10, "Cifar10"),
                                                   ("main_M2_vae", "preactresnet18", 100, "Cifar100")])
def test_unmodified_reference_train_runs_on_the_dropin(script, net, nd, dataset):
    from oracle import shotvae_oracle as O
    m2 = script == "main_M2_vae"
    B, epoch, seed, nsteps = 16, 100, 5, 2
    hyper = O.default_hyper(dataset, m2)
    hyper["br"] = not m2
    st = O.init_state(net, nd)
    batches = []
    for i in range(nsteps):
        il, ll, iu, lu = O.synthetic_batch(B, nd, 40 + i)
        batches.append(dict(il=il, ll=ll, iu=iu, lu=lu))
    # oracle: the same two steps, one continuous host RNG stream (the reference seeds once, then draws as it goes)
    ost = O.clone_state(st)
    torch.manual_seed(seed); np.random.seed(seed)
    mom, kls = {}, []
    for b in batches:
        out = (O.m2_step if m2 else O.shot_step)(ost, net, nd, b["il"], b["ll"], b["iu"], b["lu"], epoch, hyper, O.LiveDraws())
        O.sgd_step(ost, mom, hyper["lr"], hyper["momentum"], hyper["wd"])
        kls.append(out["kl_inference"])
    argv = ["--dp", "--gpu", os.environ.get("CUDA_VISIBLE_DEVICES", "0").split(",")[0], "-b", str(B), "--net-name", net,
            "--dataset", dataset, "-bp", "/tmp"]
    if not m2:
        argv.append("--br")
    args = {k: hyper[k] for k in ("akb", "aew", "ewm", "kbmc", "kbmd", "cmi", "dmi", "epochs")}
    if not m2:
        args.update({k: hyper[k] for k in ("apw", "pwm", "wrd", "wmf", "epsilon", "om")})
    with tempfile.TemporaryDirectory() as tmp:
        torch.save(dict(argv=argv, args=args, net=net, nd=nd, br=hyper["br"], state=st, batches=batches, rng_seed=seed, epoch=epoch),
                   os.path.join(tmp, "in.pt"))
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dropin_ref_runner.py"), script, os.path.join(tmp, "in.pt"),
                            os.path.join(tmp, "out.pt")], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
        got = torch.load(os.path.join(tmp, "out.pt"), weights_only=False)
    assert got["launches"] > 300, "the reference's train() did not run on libshotvae kernels"
    want_kl = float(np.mean(kls))                                   # AverageMeter over the epoch's steps (main_shot_vae.py:339,376)
    assert abs(got["scalars"]["Train/KL_Inference"] - want_kl) < 5e-3 * abs(want_kl), (got["scalars"], want_kl)
    names = O.param_names(ost)
    errs = grad_errors({k: got["state"][k].float() - st[k].float() for k in names},
                       {k: ost[k].detach().float() - st[k].float() for k in names})
    assert errs["decoder"] < 0.2 and errs["heads"] < 0.2 and errs["encoder"] < 0.7, errs
    rs = max(rel(got["state"][k], ost[k]) for k in ost if k.endswith("running_mean") or k.endswith("running_var"))
    assert rs < 5e-2, rs
    nb = (2 if m2 else 4) * nsteps
    assert all(int(got["state"][k]) == nb for k in ost if k.endswith("num_batches_tracked") and "feature_extractor" in k)
