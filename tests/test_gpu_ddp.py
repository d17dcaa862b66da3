"""2-GPU data-parallel parity (SURVEY.md section 8 rows a14 / e): two ranks (one process per GPU, NCCL) run ONE fused step
on two different shards with teacher-fed host draws.  Asserted: (1) rank 1 starts from a deliberately different
initialisation and is brought in line by GradReducer's construction-time broadcast; (2) after the step both ranks
hold bit-identical parameters and momentum; (3) those parameters equal the oracle's step on the gradient averaged
over the two shards (the reference's global-batch mean under nn.DataParallel, main_shot_vae.py:324,364-365);
(4) BatchNorm running statistics stay per replica (DataParallel semantics: each replica normalises its own shard).

Needs two visible GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_gpu_ddp.py -m gpu` (skipped on one GPU; the
committed log of that run is profiles/r02_pytest_gpu_2gpu.log)."""
import os
import sys
import tempfile

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NET, ND, B, EPOCH = "wideresnet-10-1", 10, 16, 100


def _worker(rank, world, port, tmp):
    for p in (ROOT, os.path.join(ROOT, "shot-vae_b200"), os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    os.environ["NCCL_MAX_CTAS"] = "4"               # the reserved-SM path: backward grids of 144 CTAs (ddp.GradReducer.cta_limit)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from shotvae_b200.engine import TrainStep
    from shotvae_b200.ddp import GradReducer
    from test_gpu_step import build_model, _feed
    blob = torch.load(os.path.join(tmp, "inputs.pt"), weights_only=False)
    st = blob["state"]
    if rank == 1:                                   # a replica that starts WRONG: the broadcast must repair it
        st = {k: (v + 0.01 if v.dtype == torch.float32 else v) for k, v in st.items()}
    model = build_model(NET, ND, st).train()
    model._ensure_bound()
    red = GradReducer(model._net)
    spread0 = red.state_checksum()
    ts = TrainStep(model, B, hyper=blob["hyper"], use_graph=False, device_noise=False, reducer=red)
    ts.set_epoch(EPOCH)
    sh = blob["shards"][rank]
    terms = ts.step(sh["il"], sh["ll"], sh["iu"], sh["lu"], draws=_feed(ts, sh["log"], False))
    torch.cuda.synchronize()
    spread1 = red.state_checksum()
    torch.save(dict(state={k: v.cpu() for k, v in model.state_dict().items()}, momentum=model._net.momentum.cpu(), terms=terms,
                    spread0=spread0, spread1=spread1), os.path.join(tmp, "rank%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
def test_two_rank_step_equals_oracle_on_two_shards():
    from oracle import shotvae_oracle as O
    from test_gpu_step import grad_errors, rel
    import torch.multiprocessing as mp
    hyper = O.default_hyper("Cifar10")
    st = O.init_state(NET, ND)
    shards, runs = [], []
    for r in range(2):
        il, ll, iu, lu = O.synthetic_batch(B, ND, 30 + r)
        cs = O.clone_state(st)
        torch.manual_seed(50 + r); np.random.seed(50 + r)
        draws = O.LiveDraws()
        out = O.shot_step(cs, NET, ND, il, ll, iu, lu, EPOCH, hyper, draws)
        shards.append(dict(il=il, ll=ll, iu=iu, lu=lu, log=draws.log))
        runs.append((cs, out))
    # the data-parallel step: gradient = mean over the shards, one SGD step from the common state
    fin = O.clone_state(st)
    names = O.param_names(fin)
    for k in names:
        fin[k].requires_grad_(True)
        fin[k].grad = (runs[0][0][k].grad + runs[1][0][k].grad) / 2
    O.sgd_step(fin, {}, hyper["lr"], hyper["momentum"], hyper["wd"])
    with tempfile.TemporaryDirectory() as tmp:
        torch.save(dict(state=st, hyper={k: v for k, v in hyper.items() if k != "temperature"}, shards=shards), os.path.join(tmp, "inputs.pt"))
        port = 29600 + os.getpid() % 2000
        mp.spawn(_worker, args=(2, port, tmp), nprocs=2, join=True)
        res = [torch.load(os.path.join(tmp, "rank%d.pt" % r), weights_only=False) for r in range(2)]
    assert res[0]["spread0"] == 0.0 and res[1]["spread0"] == 0.0, "construction-time broadcast did not align the replicas"
    assert res[0]["spread1"] == 0.0
    for k in names:
        assert torch.equal(res[0]["state"][k], res[1]["state"][k]), "ranks diverged after one step: " + k
    assert torch.equal(res[0]["momentum"], res[1]["momentum"])
    upd = {k: res[0]["state"][k].float() - st[k].float() for k in names}
    wupd = {k: fin[k].detach().float() - st[k].float() for k in names}
    errs = grad_errors(upd, wupd)
    assert errs["decoder"] < 0.15 and errs["heads"] < 0.15 and errs["encoder"] < 0.65, errs
    # a single-shard update is clearly different from the two-shard one (so the test is sensitive to a missing all-reduce)
    one = O.clone_state(st)
    for k in names:
        one[k].requires_grad_(True)
        one[k].grad = runs[0][0][k].grad.clone()
    O.sgd_step(one, {}, hyper["lr"], hyper["momentum"], hyper["wd"])
    errs_one = grad_errors(upd, {k: one[k].detach().float() - st[k].float() for k in names})
    assert errs_one["decoder"] > 1.5 * errs["decoder"], (errs_one, errs)
    for r in range(2):
        cs, out = runs[r]
        for k in ("rec_l", "klc_l", "rec_u", "klc_u"):
            assert abs(res[r]["terms"][k] - out[k]) < 1e-3 * abs(out[k]), (r, k)
        rs = max(rel(res[r]["state"][k], cs[k]) for k in cs if k.endswith("running_mean") or k.endswith("running_var"))
        assert rs < 3e-2, (r, rs)                   # per-replica BatchNorm statistics
    assert not torch.equal(res[0]["state"]["feature_extractor.encoder.transition.norm.running_mean"],
                           res[1]["state"]["feature_extractor.encoder.transition.norm.running_mean"])
