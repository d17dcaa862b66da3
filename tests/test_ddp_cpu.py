"""world_size-2 gloo test (CPU) of the data-parallel logic: bucket ranges cover the arena exactly once,
and summing per-rank gradients then scaling by 1/world (what GradReducer + sv_sgd_step do) equals the
gradient of the global-batch mean loss."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


class _FakeNet:
    def __init__(self, sizes):
        self.poff, off = {}, 0
        for k, n in sizes.items():
            self.poff[k] = (off, n, (n,))
            off += n
        self.n_params = off
        self.grads = torch.zeros(off)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "shot-vae_b200"))
    from shotvae_b200 import ddp
    # GradReducer's stream plumbing is CUDA-only; exercise its bucket arithmetic and the collective on CPU
    net = _FakeNet({"feature_extractor.encoder.pre_process.conv0.weight": 37, "continuous_inference.mean.fc.weight": 11,
                    "feature_reconstructor.decoder.0.weight": 101, "feature_reconstructor.decoder.3.weight": 53})
    red = ddp.GradReducer.__new__(ddp.GradReducer)
    red.net, red.group, red.world = net, None, world
    split = net.poff["feature_reconstructor.decoder.0.weight"][0]
    red.buckets = {"encoder": (0, split), "decoder": (split, net.n_params)}
    covered = sorted(red.buckets.values())
    assert covered[0][0] == 0 and covered[-1][1] == net.n_params and covered[0][1] == covered[1][0]
    # per-rank shard gradient of mean-over-local-batch loss
    torch.manual_seed(0)
    w = torch.randn(net.n_params, requires_grad=True)
    x = torch.randn(world * 4, net.n_params)
    shard = x[rank * 4:(rank + 1) * 4]
    loss = (shard @ w).pow(2).mean()
    loss.backward()
    net.grads.copy_(w.grad)
    for name in ("decoder", "encoder"):
        s, e = red.buckets[name]
        dist.all_reduce(net.grads[s:e], op=dist.ReduceOp.SUM)
    got = net.grads / world
    w2 = w.detach().clone().requires_grad_(True)
    (x @ w2).pow(2).mean().backward()
    ok = torch.allclose(got, w2.grad, atol=1e-5)
    if rank == 0:
        out.put(bool(ok))
    dist.destroy_process_group()


def test_bucketed_allreduce_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
