"""world_size-2 gloo tests (CPU) of the data-parallel logic, through the REAL GradReducer class (its CPU mode reduces
synchronously; streams are CUDA-only): construction broadcasts rank 0's state, the bucket ranges cover the arena
exactly once, and summing per-rank gradients then scaling by 1/world (what GradReducer + sv_sgd_step do) equals
the gradient of the global-batch mean loss."""
import importlib.util
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_ddp():
    """ddp.py by path: importing the package would load libshotvae.so, which this host-logic test does not need"""
    spec = importlib.util.spec_from_file_location("shotvae_ddp_under_test", os.path.join(ROOT, "shot-vae_b200", "shotvae_b200", "ddp.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class _FakeNet:
    """the flat arenas of plan.Net on the CPU"""

    def __init__(self, sizes, seed):
        self.poff, off = {}, 0
        for k, n in sizes.items():
            self.poff[k] = (off, n, (n,))
            off += n
        self.n_params = off
        g = torch.Generator().manual_seed(seed)
        self.params = torch.randn(off, generator=g)
        self.grads = torch.zeros(off)
        self.momentum = torch.randn(off, generator=g)
        self.running = torch.randn(7, generator=g)
        self.nbt = torch.full((3,), seed, dtype=torch.int64)
        self.param_epoch = 0


SIZES = {"feature_extractor.encoder.pre_process.conv0.weight": 37, "continuous_inference.mean.fc.weight": 11,
         "feature_reconstructor.decoder.0.weight": 101, "feature_reconstructor.decoder.3.weight": 53}


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ddp = _load_ddp()
    net = _FakeNet(SIZES, seed=10 + rank)          # ranks start DIFFERENT on purpose
    before = net.params.clone()
    red = ddp.GradReducer(net)                      # real constructor: asserts the layout, broadcasts rank 0's state
    ref = _FakeNet(SIZES, seed=10)
    ok = torch.equal(net.params, ref.params) and torch.equal(net.momentum, ref.momentum) and torch.equal(net.running, ref.running) \
        and torch.equal(net.nbt, ref.nbt) and net.param_epoch == 1 and (rank == 0 or not torch.equal(before, net.params))
    ok = ok and red.state_checksum() == 0.0
    covered = sorted(red.buckets.values())
    ok = ok and covered[0][0] == 0 and covered[-1][1] == net.n_params and covered[0][1] == covered[1][0] and red.world == world
    # per-rank shard gradient of the mean-over-local-batch loss
    torch.manual_seed(0)
    w = torch.randn(net.n_params, requires_grad=True)
    x = torch.randn(world * 4, net.n_params)
    shard = x[rank * 4:(rank + 1) * 4]
    (shard @ w).pow(2).mean().backward()
    net.grads.copy_(w.grad)
    red.bucket_ready("decoder")                     # the engine's call order
    red.bucket_ready("encoder")
    red.wait_all()
    got = net.grads / world
    w2 = w.detach().clone().requires_grad_(True)
    (x @ w2).pow(2).mean().backward()
    ok = ok and torch.allclose(got, w2.grad, atol=1e-5)
    # a layout where the decoder is not the arena tail must be refused
    bad = dict(list(SIZES.items())[2:] + list(SIZES.items())[:2])
    try:
        ddp.GradReducer(_FakeNet(bad, 1), broadcast=False)
        ok = False
    except AssertionError:
        pass
    out.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_grad_reducer_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = dict(q.get(timeout=5) for _ in range(2))
    assert res == {0: True, 1: True}
