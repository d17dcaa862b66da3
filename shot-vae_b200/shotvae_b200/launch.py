"""Multi-process launcher: replaces the reference's `--gpu "0,1"` + nn.DataParallel (main_shot_vae.py:105-107,191-195) by one
process per GPU (torch.distributed over NCCL / NVLink).

    python -m shotvae_b200.launch --gpus 8 train_script.py [script args]

and, inside the script:

    rank, world, local = init_distributed()            # NCCL process group, cuda device = local rank
    model = VariationalAutoEncoder(...).cuda(); model._ensure_bound()
    trainer = Trainer(model, 128, reducer=GradReducer(model._net) if world > 1 else None)

Each rank feeds its own (labelled, unlabelled) loader shard; GradReducer broadcasts rank 0's state at construction."""
import argparse
import os
import subprocess
import sys


def build_command(gpus, script, script_args, port=29500):
    if gpus <= 1:
        return [sys.executable, script] + list(script_args)
    return [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(gpus), "--master-addr", "127.0.0.1",
            "--master-port", str(port), script] + list(script_args)


def init_distributed():
    """-> (rank, world_size, local_rank); initialises NCCL when launched with WORLD_SIZE > 1"""
    import torch
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
            # NCCL on at most 4 SMs, the backward's persistent grids on the others (ddp.GradReducer.cta_limit)
            os.environ.setdefault("NCCL_MAX_CTAS", "4")
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        torch.cuda.manual_seed(1234 + rank)          # per-rank device noise (eps / gumbel draws)
    return rank, world, local


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--port", type=int, default=29500)
    ap.add_argument("script")
    ap.add_argument("script_args", nargs=argparse.REMAINDER)
    a = ap.parse_args(argv)
    return subprocess.call(build_command(a.gpus, a.script, a.script_args, a.port))


if __name__ == "__main__":
    sys.exit(main())
