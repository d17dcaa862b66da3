"""Times the UNMODIFIED reference (baseline/_ref, a verbatim copy of /root/reference that travels to the GPU box)
through its own public entry point -- `main_shot_vae.train` / `main_M2_vae.train` on list loaders -- either on the host
cores (bench.py's `--impl reference` arm and `cpu_baseline`) or on one GPU through torch's eager CUDA path
(`gpu_reference`: cuDNN / cuBLAS kernels, TF32 convolutions = what `python main_shot_vae.py` does on a B200 today).
None of libshotvae is imported here.  Prints ONE JSON line.

How the reference is made importable without touching it: sys.argv is set before the import
(argparse runs at import), `--gpu ""` plus identity `.cuda()` shims on a CPU run (the reference hard-codes .cuda()),
loaders are plain lists of (image, label) batches and the TensorBoard writer is a stub."""
import argparse
import importlib
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")

CONFIGS = {          # same table as bench.py
    "c2": dict(net="wideresnet-28-2", nd=10, dataset="Cifar10", m2=False, br=True),
    "c3": dict(net="wideresnet-28-2", nd=100, dataset="Cifar100", m2=False, br=True),
    "c4": dict(net="wideresnet-28-10", nd=10, dataset="Cifar10", m2=False, br=True),
    "c5": dict(net="preactresnet18", nd=100, dataset="Cifar100", m2=True, br=False),
}


class _Writer:
    def add_scalar(self, *a, **k):
        pass

    def add_image(self, *a, **k):
        pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--config", default="c2")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--epoch", type=int, default=100)
    ap.add_argument("--gpu-index", default="0")
    a = ap.parse_args()
    if not os.path.exists(os.path.join(REF, "main_shot_vae.py")):
        print(json.dumps({"unavailable": "baseline/_ref (copy of the reference) is not present"}))
        return
    cfg = CONFIGS[a.config]
    script = "main_M2_vae" if cfg["m2"] else "main_shot_vae"
    cuda = a.device == "cuda"
    sys.path.insert(0, REF)
    sys.argv = [script + ".py", "--dp", "--gpu", a.gpu_index if cuda else "", "-b", str(a.batch), "--net-name", cfg["net"],
                "--dataset", cfg["dataset"]] + (["--br"] if cfg["br"] else [])
    import torch
    if not cuda:
        torch.Tensor.cuda = lambda self, *x, **k: self
        torch.nn.Module.cuda = lambda self, *x, **k: self
    M = importlib.import_module(script)
    import numpy as np
    nd = cfg["nd"]
    # what main() sets per dataset before it calls train() (main_shot_vae.py:139,161-163; main_M2_vae.py:123,146-147)
    if cfg["dataset"] == "Cifar100":
        M.args.dmi = 4.6
        if cfg["m2"]:
            M.args.cmi = 1280
        else:
            M.args.akb, M.args.apw = 150, 400
    else:
        M.args.dmi = 2.3
        if cfg["m2"]:
            M.args.cmi = 200
    M.args.print_freq = 10 ** 9
    M.args.reconstruct_freq = 10 ** 9
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(1)
    model = M.VariationalAutoEncoder(encoder_name=cfg["net"], num_input_channels=3, drop_rate=0, img_size=(32, 32), data_parallel=False,
                                     continuous_latent_dim=128, disc_latent_dim=nd, sample_temperature=0.67, small_input=True)
    model = model.cuda()
    crit = M.VAECriterion(discrete_dim=nd, x_sigma=1, bce_reconstruction=cfg["br"]).cuda()
    cls = M.ClsCriterion()
    opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9, weight_decay=5e-4)
    g = torch.Generator().manual_seed(1234)
    pool = [((torch.rand(a.batch, 3, 32, 32, generator=g), torch.randint(0, nd, (a.batch,), generator=g)),
             (torch.rand(a.batch, 3, 32, 32, generator=g), torch.randint(0, nd, (a.batch,), generator=g))) for _ in range(4)]
    if cuda:
        pool = [tuple((x.pin_memory(), y.pin_memory()) for x, y in pair) for pair in pool]
    np.random.seed(100)

    def run(n):
        lu = [pool[i % len(pool)][1] for i in range(n)]
        ll = [pool[i % len(pool)][0] for i in range(n)]
        M.train(lu, ll, model=model, elbo_criterion=crit, cls_criterion=cls, optimizer=opt, epoch=a.epoch, writer=_Writer(),
                discrete_latent_dim=nd)
        if cuda:
            torch.cuda.synchronize()

    if a.warmup > 0:
        run(a.warmup)
    t0 = time.perf_counter()
    run(a.steps)
    dt = time.perf_counter() - t0
    out = dict(device=a.device, config=a.config, steps=a.steps, warmup=a.warmup, seconds=dt, ms_per_step=1e3 * dt / a.steps,
               images_per_s=a.batch * a.steps / dt, cores=threads, torch=torch.__version__, entry=script + ".train")
    if cuda:
        out.update(gpu=torch.cuda.get_device_name(0), cudnn=torch.backends.cudnn.version(), cudnn_allow_tf32=bool(torch.backends.cudnn.allow_tf32),
                   precision="FP32 storage, TF32 cuDNN convolutions (torch default), FP32 matmul")
    print(json.dumps(out))


if __name__ == "__main__":
    main()
