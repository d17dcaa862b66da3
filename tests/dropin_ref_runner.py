"""Runs the UNMODIFIED reference training script's own train() (baseline/_ref/main_shot_vae.py:261-383 or
main_M2_vae.py:242-323) on top of the drop-in packages: `shot-vae_b200/` provides `shot_vae_model`, `lib.criterion`,
`lib.utils.mixup`, `lib.utils.avgmeter` and `lib.dataloader`, and the reference directory itself is NOT on sys.path, so
every import of the script resolves to libshotvae.  Executed as a subprocess by tests/test_gpu_dropin_ref.py (argparse
runs at import of the reference script and it sets CUDA_VISIBLE_DEVICES before importing torch).

usage: python tests/dropin_ref_runner.py <main_shot_vae|main_M2_vae> <inputs.pt> <outputs.pt>"""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "shot-vae_b200")
REF = os.path.join(ROOT, "baseline", "_ref")


class StubWriter:
    def __init__(self):
        self.scalars = {}

    def add_scalar(self, tag, scalar_value, global_step=None):
        self.scalars[tag] = float(scalar_value)

    def add_image(self, *a, **k):
        pass


def main():
    script, inp, outp = sys.argv[1:4]
    import torch as _t                      # (already imported by the parent environment's sitecustomize or not: harmless)
    blob = _t.load(inp, weights_only=False)
    sys.argv = [script + ".py"] + blob["argv"]
    sys.path[:0] = [PKG]
    assert not any(os.path.abspath(p) == REF for p in sys.path if p)
    spec = importlib.util.spec_from_file_location("reference_" + script, os.path.join(REF, script + ".py"))
    M = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(M)              # argparse + CUDA_VISIBLE_DEVICES happen here, as in `python main_shot_vae.py ...`
    import numpy as np
    import torch
    import shot_vae_model.vae
    import lib.criterion
    assert os.path.abspath(shot_vae_model.vae.__file__).startswith(PKG) and os.path.abspath(lib.criterion.__file__).startswith(PKG)
    for k, v in blob["args"].items():       # per-dataset overrides main() would apply (main_shot_vae.py:139,161-163)
        setattr(M.args, k, v)
    M.args.print_freq = 10 ** 9
    M.args.reconstruct_freq = 10 ** 9
    nd = blob["nd"]
    model = M.VariationalAutoEncoder(encoder_name=blob["net"], num_input_channels=3, drop_rate=0, img_size=(32, 32),
                                     data_parallel=False, continuous_latent_dim=128, disc_latent_dim=nd,
                                     sample_temperature=0.67, small_input=True)
    model.load_state_dict(blob["state"])
    model = model.cuda()
    crit = M.VAECriterion(discrete_dim=nd, x_sigma=1, bce_reconstruction=blob["br"]).cuda()
    cls = M.ClsCriterion()
    opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9, weight_decay=5e-4)
    writer = StubWriter()
    loader_u = [(b["iu"], b["lu"]) for b in blob["batches"]]
    loader_l = [(b["il"], b["ll"]) for b in blob["batches"]]
    torch.manual_seed(blob["rng_seed"]); np.random.seed(blob["rng_seed"])
    from shotvae_b200 import _abi
    n0 = _abi.launch_count()
    M.train(loader_u, loader_l, model=model, elbo_criterion=crit, cls_criterion=cls, optimizer=opt, epoch=blob["epoch"],
            writer=writer, discrete_latent_dim=nd)
    torch.cuda.synchronize()
    torch.save(dict(scalars=writer.scalars, state={k: v.cpu() for k, v in model.state_dict().items()},
                    launches=_abi.launch_count() - n0), outp)


if __name__ == "__main__":
    main()
