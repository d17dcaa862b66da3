"""Generates tests/golden/*.json by EXECUTING THE UNMODIFIED REFERENCE (/root/reference) in the build
container.  The reference is Python and cannot travel to the GPU box, so its outputs are committed
here as small fixtures: every loss term, the host RNG draws' scalars, both pairing indices, and
fixed-position samples + norms of every output tensor / parameter gradient / post-step state entry.

Run:  python tests/golden/make_golden.py            (needs /root/reference; CPU only, ~2 min)

How the reference is driven: `sys.argv` is patched before importing
`main_shot_vae` (argparse runs at import), `.cuda()` is patched to identity on this GPU-less host,
the loaders are plain lists, and the reference's own `train()` is called with recording wrappers
injected for `model`, `elbo_criterion`, `cls_criterion` and `optimizer` -- the step itself is the
reference's code, untouched.
"""
import json
import os
import sys
import importlib

import numpy as np
import torch

REF = os.environ.get("SHOTVAE_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
NSAMP = 12


def sample_positions(numel, n=NSAMP):
    """Deterministic sample positions shared with the tests."""
    if numel <= n:
        return list(range(numel))
    return [int((i * 2654435761 + 12345) % numel) for i in range(n)]


def summarize(t):
    t = t.detach().double().flatten()
    pos = sample_positions(t.numel())
    return {"numel": int(t.numel()), "l2": float(t.norm()), "sum": float(t.sum()),
            "samples": [float(t[p]) for p in pos]}


class Recorder:
    """Callable wrapper that records the outputs of every call to the wrapped reference object."""

    def __init__(self, inner, name, log):
        self.inner, self.name, self.log = inner, name, log

    def __call__(self, *a, **k):
        out = self.inner(*a, **k)
        self.log.append((self.name, out))
        return out

    def __getattr__(self, item):
        return getattr(self.inner, item)


class RecOptimizer:
    def __init__(self, inner, model, log):
        self.inner, self.model, self.log = inner, model, log

    def zero_grad(self):
        self.inner.zero_grad()

    def step(self):
        self.log.append(("grads", {k: p.grad.detach().clone() for k, p in self.model.named_parameters()}))
        self.inner.step()

    def __getattr__(self, item):
        return getattr(self.inner, item)


class StubWriter:
    def __init__(self): self.scalars = {}
    def add_scalar(self, tag, scalar_value, global_step=None): self.scalars[tag] = float(scalar_value)
    def add_image(self, **kw): pass


def import_reference(script, argv):
    sys.path.insert(0, REF)
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    sys.argv = argv
    return importlib.import_module(script)


def run_case(M, case):
    name, net, nd, batch, epoch, om, m2 = (case[k] for k in ("name", "net", "nd", "batch", "epoch", "om", "m2"))
    sys.path.insert(0, ROOT)
    from oracle import shotvae_oracle as O
    dataset = "Cifar100" if nd == 100 else "Cifar10"
    hyper = O.default_hyper(dataset, m2)
    hyper["om"] = om
    hyper["br"] = case.get("br", True)
    for k in ("akb", "aew", "apw", "ewm", "kbmc", "kbmd", "cmi", "dmi", "epochs"):
        setattr(M.args, k, hyper[k])
    if not m2:
        for k in ("pwm", "wrd", "wmf", "epsilon", "om"):
            setattr(M.args, k, hyper[k])
    M.args.print_freq = 10 ** 9
    M.args.reconstruct_freq = 10 ** 9 if epoch else 1
    torch.manual_seed(1)
    model = M.VariationalAutoEncoder(encoder_name=net, num_input_channels=3, drop_rate=0, img_size=(32, 32),
                                     data_parallel=False, continuous_latent_dim=128, disc_latent_dim=nd,
                                     sample_temperature=0.67, small_input=True)
    init_ok = all(torch.equal(v, O.init_state(net, nd)[k]) for k, v in model.state_dict().items())
    crit = M.VAECriterion(discrete_dim=nd, x_sigma=1, bce_reconstruction=hyper["br"])
    cls = M.ClsCriterion()
    opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9, weight_decay=5e-4)
    il, ll, iu, lu = O.synthetic_batch(batch, nd, case["data_seed"])
    # record the host draws the reference makes (wrap the RNG entry points it calls)
    draws = []
    o_randn, o_rand, o_perm, o_beta = torch.randn, torch.rand, torch.randperm, np.random.beta
    torch.randn = lambda *a, **k: (lambda t: (draws.append(("randn", t)), t)[1])(o_randn(*a, **k))
    torch.rand = lambda *a, **k: (lambda t: (draws.append(("rand", t)), t)[1])(o_rand(*a, **k))
    torch.randperm = lambda *a, **k: (lambda t: (draws.append(("randperm", t)), t)[1])(o_perm(*a, **k))
    np.random.beta = lambda *a, **k: (lambda v: (draws.append(("beta", float(v))), v)[1])(o_beta(*a, **k))
    log = []
    torch.manual_seed(case["rng_seed"]); np.random.seed(case["rng_seed"])
    try:
        writer = StubWriter()
        M.train([(iu, lu)], [(il, ll)], model=Recorder(model, "model", log), elbo_criterion=Recorder(crit, "elbo", log),
                cls_criterion=Recorder(cls, "cls", log), optimizer=RecOptimizer(opt, model, log), epoch=epoch,
                writer=writer, discrete_latent_dim=nd)
    finally:
        torch.randn, torch.rand, torch.randperm, np.random.beta = o_randn, o_rand, o_perm, o_beta
    g = {"case": case, "init_matches_oracle_init": bool(init_ok), "kl_inference": writer.scalars["Train/KL_Inference"]}
    g["draw_kinds"] = [k for k, _ in draws]
    g["betas"] = [v for k, v in draws if k == "beta"]
    g["randperms"] = [v.tolist() for k, v in draws if k == "randperm"]
    g["draw_summaries"] = [summarize(v) if k in ("randn", "rand") else None for k, v in draws]
    models = [o for n, o in log if n == "model"]
    g["model_outputs"] = [[summarize(t) for t in o[:4]] for o in models]
    g["elbo_terms"] = [[float(t) for t in o] for n, o in log if n == "elbo"]
    g["cls_terms"] = [float(o) for n, o in log if n == "cls"]
    grads = [o for n, o in log if n == "grads"][0]
    g["grads"] = {k: summarize(v) for k, v in grads.items()}
    g["post_state"] = {k: summarize(v.float()) for k, v in model.state_dict().items()}
    if om and not m2:
        # the pairing the reference's own loop (mixup.py:11-18) produces on the P3 latents
        mu, ls = models[2][1].detach(), models[2][2].detach()
        from lib.utils.mixup import gaussian_kl_divergence_calculation as gk
        b = mu.size(0)
        kl = torch.zeros(b, b)
        for i in range(b):
            for j in range(b):
                kl[i, j] = gk(mu[i], ls[i], mu[j], ls[j])
        idx = torch.topk(kl, 2, largest=False)[1][:, 1]
        g["om_index"] = idx.tolist()
        g["om_mu"] = mu.tolist()          # the full FP32 latents the pairing was computed on, so the CUDA
        g["om_ls"] = ls.tolist()          # kernel can be checked for bit-exact indices on the GPU box
        g["om_oracle_matrix_bitexact"] = bool(torch.equal(kl, O.pairwise_kl_matrix(mu, ls)))
        g["om_oracle_index_equal"] = bool(torch.equal(idx, O.optimal_match_index(mu, ls)))
        srt = torch.sort(kl, dim=1)[0]
        g["om_min_gap_2nd_3rd"] = float((srt[:, 2] - srt[:, 1]).min())
    return g


def eval_state(O, net, nd, seed):
    """constructor-initialised state with non-trivial BatchNorm running statistics (shared with the tests)"""
    st = O.init_state(net, nd)
    g = torch.Generator().manual_seed(seed)
    for k in st:
        if k.endswith("running_mean"):
            st[k] = 0.2 * torch.randn(st[k].shape, generator=g)
        elif k.endswith("running_var"):
            st[k] = 0.5 + torch.rand(st[k].shape, generator=g)
    return st


def run_eval_case(M, case):
    """the reference MODULE in eval mode (what its valid()/test() call, main_shot_vae.py:414-455): BatchNorm with
    running statistics, sampling unchanged"""
    sys.path.insert(0, ROOT)
    from oracle import shotvae_oracle as O
    net, nd, batch = case["net"], case["nd"], case["batch"]
    model = M.VariationalAutoEncoder(encoder_name=net, num_input_channels=3, drop_rate=0, img_size=(32, 32),
                                     data_parallel=False, continuous_latent_dim=128, disc_latent_dim=nd,
                                     sample_temperature=0.67, small_input=True)
    st = eval_state(O, net, nd, case["state_seed"])
    model.load_state_dict(st)
    model.eval()
    il, ll, iu, lu = O.synthetic_batch(batch, nd, case["data_seed"])
    g = dict(case=case, outputs={})
    with torch.no_grad():
        torch.manual_seed(case["rng_seed"])
        out = model(iu, disc_label=lu)                      # labelled form (one-hot y)
        g["outputs"]["with_label"] = [summarize(t) for t in out]
        torch.manual_seed(case["rng_seed"])
        out = model(iu)                                     # gumbel-softmax form
        g["outputs"]["gumbel"] = [summarize(t) for t in out]
    after = model.state_dict()
    g["running_stats_untouched"] = bool(all(torch.equal(after[k], st[k]) for k in st))
    return g


CASES_EVAL = [
    dict(name="eval_wrn28x2_nd10_b16", net="wideresnet-28-2", nd=10, batch=16, data_seed=31, rng_seed=12, state_seed=9),
    dict(name="eval_preact18_nd10_b8", net="preactresnet18", nd=10, batch=8, data_seed=32, rng_seed=13, state_seed=9),
]
CASES_SHOT = [
    dict(name="c2_wrn28x2_nd10_b16_e100", net="wideresnet-28-2", nd=10, batch=16, epoch=100, om=False, m2=False, data_seed=11, rng_seed=5),
    dict(name="c2_wrn28x2_nd10_b128_e0", net="wideresnet-28-2", nd=10, batch=128, epoch=0, om=False, m2=False, data_seed=12, rng_seed=6),
    dict(name="c2_wrn28x2_nd10_b32_e400_om", net="wideresnet-28-2", nd=10, batch=32, epoch=400, om=True, m2=False, data_seed=13, rng_seed=7),
    dict(name="c3_wrn28x2_nd100_b16_e100_mse", net="wideresnet-28-2", nd=100, batch=16, epoch=100, om=False, m2=False, br=False, data_seed=14, rng_seed=8),
    dict(name="wrn10x1_nd10_b8_e100", net="wideresnet-10-1", nd=10, batch=8, epoch=100, om=False, m2=False, data_seed=15, rng_seed=9),
    dict(name="preact18_shot_nd10_b8_e100", net="preactresnet18", nd=10, batch=8, epoch=100, om=False, m2=False, data_seed=16, rng_seed=10),
    dict(name="c2_wrn28x2_nd10_b16_e0_seed2", net="wideresnet-28-2", nd=10, batch=16, epoch=0, om=False, m2=False, data_seed=2, rng_seed=2),
    dict(name="c3_wrn28x2_nd100_b16_e400_seed3", net="wideresnet-28-2", nd=100, batch=16, epoch=400, om=False, m2=False, data_seed=3, rng_seed=3),
]
CASES_M2 = [
    dict(name="c5_m2_preact18_nd100_b16_e100", net="preactresnet18", nd=100, batch=16, epoch=100, om=False, m2=True, br=False, data_seed=21, rng_seed=3),
    dict(name="m2_wrn28x2_nd10_b16_e400", net="wideresnet-28-2", nd=10, batch=16, epoch=400, om=False, m2=True, data_seed=22, rng_seed=4),
]


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "shot"
    only = sys.argv[2] if len(sys.argv) > 2 else None      # optional: regenerate one case by name (read before argv is patched)
    if which in ("shot", "eval"):
        M = import_reference("main_shot_vae", ["main_shot_vae.py", "--gpu", "", "--dp", "--br", "-b", "128"])
        cases = CASES_SHOT
    else:
        M = import_reference("main_M2_vae", ["main_M2_vae.py", "--gpu", "", "--dp", "-b", "128",
                                             "--net-name", "preactresnet18", "--dataset", "Cifar100"])
        cases = CASES_M2
    if which == "eval" or (which == "shot" and not only):
        for case in CASES_EVAL:
            g = run_eval_case(M, case)
            with open(os.path.join(HERE, case["name"] + ".json"), "w") as f:
                json.dump(g, f)
            print(case["name"], "running stats untouched:", g["running_stats_untouched"])
    if which == "eval":
        return
    for case in cases:
        if only and case["name"] != only:
            continue
        g = run_case(M, case)
        with open(os.path.join(HERE, case["name"] + ".json"), "w") as f:
            json.dump(g, f)
        print(case["name"], "init==oracle_init:", g["init_matches_oracle_init"], "elbo:", g["elbo_terms"],
              {k: g[k] for k in g if k.startswith("om_o") or k.startswith("om_m")})


if __name__ == "__main__":
    # the two reference scripts both parse argv at import and define `args` globals; run them in
    # separate processes
    if len(sys.argv) == 1:
        import subprocess
        for w in ("shot", "m2"):
            subprocess.check_call([sys.executable, os.path.abspath(__file__), w])
    else:
        main()
