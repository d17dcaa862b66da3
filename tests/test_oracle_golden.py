"""Pins the CPU oracle (oracle/shotvae_oracle.py) against outputs of the UNMODIFIED reference recorded
by tests/golden/make_golden.py.  Both sides are torch FP32 on the CPU, so the tolerance is tight
(1e-5 relative): they issue the same ATen ops in the same order."""
import glob
import json
import os

import numpy as np
import pytest
import torch

from oracle import shotvae_oracle as O
from tests.golden.make_golden import sample_positions

_ALL = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.json")))
GOLD = [p for p in _ALL if not os.path.basename(p).startswith(("eval_", "aux_"))]      # aux_*: tests/test_aux_cpu.py
GOLD_EVAL = [p for p in _ALL if os.path.basename(p).startswith("eval_")]
RTOL = 1e-5


def close(a, b, rtol=RTOL, atol=1e-7):
    return abs(a - b) <= atol + rtol * max(abs(a), abs(b))


def check_summary(t, g, what, rtol=RTOL):
    t = t.detach().double().flatten()
    assert t.numel() == g["numel"], what
    assert close(float(t.norm()), g["l2"], rtol), (what, float(t.norm()), g["l2"])
    scale = g["l2"] / max(1.0, g["numel"] ** 0.5)
    for p, v in zip(sample_positions(t.numel()), g["samples"]):
        assert abs(float(t[p]) - v) <= rtol * max(abs(v), scale) + 1e-9, (what, p, float(t[p]), v)


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-5] for p in GOLD])
def test_oracle_matches_reference_golden(path):
    g = json.load(open(path))
    c = g["case"]
    assert g["init_matches_oracle_init"]
    dataset = "Cifar100" if c["nd"] == 100 else "Cifar10"
    hyper = O.default_hyper(dataset, c["m2"])
    hyper["om"], hyper["br"] = c["om"], c.get("br", True)
    st = O.init_state(c["net"], c["nd"])
    il, ll, iu, lu = O.synthetic_batch(c["batch"], c["nd"], c["data_seed"])
    torch.manual_seed(c["rng_seed"]); np.random.seed(c["rng_seed"])
    draws = O.LiveDraws()
    step = O.m2_step if c["m2"] else O.shot_step
    out = step(st, c["net"], c["nd"], il, ll, iu, lu, c["epoch"], hyper, draws, keep=True)
    # host draws: same kinds in the same order, same values
    assert [k for k, _ in draws.log] == g["draw_kinds"]
    assert [v for k, v in draws.log if k == "beta"] == g["betas"]
    assert [v.tolist() for k, v in draws.log if k == "randperm"] == g["randperms"]
    # loss terms
    e = g["elbo_terms"]
    for got, want in zip((out["rec_l"], out["klc_l"], out["kld_l"]), e[0]):
        assert close(got, want), (got, want)
    for got, want in zip((out["rec_u"], out["klc_u"], out["kld_u"]), e[1]):
        assert close(got, want), (got, want)
    assert close(out["kl_inference"], g["kl_inference"])
    T = out["tensors"]
    passes = [("rec_l", "mu_l", "ls_l", "la_l"), ("rec_u", "mu_u", "ls_u", "la_u")] if c["m2"] else \
        [("rec_l", "mu_l", "ls_l", "la_l"), ("rec2", "mu2", "ls2", "la2"), ("rec_u", "mu_u", "ls_u", "la_u"),
         ("rec4", "mu4", "ls4", "la4")]
    for names, gp in zip(passes, g["model_outputs"]):
        for n, gs in zip(names, gp):
            check_summary(T[n], gs, n)
    if c["om"]:
        assert T["idx_u"].tolist() == g["om_index"]
        assert g["om_oracle_matrix_bitexact"] and g["om_oracle_index_equal"]
    # parameter gradients and post-SGD state (rtol 1e-4: summation order inside autograd may differ)
    for k, gs in g["grads"].items():
        check_summary(st[k].grad, gs, "grad " + k, rtol=1e-4)
    O.sgd_step(st, {}, lr=0.1)
    for k, gs in g["post_state"].items():
        check_summary(st[k].float(), gs, "state " + k, rtol=1e-4)


@pytest.mark.parametrize("path", GOLD_EVAL, ids=[os.path.basename(p)[:-5] for p in GOLD_EVAL])
def test_oracle_eval_forward_matches_reference_module(path):
    """model.eval() forward of the reference module (BatchNorm on running statistics; what valid()/test() run,
    main_shot_vae.py:414-455) against the oracle's training=False path"""
    from tests.golden.make_golden import eval_state
    g = json.load(open(path))
    c = g["case"]
    assert g["running_stats_untouched"]
    st = eval_state(O, c["net"], c["nd"], c["state_seed"])
    il, ll, iu, lu = O.synthetic_batch(c["batch"], c["nd"], c["data_seed"])
    topo = O.encoder_topology(c["net"])
    with torch.no_grad():
        torch.manual_seed(c["rng_seed"])
        out = O.vae_forward(st, topo, iu, O.LiveDraws(), 0.67, disc_label=lu, training=False)
        for t, gs, n in zip(out, g["outputs"]["with_label"], ("rec", "mu", "ls", "la")):
            check_summary(t, gs, "eval/with_label/" + n)
        torch.manual_seed(c["rng_seed"])
        out = O.vae_forward(st, topo, iu, O.LiveDraws(), 0.67, training=False)
        for t, gs, n in zip(out, g["outputs"]["gumbel"], ("rec", "mu", "ls", "la")):
            check_summary(t, gs, "eval/gumbel/" + n)
