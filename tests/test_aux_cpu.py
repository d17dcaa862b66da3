"""CPU tests for the rows beside the training step: the oracle's restatements of the
reference's never-called criteria, vectorised distances, SSL samplers and the augmentation arithmetic are pinned
against tests/golden/aux_f3_f4.json, recorded by executing the unmodified reference / torchvision
(tests/golden/make_golden_aux.py).  Host logic of lib.dataloader (sampler draws, loader length) is checked too."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from oracle import shotvae_oracle as O
from tests.golden.make_golden_aux import aux_inputs, aug_inputs
from tests.test_oracle_golden import check_summary, close

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "aux_f3_f4.json")))


def sha(t):
    return hashlib.sha256(np.ascontiguousarray(t.numpy() if torch.is_tensor(t) else t).tobytes()).hexdigest()


def test_unused_criteria_match_reference_golden():
    d, nd = aux_inputs()
    for got, want in zip(O.m1_criterion(d["x"], d["x_rec"], d["mu"], d["ls"], 1, True), GOLD["m1_bce"]):
        assert close(float(got), want)
    for got, want in zip(O.m1_criterion(d["x"], d["x_rec"], d["mu"], d["ls"], 0.5, False), GOLD["m1_mse"]):
        assert close(float(got), want)
    for got, want in zip(O.m2_criterion(d["mu"], d["ls"], d["la"], nd), GOLD["m2"]):
        assert close(float(got), want)
    assert close(float(O.reconstruction_criterion(d["x"], d["x_rec"], 1, True)), GOLD["rec_bce"])
    assert close(float(O.reconstruction_criterion(d["x"], d["x_rec"], 2.0, False)), GOLD["rec_mse"])
    assert close(float(O.kl_norm_criterion(d["mu"], d["ls"])), GOLD["klnorm_prior"])
    assert close(float(O.kl_norm_criterion(d["mu"], d["ls"], d["mu_gt"], d["sigma_gt"])), GOLD["klnorm_pair"])
    assert close(float(O.kl_disc_criterion(d["la"], d["p_gt"], True)), GOLD["kldisc_qp"])
    assert close(float(O.kl_disc_criterion(d["la"], d["p_gt"], False)), GOLD["kldisc_pq"])
    xs = [d[k].clone().requires_grad_(True) for k in ("mu", "ls", "mu_gt", "sigma_gt")]
    O.kl_norm_criterion(*xs).backward()
    for x, g in zip(xs, GOLD["klnorm_pair_grads"]):
        check_summary(x.grad, g, "klnorm grad")
    for name, order in (("kldisc_qp_grads", True), ("kldisc_pq_grads", False)):
        ys = [d[k].clone().requires_grad_(True) for k in ("la", "p_gt")]
        O.kl_disc_criterion(ys[0], ys[1], order).backward()
        for y, g in zip(ys, GOLD[name]):
            check_summary(y.grad, g, name)


def test_distance_helpers_match_reference_golden():
    d, _ = aux_inputs()
    check_summary(O.pairwise_norm_kl_dist(d["u1"], d["ls1"], d["u2"], d["ls2"]), GOLD["dist_kl"], "kl")
    check_summary(O.pairwise_square_euclidean(d["u1"], d["u2"]), GOLD["dist_euclid"], "euclid")
    check_summary(O.pairwise_norm_wasserstein_dist(d["u1"], d["ls1"], d["u2"], d["ls2"]), GOLD["dist_wasserstein"], "wd")
    check_summary(O.mean_dist_pairwise(d["u1"], d["u2"], "cosine"), GOLD["dist_cosine_numpy"], "cosine")
    check_summary(O.pairwise_norm_kl_dist(d["u1"], d["ls1"], d["u1"], d["ls1"]), GOLD["dist_kl_vec_numpy"], "kl vec", rtol=1e-4)
    # the --om metric (mixup.py:93-99) is the same quantity: cross-check the two statements
    kl = O.pairwise_kl_matrix(d["u1"], d["ls1"])
    assert torch.allclose(kl, O.pairwise_norm_kl_dist(d["u1"], d["ls1"], d["u1"], d["ls1"]), rtol=1e-4, atol=1e-4)


def test_ssl_samplers_match_reference_golden():
    labels = torch.randint(0, 10, (600,), generator=torch.Generator().manual_seed(3), dtype=torch.int32)
    torch.manual_seed(11)
    valid, lab, unl = O.per_class_split(labels, 10, 5, 8)
    g = GOLD["ssl_cifar10"]
    assert valid == g["valid"] and lab == g["train_l"] and len(unl) == g["train_u_len"]
    assert sha(np.asarray(unl, dtype=np.int64)) == g["train_u_sha"]
    assert set(lab) <= set(unl) and not (set(valid) & set(unl))          # unlabelled includes labelled, excludes validation
    torch.manual_seed(12)
    valid, train = O.per_class_split(labels, 10, 4)
    g = GOLD["sl_cifar10"]
    assert valid == g["valid"] and len(train) == g["train_len"] and sha(np.asarray(train, dtype=np.int64)) == g["train_sha"]


def test_augmentation_oracle_matches_torchvision_golden():
    data, index, params, mnist = aug_inputs()
    assert sha(O.augment_batch(data, index, params)) == GOLD["augment_train_sha"]          # bit-exact
    assert sha(O.augment_batch(data, index, None)) == GOLD["augment_test_sha"]
    mn = O.augment_batch(mnist, [0, 1, 2], [[0, 0, 0], [4, 4, 0], [2, 3, 0]], pad=4, out_size=32)
    assert sha(mn) == GOLD["augment_mnist_sha"]


def _load_dataloader_module():
    """lib/dataloader.py by path (its kernel binding is imported lazily, so the host logic runs without the GPU library)"""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("shotvae_dataloader_under_test", os.path.join(root, "shot-vae_b200", "lib", "dataloader.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_dropin_samplers_reproduce_reference_draws():
    DL = _load_dataloader_module()
    labels = torch.randint(0, 10, (600,), generator=torch.Generator().manual_seed(3), dtype=torch.int32)
    g = GOLD["ssl_cifar10"]
    for fn in (DL.get_cifar10_ssl_sampler, DL.get_ssl_sampler):
        torch.manual_seed(11)
        sv, sl, su = fn(labels, 5, 8, 10)
        assert list(sv.indices) == g["valid"] and list(sl.indices) == g["train_l"]
        assert sha(np.asarray(su.indices, dtype=np.int64)) == g["train_u_sha"]
    torch.manual_seed(12)
    sv, st = DL.get_cifar10_sl_sampler(labels, 4, 10)
    assert list(sv.indices) == GOLD["sl_cifar10"]["valid"] and sha(np.asarray(st.indices, dtype=np.int64)) == GOLD["sl_cifar10"]["train_sha"]
    torch.manual_seed(11)
    labels100 = torch.randint(0, 100, (3000,), generator=torch.Generator().manual_seed(4), dtype=torch.int32)
    sv, sl, su = DL.get_cifar100_ssl_sampler(labels100, 2, 3)
    torch.manual_seed(11)
    assert (list(sv.indices), list(sl.indices), list(su.indices)) == O.per_class_split(labels100, 100, 2, 3)
