"""Generates tests/golden/aux_f3_f4.json by executing the UNMODIFIED reference's never-called criteria
(lib/criterion.py:59-91,111-177), vectorised distance helpers (lib/utils/calculate_dist.py) and SSL samplers
(lib/dataloader.py:73-193), plus torchvision's own transform pipeline for the augmentation arithmetic, on seeded
inputs in the build container.  Run: python tests/golden/make_golden_aux.py   (needs /root/reference, torchvision; CPU)"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SHOTVAE_REFERENCE", "/root/reference")
sys.path.insert(0, HERE)
from make_golden import summarize          # noqa: E402


def aux_inputs():
    """seeded inputs shared with the tests"""
    g = torch.Generator().manual_seed(77)
    B, D, nd = 12, 128, 10
    r = lambda *s: torch.randn(*s, generator=g)
    d = dict(x=torch.rand(B, 3, 32, 32, generator=g), x_rec=r(B, 3, 32, 32) * 2, mu=r(B, D), ls=r(B, D) * 0.3,
             mu_gt=r(B, D), sigma_gt=torch.rand(B, D, generator=g) + 0.2,
             la=torch.log_softmax(r(B, nd), 1), p_gt=torch.softmax(r(B, nd), 1),
             u1=r(9, D), ls1=r(9, D) * 0.3, u2=r(7, D), ls2=r(7, D) * 0.3)
    return d, nd


def aug_inputs():
    rng = np.random.RandomState(5)
    data = rng.randint(0, 256, size=(6, 32, 32, 3), dtype=np.uint8)
    index = [4, 0, 5, 2, 2]
    params = [[0, 0, 0], [8, 8, 1], [3, 5, 1], [4, 4, 0], [7, 1, 1]]
    mnist = rng.randint(0, 256, size=(3, 28, 28, 1), dtype=np.uint8)
    return data, index, params, mnist


def sha(t):
    return hashlib.sha256(np.ascontiguousarray(t.numpy() if torch.is_tensor(t) else t).tobytes()).hexdigest()


def main():
    sys.path.insert(0, REF)
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    from lib import criterion as RC
    from lib.utils import calculate_dist as RD
    from lib import dataloader as RL
    d, nd = aux_inputs()
    out = {}
    f = lambda t: float(t)
    out["m1_bce"] = [f(v) for v in RC.M1Criterion(1, True)(d["x"], d["x_rec"], d["mu"], d["ls"])]
    out["m1_mse"] = [f(v) for v in RC.M1Criterion(0.5, False)(d["x"], d["x_rec"], d["mu"], d["ls"])]
    out["m2"] = [f(v) for v in RC.M2Criterion(nd)(d["mu"], d["ls"], d["la"])]
    out["rec_bce"] = f(RC.ReconstructionCriterion(1, True)(d["x"], d["x_rec"]))
    out["rec_mse"] = f(RC.ReconstructionCriterion(2.0, False)(d["x"], d["x_rec"]))
    out["klnorm_prior"] = f(RC.KLNormCriterion()(d["mu"], d["ls"]))
    out["klnorm_pair"] = f(RC.KLNormCriterion()(d["mu"], d["ls"], d["mu_gt"], d["sigma_gt"]))
    out["kldisc_qp"] = f(RC.KLDiscCriterion()(d["la"], d["p_gt"], True))
    out["kldisc_pq"] = f(RC.KLDiscCriterion()(d["la"], d["p_gt"], False))
    # gradients of the two-distribution forms (autograd through the reference's own forward)
    xs = [d[k].clone().requires_grad_(True) for k in ("mu", "ls", "mu_gt", "sigma_gt")]
    RC.KLNormCriterion()(*xs).backward()
    out["klnorm_pair_grads"] = [summarize(x.grad) for x in xs]
    for name, order in (("kldisc_qp_grads", True), ("kldisc_pq_grads", False)):
        ys = [d[k].clone().requires_grad_(True) for k in ("la", "p_gt")]
        RC.KLDiscCriterion()(ys[0], ys[1], order).backward()
        out[name] = [summarize(y.grad) for y in ys]
    out["dist_kl"] = summarize(RD.pairwise_norm_kl_dist_gpu(d["u1"], d["ls1"], d["u2"], d["ls2"]))
    out["dist_euclid"] = summarize(RD.pairwise_square_euclidean_gpu(d["u1"], d["u2"]))
    out["dist_wasserstein"] = summarize(RD.pairwise_norm_wasserstein_dist_gpu(d["u1"], d["ls1"], d["u2"], d["ls2"]))
    out["dist_cosine_numpy"] = summarize(torch.from_numpy(RD.calculate_mean_dist_pairwise(d["u1"].numpy(), d["u2"].numpy(), False, "cosine")))
    out["dist_kl_vec_numpy"] = summarize(torch.from_numpy(RD.gaussian_kl_calculation_vec(d["u1"].numpy(), d["ls1"].numpy())))
    # samplers: the reference's own functions on synthetic labels
    labels = torch.randint(0, 10, (600,), generator=torch.Generator().manual_seed(3), dtype=torch.int32)
    torch.manual_seed(11)
    sv, sl, su = RL.get_cifar10_ssl_sampler(labels, 5, 8, 10)
    out["ssl_cifar10"] = dict(valid=list(sv.indices), train_l=list(sl.indices), train_u_sha=sha(np.asarray(su.indices, dtype=np.int64)),
                              train_u_len=len(su.indices))
    torch.manual_seed(12)
    sv, st = RL.get_cifar10_sl_sampler(labels, 4, 10)
    out["sl_cifar10"] = dict(valid=list(sv.indices), train_sha=sha(np.asarray(st.indices, dtype=np.int64)), train_len=len(st.indices))
    # augmentation arithmetic: torchvision's own ops in the order of the reference's Compose (dataloader.py:63-66)
    import torchvision.transforms.functional as TF
    from PIL import Image
    data, index, params, mnist = aug_inputs()
    imgs = []
    for i, (ci, cj, flip) in zip(index, params):
        im = TF.pad(Image.fromarray(data[i]), 4, padding_mode="reflect")
        if flip:
            im = TF.hflip(im)
        imgs.append(TF.to_tensor(TF.crop(im, ci, cj, 32, 32)))
    out["augment_train_sha"] = sha(torch.stack(imgs))
    out["augment_test_sha"] = sha(torch.stack([TF.to_tensor(Image.fromarray(data[i])) for i in index]))
    mn = []
    for k, (ci, cj) in enumerate([(0, 0), (4, 4), (2, 3)]):
        mn.append(TF.to_tensor(TF.crop(TF.pad(Image.fromarray(mnist[k, :, :, 0]), 4, padding_mode="reflect"), ci, cj, 32, 32)))
    out["augment_mnist_sha"] = sha(torch.stack(mn))
    json.dump(out, open(os.path.join(HERE, "aux_f3_f4.json"), "w"), indent=1)
    print("wrote aux_f3_f4.json")


if __name__ == "__main__":
    main()
