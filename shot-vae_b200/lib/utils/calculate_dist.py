"""Drop-in `lib.utils.calculate_dist` (reference lib/utils/calculate_dist.py): pairwise distances between batches
of diagonal Gaussians / vectors.  The reference never imports this module; it is the
authors' own vectorised statement of the `--om` pairing metric.  The torch-tensor (GPU) entry points are one
libshotvae kernel each (sv_pairwise_dist) instead of n1 x n2 x d broadcast temporaries; the NumPy entry points of
the reference take host arrays and stay host-side NumPy exactly as there (they are not device code)."""
import numpy as np
import torch

from shotvae_b200 import _abi
from shotvae_b200._abi import lib, check, ptr

_KL, _EUCLID, _WASSERSTEIN, _COSINE = 0, 1, 2, 3


def _pairwise(mode, u1, u2, ls1=None, ls2=None):
    ts = [t for t in (u1, u2, ls1, ls2) if t is not None]
    for t in ts:
        if not t.is_cuda:
            raise _abi.ShotVaeError("lib.utils.calculate_dist needs CUDA tensors; libshotvae has no CPU path")
    u1, u2 = u1.detach().contiguous().float(), u2.detach().contiguous().float()
    ls1 = None if ls1 is None else ls1.detach().contiguous().float()
    ls2 = None if ls2 is None else ls2.detach().contiguous().float()
    n1, d = u1.shape
    n2 = u2.size(0)
    assert u2.size(1) == d
    out = torch.empty(n1, n2, dtype=torch.float32, device=u1.device)
    check(lib.sv_pairwise_dist(ptr(u1), ptr(ls1), ptr(u2), ptr(ls2), n1, n2, d, mode, ptr(out), _abi.stream()))
    return out


def pairwise_norm_kl_dist_gpu(u1, log_sigma1, u2, log_sigma2):
    """[i, j] = KL(N(u1[i], exp(log_sigma1[i])) || N(u2[j], exp(log_sigma2[j])))   (:94-107)"""
    return _pairwise(_KL, u1, u2, log_sigma1, log_sigma2)


def pairwise_square_euclidean_gpu(v1, v2):
    """[i, j] = ||v1[i] - v2[j]||^2   (:110-117)"""
    return _pairwise(_EUCLID, v1, v2)


def pairwise_norm_wasserstein_dist_gpu(u1, log_sigma1, u2, log_sigma2):
    """[i, j] = ||u1[i] - u2[j]||^2 + ||exp(ls1[i]) - exp(ls2[j])||^2   (:120-130)"""
    return _pairwise(_WASSERSTEIN, u1, u2, log_sigma1, log_sigma2)


def _dev(x):
    return (torch.from_numpy(np.ascontiguousarray(x)).float() if isinstance(x, np.ndarray) else x).cuda()


def gaussian_kl_calculation_vec(u, log_sigma, GPU_flag=False):
    """n x n matrix, [i, j] = KL(i || j)   (:35-58)"""
    if GPU_flag:
        u, log_sigma = _dev(u), _dev(log_sigma)
        return _pairwise(_KL, u, u, log_sigma, log_sigma)
    return gaussian_kl_calculation_vec_pairwise(u, log_sigma, u, log_sigma)


def gaussian_kl_calculation_vec_pairwise(u1, log_sigma1, u2, log_sigma2, GPU_flag=False):
    """n1 x n2 matrix, [i, j] = KL((u1[i], ls1[i]) || (u2[j], ls2[j]))   (:61-91)"""
    if GPU_flag:
        return _pairwise(_KL, _dev(u1), _dev(u2), _dev(log_sigma1), _dev(log_sigma2))
    v1, v2 = np.exp(log_sigma1) ** 2, np.exp(log_sigma2) ** 2
    ratio = v1[:, None, :] / v2[None, :, :]
    shift = (u1[:, None, :] - u2[None, :, :]) ** 2 / v2[None, :, :]
    return 0.5 * (-np.log(ratio).sum(2) + ratio.sum(2) + shift.sum(2) - v1.shape[1])


def calculate_mean_dist_pairwise(u1, u2, GPU_flag=False, distance="euclidean"):
    """(:133-160); "cosine" divides by the SQUARED norms, as the reference writes it"""
    if distance not in ("euclidean", "cosine"):
        raise NotImplementedError("distance {} not implemented".format(distance))
    if GPU_flag:
        return _pairwise(_EUCLID if distance == "euclidean" else _COSINE, _dev(u1), _dev(u2))
    if distance == "euclidean":
        return ((u1[:, None, :] - u2[None, :, :]) ** 2).sum(2)
    return u1.dot(u2.T) / ((u1 ** 2).sum(1)[:, None] * (u2 ** 2).sum(1)[None, :])


def gaussian_wd_calculation(u1, u2, log_sigma1, log_sigma2, diagflag=True):
    """Wasserstein-2 distance of one pair of diagonal Gaussians, host NumPy   (:5-10)"""
    if not diagflag:
        raise NotImplementedError("No diag covariance matrix not implemented")
    return np.sum((u1 - u2) ** 2) + np.sum((np.exp(log_sigma1) - np.exp(log_sigma2)) ** 2)


def gaussian_kl_calculation(u1, u2, log_sigma1, log_sigma2, diag1flag=True, diag2flag=True):
    """KL of one pair of diagonal Gaussians, host NumPy   (:13-32)"""
    if not (diag1flag and diag2flag):
        raise ValueError("Undefined value for diag1flag and diag2flag")
    v1, v2 = np.exp(log_sigma1) ** 2, np.exp(log_sigma2) ** 2
    return 0.5 * (np.sum(np.log(v2)) - np.sum(np.log(v1)) + np.sum(v1 / v2) + np.sum((u1 - u2) ** 2 / v2) - u1.shape[0])
