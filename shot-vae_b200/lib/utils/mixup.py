"""Drop-in `lib.utils.mixup` (reference lib/utils/mixup.py:5-41,93-99): optimal-interpolation mixup and
label smoothing.  lambda and the random pairing are drawn on the host exactly where the reference
draws them (np.random.beta, torch.randperm); the gather + interpolation of images / latents and the
`--om` pairwise-KL pairing are libshotvae kernels (the reference's O(B^2) Python loop with one
device->host sync per pair becomes one launch)."""
import numpy as np
import torch

from shotvae_b200 import _abi
from shotvae_b200._abi import lib, check, ptr


def _need_cuda(*ts):
    for t in ts:
        if not t.is_cuda:
            raise _abi.ShotVaeError("lib.utils.mixup needs CUDA tensors; libshotvae has no CPU path")


def optimal_match_index(z_mean, z_log_sigma, return_matrix=False):
    """index[i] = second entry of the ascending top-2 of row i of KL(N_i || N_j) (mixup.py:11-18)"""
    _need_cuda(z_mean, z_log_sigma)
    B, D = z_mean.shape
    mu, ls = z_mean.detach().contiguous().float(), z_log_sigma.detach().contiguous().float()
    index = torch.empty(B, dtype=torch.int64, device=mu.device)
    kl = torch.empty(B, B, dtype=torch.float32, device=mu.device) if return_matrix else None
    check(lib.sv_pairwise_kl_second_nearest(ptr(mu), ptr(ls), B, D, ptr(index), ptr(kl), _abi.stream()))
    return (index, kl) if return_matrix else index


def _lerp(image, z_mean, z_log_sigma, disc_log_alpha, index, lam):
    _need_cuda(image, z_mean, z_log_sigma, disc_log_alpha)
    B = image.size(0)
    img = image.detach().contiguous().float()
    mu, ls, la = (t.detach().contiguous().float() for t in (z_mean, z_log_sigma, disc_log_alpha))
    D, nd = mu.size(1), la.size(1)
    ch, hw = img.size(1), img[0, 0].numel()
    dev = img.device
    lam_dev = torch.tensor([float(lam), float(1 - lam)], dtype=torch.float32).to(dev)
    index = index.to(dev, torch.int64).contiguous()
    m_img, m_mu, m_sig, m_alpha = torch.empty_like(img), torch.empty_like(mu), torch.empty_like(mu), torch.empty_like(la)
    check(lib.sv_mixup_lerp(ptr(img), ptr(mu), ptr(ls), ptr(la), ptr(index), ptr(lam_dev), B, ch, hw, D, nd, ptr(m_img), None, 0,
                            ptr(m_mu), ptr(m_sig), ptr(m_alpha), _abi.stream()))
    return m_img, m_mu, m_sig, m_alpha, index


def mixup_vae_data(image, z_mean, z_log_sigma, disc_log_alpha, optimal_match=False):
    '''Returns mixed inputs, pairs of targets, and lambda'''
    lam = np.random.beta(2.0, 2.0)
    batch_size = image.size()[0]
    if optimal_match:
        index = optimal_match_index(z_mean, z_log_sigma)
    else:
        index = torch.randperm(batch_size).cuda()
    mixed_image, mixed_z_mean, mixed_z_sigma, mixed_disc_alpha, _ = _lerp(image, z_mean, z_log_sigma, disc_log_alpha, index, lam)
    return mixed_image, mixed_z_mean, mixed_z_sigma, mixed_disc_alpha, lam


def label_smoothing(image, z_mean, z_log_sigma, disc_log_alpha, epsilon=0.1, disc_label=None):
    if epsilon > 0:
        lam = np.random.beta(epsilon, epsilon)
    else:
        lam = 1
    batch_size = image.size()[0]
    index = torch.randperm(batch_size).cuda()
    s_image, s_mean, s_sigma, s_alpha, index = _lerp(image, z_mean, z_log_sigma, disc_log_alpha, index, lam)
    smoothed_disc_label = disc_label[index]
    return s_image, s_mean, s_sigma, s_alpha, smoothed_disc_label, lam


def gaussian_kl_divergence_calculation(z_mean_1, z_log_sigma_1, z_mean_2, z_log_sigma_2):
    """KL(N_1 || N_2) of one pair (mixup.py:93-99), through the pairwise kernel on a 2-row batch."""
    mu = torch.stack([z_mean_1, z_mean_2]).float()
    ls = torch.stack([z_log_sigma_1, z_log_sigma_2]).float()
    _, kl = optimal_match_index(mu, ls, return_matrix=True)
    return kl[0, 1]
