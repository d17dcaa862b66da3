"""Drop-in `lib.criterion` (reference lib/criterion.py): same class names, constructor arguments and
return values; the reductions and their gradients are libshotvae kernels.

VAECriterion / ClsCriterion are the two criteria the training step calls (main_shot_vae.py:289-292,
316-318,358).  M1Criterion, M2Criterion, ReconstructionCriterion, KLNormCriterion and
KLDiscCriterion are never called by the reference; they keep their entry points here, built from
the same kernels where those apply."""
import torch
from torch import nn

from shotvae_b200 import _abi
from shotvae_b200._abi import lib, check, ptr

eps = 1e-7


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _abi.ShotVaeError("lib.criterion needs CUDA tensors; libshotvae has no CPU path")


def _layout(x_rec):
    """(tensor, is_nhwc): accepts NCHW-contiguous tensors and the NHWC-backed views the VAE returns"""
    if x_rec.is_contiguous():
        return x_rec, 0
    nhwc = x_rec.permute(0, 2, 3, 1)
    if nhwc.is_contiguous():
        return nhwc, 1
    return x_rec.contiguous(), 0


class _RecFn(torch.autograd.Function):
    """reconstruction term: forward computes the loss AND d loss / d x_rec in one pass"""

    @staticmethod
    def forward(ctx, x_rec, x, bce, x_sigma):
        _need_cuda(x_rec, x)
        B, ch = x.size(0), x.size(1)
        hw = x[0, 0].numel()
        xr, nhwc = _layout(x_rec.float())
        xc = x.contiguous().float()
        terms = torch.zeros(3, dtype=torch.float32, device=x.device)
        g = torch.empty_like(xr)
        check(lib.sv_elbo_rec_fwd_bwd(ptr(xc), ptr(xr), nhwc, B, ch, hw, 1 if bce else 0, float(x_sigma), None, ptr(terms),
                                      None, 0, ptr(g), _abi.stream()))
        ctx.g = g.permute(0, 3, 1, 2) if nhwc else g
        return terms[0]

    @staticmethod
    def backward(ctx, go):
        return ctx.g * go, None, None, None


class _KLFn(torch.autograd.Function):
    """Gaussian KL on z and categorical KL on y against the uniform prior (criterion.py:50-56)"""

    @staticmethod
    def forward(ctx, mu, ls, la):
        _need_cuda(mu, ls, la)
        B, D, nd = mu.size(0), mu.size(1), la.size(1)
        mu, ls, la = mu.contiguous().float(), ls.contiguous().float(), la.contiguous().float()
        terms = torch.zeros(3, dtype=torch.float32, device=mu.device)
        st = _abi.stream()
        check(lib.sv_elbo_kl_fwd(ptr(mu), ptr(ls), ptr(la), B, D, nd, ptr(terms), st))
        g_mu, g_ls, g_la = torch.empty_like(mu), torch.empty_like(ls), torch.empty_like(la)
        check(lib.sv_elbo_kl_bwd(ptr(mu), ptr(ls), ptr(la), None, None, 1, B, D, nd, ptr(g_mu), ptr(g_ls), ptr(g_la), 0, st))
        ctx.save_for_backward(g_mu, g_ls, g_la)
        return terms[1], terms[2]

    @staticmethod
    def backward(ctx, g_kc, g_kd):
        g_mu, g_ls, g_la = ctx.saved_tensors
        return (g_mu * g_kc if g_kc is not None else None, g_ls * g_kc if g_kc is not None else None,
                g_la * g_kd if g_kd is not None else None)


class _SoftCEFn(torch.autograd.Function):
    """-(1/B) sum_b w_b sum_c predict*label  (criterion.py:104-107)"""

    @staticmethod
    def forward(ctx, predict, label, batch_weight):
        _need_cuda(predict, label)
        B, nd = predict.shape
        pr = predict.contiguous().float()
        tgt = label.contiguous().float()
        if batch_weight is not None:
            tgt = tgt * batch_weight.view(B, -1).float()
        terms = torch.zeros(2, dtype=torch.float32, device=predict.device)
        g = torch.empty_like(pr)
        check(lib.sv_posterior_fwd_bwd(ptr(pr), ptr(tgt), None, None, None, None, None, None, None, None, B, 1, nd, ptr(terms),
                                       ptr(g), None, None, 0, _abi.stream()))
        ctx.save_for_backward(g)
        return terms[0]

    @staticmethod
    def backward(ctx, go):
        (g,) = ctx.saved_tensors
        return g * go, None, None


class VAECriterion(nn.Module):
    def __init__(self, discrete_dim=10, x_sigma=1, bce_reconstruction=True):
        super().__init__()
        self.x_sigma = x_sigma
        self.bce_reconstruction = bce_reconstruction
        self.discrete_dim = discrete_dim

    def forward(self, x, x_reconstructed, z_mean, z_log_sigma, disc_log_alpha):
        rec = _RecFn.apply(x_reconstructed, x, self.bce_reconstruction, self.x_sigma)
        klc, kld = _KLFn.apply(z_mean, z_log_sigma, disc_log_alpha)
        return rec, klc, kld


class ClsCriterion(nn.Module):
    def forward(self, predict, label, batch_weight=None):
        return _SoftCEFn.apply(predict, label, batch_weight)


class ReconstructionCriterion(nn.Module):
    def __init__(self, x_sigma=1, bce_reconstruction=True):
        super().__init__()
        self.x_sigma, self.bce_reconstruction = x_sigma, bce_reconstruction

    def forward(self, x, x_reconstructed):
        return _RecFn.apply(x_reconstructed, x, self.bce_reconstruction, self.x_sigma)


class M1Criterion(nn.Module):
    def __init__(self, x_sigma=1, bce_reconstruction=True):
        super().__init__()
        self.x_sigma, self.bce_reconstruction = x_sigma, bce_reconstruction

    def forward(self, x, x_reconstructed, M1_mean, M1_log_sigma):
        rec = _RecFn.apply(x_reconstructed, x, self.bce_reconstruction, self.x_sigma)
        dummy = torch.full((M1_mean.size(0), 2), -0.6931471805599453, device=M1_mean.device)
        klc, _ = _KLFn.apply(M1_mean, M1_log_sigma, dummy)
        return rec, klc


class M2Criterion(nn.Module):
    def __init__(self, discrete_dim=10):
        super().__init__()
        self.discrete_dim = discrete_dim

    def forward(self, M2_mean, M2_log_sigma, disc_log_alpha):
        return _KLFn.apply(M2_mean, M2_log_sigma, disc_log_alpha)


class _KLPairFn(torch.autograd.Function):
    """two-distribution KL forms (criterion.py:151-157,172-176): loss and every input gradient in one pass"""

    @staticmethod
    def forward(ctx, mode, *xs):
        xs = [x.contiguous().float() for x in xs]
        _need_cuda(*xs)
        batch = xs[0].size(0)
        loss = torch.zeros(1, dtype=torch.float32, device=xs[0].device)
        grads = [torch.empty_like(x) if ctx.needs_input_grad[i + 1] else None for i, x in enumerate(xs)]
        pad = [None] * (4 - len(xs))
        check(lib.sv_kl_pair_fwd_bwd(mode, *[ptr(x) for x in xs + pad], xs[0].numel(), batch, ptr(loss),
                                     *[ptr(g) for g in grads + pad], _abi.stream()))
        ctx.grads = grads
        return loss[0]

    @staticmethod
    def backward(ctx, go):
        return (None,) + tuple(None if g is None else g * go for g in ctx.grads)


class KLNormCriterion(nn.Module):
    """criterion.py:134-158 (never called by the reference): KL to the standard normal through the fused ELBO-KL
    kernel, KL between two diagonal Gaussians through sv_kl_pair_fwd_bwd."""

    def forward(self, z_mean_pre, z_log_sigma_pre, z_mean_gt=None, z_sigma_gt=None):
        if z_mean_gt is None or z_sigma_gt is None:
            dummy = torch.full((z_mean_pre.size(0), 2), -0.6931471805599453, device=z_mean_pre.device)
            return _KLFn.apply(z_mean_pre, z_log_sigma_pre, dummy)[0]
        return _KLPairFn.apply(0, z_mean_pre, z_log_sigma_pre, z_mean_gt, z_sigma_gt)


class KLDiscCriterion(nn.Module):
    """criterion.py:161-177 (never called by the reference): sum_j KL between predicted log-probabilities and target
    probabilities, in either order."""

    def forward(self, disc_log_pre, disc_gt, qp_order=True):
        return _KLPairFn.apply(1 if qp_order else 2, disc_log_pre, disc_gt)
