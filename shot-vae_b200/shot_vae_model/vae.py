"""Drop-in `shot_vae_model.vae` (reference vae.py:10-151).

`VariationalAutoEncoder(...)` takes the reference's constructor arguments, owns parameters under the
reference's state_dict keys, and `forward(input_img, mixup=False, disc_label=None,
disc_pseudo_label=None, mixup_lam=None)` returns `(reconstruction, norm_mean, norm_log_sigma,
disc_log_alpha)` as autograd-connected CUDA tensors.  Everything between the arguments and the
results is libshotvae: one `torch.autograd.Function` whose forward/backward launch the sm_100a
kernels through the C ABI.  There is no eager/PyTorch fallback -- CPU tensors raise.
"""
import os
from collections import OrderedDict

import numpy as np
import torch
from torch import nn

from .wideresnet import get_wide_resnet
from .preactresnet import get_preact_resnet
from .decoder import Decoder
from shotvae_b200 import _abi
from shotvae_b200._abi import lib, check, ptr
from shotvae_b200.plan import Net, Ctx, pad16


class _Inference(nn.Sequential):
    def __init__(self, num_input_channels, latent_dim, disc_variable=True):
        super().__init__()
        self.add_module('fc', nn.Linear(num_input_channels, latent_dim))
        if disc_variable:
            self.add_module('log_softmax', nn.LogSoftmax(dim=1))


class Sample(nn.Module):
    """Holder for the sampling temperature; the sampling itself is sv_sample_fwd.  Noise is drawn on
    the host generator in the reference's order (vae.py:69,82-84) unless `device_noise` is set."""

    def __init__(self, temperature):
        super().__init__()
        self._temperature = temperature


class _VAEFunction(torch.autograd.Function):
    @staticmethod
    def forward(fctx, anchor, model, x, mixup, disc_label, disc_pseudo_label, mixup_lam, grad_enabled):
        net = model._net
        B = x.size(0)
        ctx = model._acquire_ctx(B)
        ctx.reset()
        model._pack_if_needed()
        st = _abi.stream()
        x = x.contiguous().float()
        cp = pad16(net.in_ch)
        x_img = ctx.t("x_img", (B, 32, 32, cp))
        check(net.fn("sv_pack_image")(ptr(x), ptr(x_img), B, net.in_ch, 32 * 32, cp, st))
        # model.eval(): BatchNorm uses the running statistics (reference valid()/test(), main_shot_vae.py:414-455);
        # everything else, including the sampling, is what the reference does in both modes
        net.eval_bn = not model.training
        if net.eval_bn and grad_enabled and anchor.requires_grad:
            net.eval_bn = False
            model._release_ctx(ctx)
            raise NotImplementedError("backward through an eval-mode forward is not implemented: wrap it in torch.no_grad() "
                                      "as the reference's valid()/test() do")
        try:
            return _VAEFunction._forward_body(fctx, anchor, model, x, x_img, ctx, mixup, disc_label, disc_pseudo_label, mixup_lam,
                                              grad_enabled)
        finally:
            net.eval_bn = False

    @staticmethod
    def _forward_body(fctx, anchor, model, x, x_img, ctx, mixup, disc_label, disc_pseudo_label, mixup_lam, grad_enabled):
        net = model._net
        B, st = x.size(0), _abi.stream()
        feat = net.encoder_fwd(ctx, x_img)
        mu, ls, la = net.heads_fwd(ctx, feat)
        # host noise, reference order: randn for z first, then rand for the gumbel sample
        dev = x.device
        eps = model._draw_normal((B, net.ldc), dev)
        if disc_label is not None:
            label = disc_label.to(dev, torch.int64).contiguous()
            if mixup:
                lam_dev = torch.tensor([float(mixup_lam), float(1 - mixup_lam)], dtype=torch.float32).to(dev)
                label_mix = disc_pseudo_label.to(dev, torch.int64).contiguous()
                lat = net.sample_fwd(ctx, 0, 1, eps, label=label, label_mix=label_mix, lam_dev=lam_dev)
                ctx.keep = (label, label_mix, lam_dev)
            else:
                lat = net.sample_fwd(ctx, 0, 0, eps, label=label)
                ctx.keep = (label,)
        else:
            unif = model._draw_uniform((B, net.nd), dev)
            lat = net.sample_fwd(ctx, 0, 2, eps, unif=unif)
            ctx.keep = (unif,)
        rec_nhwc = net.decoder_fwd(ctx, lat)
        rec = torch.empty(B, net.in_ch, 32, 32, dtype=torch.float32, device=dev)
        check(lib.sv_nhwc_to_nchw_f32(ptr(rec_nhwc), ptr(rec), B, net.in_ch, 32 * 32, st))
        if model.training:
            net.bn_running_update([(ctx, 0)])
        fctx.model, fctx.pctx = model, ctx
        fctx.set_materialize_grads(False)
        needs_bwd = grad_enabled and anchor.requires_grad   # (grad mode is always off inside Function.forward)
        outs = (rec, mu.clone(), ls.clone(), la.clone())
        if not needs_bwd:
            model._release_ctx(ctx)
            fctx.pctx = None
        return outs

    @staticmethod
    def backward(fctx, g_rec, g_mu, g_ls, g_la):
        model, ctx = fctx.model, fctx.pctx
        if ctx is None:
            raise RuntimeError("backward through a VAE forward that ran under no_grad")
        net = model._net
        B, st = ctx.B, _abi.stream()
        model._attach_grads()
        gm = ctx.t("g.mu", (B, net.ldc), torch.float32)
        gl = ctx.t("g.ls", (B, net.ldc), torch.float32)
        ga = ctx.t("g.la", (B, net.nd), torch.float32)
        for dst, src in ((gm, g_mu), (gl, g_ls), (ga, g_la)):
            if src is None:
                dst.zero_()
            else:
                dst.copy_(src)
        if g_rec is not None:
            cp = pad16(net.in_ch)
            g_img = ctx.t("g.rec", (B, 32, 32, cp))
            g_rec = g_rec.contiguous().float()      # (held in a variable: ptr() of a temporary would dangle)
            check(net.fn("sv_pack_image")(ptr(g_rec), ptr(g_img), B, net.in_ch, 32 * 32, cp, st))
            g_lat = net.decoder_bwd(ctx, g_img)
            net.sample_bwd(ctx, 0, g_lat, gm, gl, ga, accumulate=1)
        g_feat = net.heads_bwd(ctx, gm, gl, ga)
        net.encoder_bwd(ctx, g_feat)
        model._release_ctx(ctx)
        fctx.pctx = None
        return (None,) * 8


class VariationalAutoEncoder(nn.Module):
    def __init__(self, encoder_name, num_input_channels=1, drop_rate=0, img_size=(160, 160), data_parallel=True,
                 continuous_latent_dim=100, disc_latent_dim=10, sample_temperature=0.67, small_input=False):
        super().__init__()
        if "wideresnet" in encoder_name:
            self.feature_extractor = get_wide_resnet(encoder_name, drop_rate, input_channels=num_input_channels,
                                                     small_input=small_input, data_parallel=data_parallel)
        elif "preactresnet" in encoder_name:
            self.feature_extractor = get_preact_resnet(encoder_name, drop_rate, input_channels=num_input_channels,
                                                       small_input=small_input, data_parallel=data_parallel)
        else:
            # densenet is outside the hot path
            raise NotImplementedError("{} not implemented".format(encoder_name))
        if tuple(img_size) != (32, 32):
            raise NotImplementedError("libshotvae implements 32x32 inputs (img_size=%s)" % (img_size,))
        self.global_avg = nn.AdaptiveAvgPool2d(output_size=(1, 1))
        feat = self.feature_extractor.num_feature_channel
        self.continuous_inference = nn.Sequential()
        self.continuous_inference.add_module("mean", _Inference(feat, continuous_latent_dim, disc_variable=False))
        self.continuous_inference.add_module("log_sigma", _Inference(feat, continuous_latent_dim, disc_variable=False))
        self._disc_latent_dim = disc_latent_dim
        self.disc_latent_inference = _Inference(feat, disc_latent_dim, disc_variable=True)
        self.sample = Sample(temperature=sample_temperature)
        ksz = tuple(int(s / 32) for s in img_size)
        self.feature_reconstructor = Decoder(num_channel=num_input_channels,
                                             latent_dim=int(continuous_latent_dim + np.sum(disc_latent_dim)),
                                             data_parallel=data_parallel, kernel_size=ksz)
        self._cfg = dict(encoder_name=self.feature_extractor.plan_name, nd=int(disc_latent_dim), ldc=int(continuous_latent_dim),
                         in_ch=int(num_input_channels), temperature=float(sample_temperature))
        self._net = None
        self._ctx_pool = {}
        self._pack_version = None
        # "bf16" (production: bf16 tensor-core operands) or "fp32" (parity-grade mode: fp32 tensors and FP32 kernels, ~20x
        # slower; same launch sequence).  Takes effect at the next forward (the arena is re-bound).
        self.precision = os.environ.get("SHOTVAE_PRECISION", "bf16")
        self.device_noise = False      # True: draw eps/u on the device generator (benchmark mode)
        self.noise_source = None       # optional object with randn(*shape) / rand(*shape) (parity replays)

    # ---- reference checkpoints may carry nn.DataParallel's ".module." infix ---------------------
    def load_state_dict(self, state_dict, strict=True, **kw):
        cleaned = OrderedDict((k.replace(".module.", "."), v) for k, v in state_dict.items())
        res = super().load_state_dict(cleaned, strict=strict, **kw)
        self._pack_version = None           # the masters changed: repack before the next forward
        return res

    # ---- arena binding ---------------------------------------------------------------------------
    def _bind(self):
        params = OrderedDict(self.named_parameters())
        bufs = OrderedDict(self.named_buffers())
        first = next(iter(params.values()))
        if not first.is_cuda:
            raise _abi.ShotVaeError("VariationalAutoEncoder must be on a CUDA device (call .cuda()); no CPU path exists")
        net = Net(named_params=params, named_buffers=bufs, device=first.device, precision=self.precision, **self._cfg)
        for k, p in params.items():
            p.data = net.p(k)
            p.grad = None
        for k, b in bufs.items():
            b.data = net.b(k)
        self._net, self._ctx_pool, self._pack_version = net, {}, None
        self.__dict__["_first_name"], self.__dict__["_first"] = next(iter(params.items()))
        self.__dict__["_plist"] = list(params.values())

    def _ensure_bound(self):
        if self._net is None or self._first.data_ptr() != self._net.p(self._first_name).data_ptr() or \
                self._net.precision != self.precision:
            self._bind()

    def _pack_if_needed(self):
        """Re-derive the bf16 operand copies of the conv / convT weights whenever the FP32 masters may have moved.
        The Parameters are views of the arena with their OWN version counters (optimizer.step(), p.add_(),
        load_state_dict() bump p._version, never the arena's), and the fused sv_sgd_step kernel bumps nothing, so
        the signal is (sum of the parameters' versions, the net's explicit param_epoch that TrainStep advances)."""
        v = (sum(p._version for p in self._plist), self._net.param_epoch)
        if v != self._pack_version:
            self._net.pack_weights()
            self._pack_version = v

    def _attach_grads(self):
        """parameter .grad tensors are views of the flat gradient arena; if the optimizer dropped them
        (zero_grad(set_to_none=True)) the arena is cleared and re-attached."""
        net = self._net
        if self._first.grad is None or self._first.grad.data_ptr() != net.g(self._first_name).data_ptr():
            net.zero_grads()
            for k, p in self.named_parameters():
                p.grad = net.g(k)

    def _acquire_ctx(self, B):
        pool = self._ctx_pool.setdefault(B, [])
        return pool.pop() if pool else Ctx(self._net, 1, B)

    def _release_ctx(self, ctx):
        self._ctx_pool.setdefault(ctx.B, []).append(ctx)

    def _draw_normal(self, shape, dev):
        if self.noise_source is not None:
            return self.noise_source.randn(*shape).to(dev)
        if self.device_noise:
            return torch.randn(shape, device=dev)
        return torch.randn(shape).to(dev)

    def _draw_uniform(self, shape, dev):
        if self.noise_source is not None:
            return self.noise_source.rand(*shape).to(dev)
        if self.device_noise:
            return torch.rand(shape, device=dev)
        return torch.rand(shape).to(dev)

    def forward(self, input_img, mixup=False, disc_label=None, disc_pseudo_label=None, mixup_lam=None):
        if not input_img.is_cuda:
            raise _abi.ShotVaeError("input must be a CUDA tensor; libshotvae has no CPU path")
        self._ensure_bound()
        return _VAEFunction.apply(self._first, self, input_img, mixup, disc_label, disc_pseudo_label, mixup_lam,
                                  torch.is_grad_enabled())
