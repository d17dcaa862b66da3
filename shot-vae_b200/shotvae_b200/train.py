"""Epoch driver on the fused engine: what `main()` / `train()` of main_shot_vae.py (:202-258, :261-383) and
main_M2_vae.py do around the step -- learning-rate warm-up and milestones, the per-epoch loss schedules, the `ewm x 5`
rule, the (labelled, unlabelled) loader pairing with its short last batches, the epoch's KL_Inference average, and
checkpoints in the reference's format {'epoch', 'args', 'state_dict', 'optimizer'} (loadable both ways).

Batches of the common size run through `TrainStep` (one CUDA graph per batch size; a second graph is captured for the
tail size).  SHOT-VAE also accepts B_l != B_u (main_shot_vae.py:280-282 only asserts nothing): those steps run the same
kernels through the drop-in modules' autograd path (one pass group per forward, any batch) followed by the same fused
SGD kernel on the shared parameter / momentum arenas.  M2 truncates both batches to the common size
(main_M2_vae.py:259-266)."""
import bisect
import itertools
import os

import numpy as np
import torch
import torch.nn.functional as F

from . import _abi
from ._abi import lib, check, ptr
from .engine import TrainStep, default_hyper, alpha_schedule


class Trainer:
    def __init__(self, model, batch, hyper=None, dataset="Cifar10", m2=False, adjust_lr=(400, 500, 550), annotated_ratio=0.1,
                 reducer=None, use_graph=True, device_noise=True):
        self.model, self.batch, self.m2, self.dataset = model, int(batch), bool(m2), dataset
        self.h = dict(default_hyper(dataset, m2) if hyper is None else hyper)
        self.adjust_lr, self.annotated_ratio = tuple(adjust_lr), float(annotated_ratio)
        self.reducer, self.use_graph, self.device_noise = reducer, use_graph, device_noise
        self.base_lr, self.base_ewm = float(self.h["lr"]), float(self.h["ewm"])
        self.steps = {}              # batch size -> TrainStep (they share the model's parameter / gradient / momentum arenas)
        self.start_epoch = 0
        self._opt_state = None       # optimizer state waiting for the first TrainStep (resume)
        self._crit = None
        model._ensure_bound()

    # ---- schedules (host) ---------------------------------------------------------------------------------------
    def lr_at(self, epoch):
        """epoch 0 runs at 0.2 * lr (warm-up, :223-225); MultiStepLR(milestones, gamma 0.1) stepped with `epoch` AFTER each
        epoch (:252), so the rate drops from the epoch after a milestone on"""
        if epoch == 0:
            return 0.2 * self.base_lr
        return self.base_lr * (0.1 ** bisect.bisect_right(self.adjust_lr, epoch - 1))

    def ewm_at(self, epoch):
        """args.ewm *= 5 once epoch == adjust_lr[0] has been trained, Cifar10 with annotated_ratio >= 0.05 only (:255-258)"""
        if (not self.m2) and self.dataset == "Cifar10" and self.annotated_ratio >= 0.05 and epoch > self.adjust_lr[0]:
            return 5.0 * self.base_ewm
        return self.base_ewm

    # ---- steps ---------------------------------------------------------------------------------------------------
    def _step_for(self, B):
        ts = self.steps.get(B)
        if ts is None:
            ts = TrainStep(self.model, B, hyper=self.h, m2=self.m2, use_graph=self.use_graph, device_noise=self.device_noise,
                           reducer=self.reducer)
            first = next(iter(self.steps.values()), None)
            if first is not None:            # one optimizer: every TrainStep shares the arenas, so it shares the step count too
                ts._steps_done = first._steps_done
            elif self._opt_state is not None:
                ts.load_state_dict(self._opt_state)
                self._opt_state = None
            self.steps[B] = ts
        return ts

    def _sync_optimizer(self, ts, epoch):
        n = max(t._steps_done for t in self.steps.values())
        for t in self.steps.values():
            t._steps_done = n
        ts.h["ewm"] = self.ewm_at(epoch)
        ts.set_epoch(epoch)
        ts.set_lr(self.lr_at(epoch))

    def _ragged_step(self, image_l, label_l, image_u, label_u, epoch, ts_any):
        """B_l != B_u: the loop body of main_shot_vae.train (:284-364) on the drop-in modules, then the fused SGD kernel"""
        from lib.criterion import VAECriterion, ClsCriterion
        from lib.utils.mixup import mixup_vae_data, label_smoothing
        model, h, net = self.model, self.h, self.model._net
        nd = net.nd
        if self._crit is None:
            self._crit = (VAECriterion(nd, h["x_sigma"], h["br"]).cuda(), ClsCriterion())
        elbo_criterion, cls_criterion = self._crit
        s = dict(ts_any.sched)
        dev = net.device
        image_l, image_u = image_l.to(dev).float(), image_u.to(dev).float()
        label_l, label_u = label_l.to(dev).long(), label_u.to(dev).long()
        onehot = lambda y: torch.zeros(y.size(0), nd, device=dev).scatter_(1, y.view(-1, 1), 1)
        model.train()
        model.device_noise = self.device_noise
        bl, bu = image_l.size(0), image_u.size(0)
        rec, mu, ls, la = model(image_l, disc_label=label_l)
        rl, kc, kd = elbo_criterion(image_l, rec, mu, ls, la)
        elbo = rl + s["kbc"] * torch.abs(kc - s["cmi"]) + s["kbd"] * torch.abs(kd - s["dmi"])
        with torch.no_grad():
            s_img, s_mu, s_sig, s_al, s_lab, lam = label_smoothing(image_l, mu, ls, la, epsilon=h["epsilon"], disc_label=label_l)
        _, mu2, ls2, la2 = model(s_img, True, label_l, s_lab, lam)
        disc_post = lam * cls_criterion(la2, onehot(label_l)) + (1 - lam) * cls_criterion(la2, onehot(s_lab))
        cont_post = (F.mse_loss(mu2, s_mu, reduction="sum") + F.mse_loss(torch.exp(ls2), s_sig, reduction="sum")) / bl
        (s["ew"] * (elbo + s["kbc"] * s["pwm"] * cont_post) + disc_post).backward()
        rec, mu, ls, la = model(image_u)
        with torch.no_grad():
            sm = torch.full((bu, nd), 0.001 / (nd - 1), device=dev).scatter_(1, label_u.view(-1, 1), 1 - 0.001)
            kl_inf = float(torch.sum(torch.exp(la) * (la - torch.log(sm))) / bu)          # (:331-339)
        ru, kcu, kdu = elbo_criterion(image_u, rec, mu, ls, la)
        elbo_u = ru + s["kbc"] * torch.abs(kcu - s["cmi"]) + s["kbd"] * torch.abs(kdu - s["dmi"])
        with torch.no_grad():
            m_img, m_mu, m_sig, m_al, lam_u = mixup_vae_data(image_u, mu, ls, la, optimal_match=h["om"])
        _, mu4, ls4, la4 = model(m_img)
        cont_u = (F.mse_loss(mu4, m_mu, reduction="sum") + F.mse_loss(torch.exp(ls4), m_sig, reduction="sum")) / bu
        (s["ew"] * (elbo_u + s["kbc"] * s["pwm"] * cont_u) + s["ucw"] * cls_criterion(la4, m_al)).backward()
        if self.reducer is not None:
            self.reducer.bucket_ready("decoder"); self.reducer.bucket_ready("encoder"); self.reducer.wait_all()
        if ts_any._steps_done == 1:
            ts_any.sgd_hyper[4:5].zero_()
        check(lib.sv_sgd_step(ptr(net.params), ptr(net.grads), ptr(net.momentum), ptr(ts_any.sgd_hyper), net.n_params, _abi.stream()))
        for t in self.steps.values():
            t._steps_done += 1
        net.param_epoch += 1
        return kl_inf

    def train_epoch(self, loader_u, loader_l, epoch):
        """one epoch of `zip(cycle(loader_l), loader_u)` (:280); returns dict(kl_inference, steps, images, ragged_steps)"""
        kl_sum, kl_n, n_steps, n_img, n_ragged = 0.0, 0, 0, 0, 0
        pending = []                                  # (TrainStep, batch) whose device-side KL accumulators are read at the end
        for ts in self.steps.values():
            ts.kl_sum.zero_()
        for (image_l, label_l), (image_u, label_u) in zip(itertools.cycle(loader_l), loader_u):
            bl, bu = image_l.size(0), image_u.size(0)
            if self.m2 and bl != bu:                  # main_M2_vae.py:259-266
                b = min(bl, bu)
                image_l, label_l, image_u, label_u = image_l[:b], label_l[:b], image_u[:b], label_u[:b]
                bl = bu = b
            if bl == bu:
                ts = self._step_for(bl)
                if ts._epoch != epoch or ts._lr != self.lr_at(epoch) or ts.h["ewm"] != self.ewm_at(epoch):
                    self._sync_optimizer(ts, epoch)
                    ts.kl_sum.zero_()
                ts._steps_done = max(t._steps_done for t in self.steps.values())
                if ts._steps_done >= 1:
                    ts.sgd_hyper[4:5].zero_()
                ts.load_inputs(image_l, label_l, image_u, label_u)
                ts.run_resident()
                if ts not in [p[0] for p in pending]:
                    pending.append((ts, bl))
            else:
                ts = self._step_for(self.batch)
                if ts._epoch != epoch or ts._lr != self.lr_at(epoch) or ts.h["ewm"] != self.ewm_at(epoch):
                    self._sync_optimizer(ts, epoch)
                kl = self._ragged_step(image_l, label_l, image_u, label_u, epoch, ts)
                kl_sum += kl * bu; kl_n += bu; n_ragged += 1
            n_steps += 1
            n_img += bu
        torch.cuda.synchronize()
        for ts, b in pending:                         # AverageMeter.update(value, n = batch): batch-weighted mean (:339)
            kl_sum += float(ts.kl_sum) * b
            kl_n += int(ts.kl_count) * b
            ts.kl_count = 0
        return dict(kl_inference=kl_sum / max(kl_n, 1), steps=n_steps, images=n_img, ragged_steps=n_ragged)

    # ---- checkpoints (reference format, :237-251, :386-406) ------------------------------------------------------
    def checkpoint(self, epoch):
        ts = next(iter(self.steps.values()), None)
        opt = ts.state_dict() if ts is not None else self._opt_state
        args = dict(self.h, dataset=self.dataset, adjust_lr=list(self.adjust_lr), annotated_ratio=self.annotated_ratio, m2=self.m2,
                    batch_size=self.batch)
        return {"epoch": epoch + 1, "args": args, "state_dict": {k: v.detach().clone() for k, v in self.model.state_dict().items()},
                "optimizer": opt}

    def save_checkpoint(self, path, epoch):
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        torch.save(self.checkpoint(epoch), path)

    def load_checkpoint(self, ckpt):
        """ckpt: a dict from checkpoint() / torch.load of a file written by save_checkpoint OR by the reference's own
        save_checkpoint (args is then an argparse.Namespace; state_dict keys may carry nn.DataParallel's '.module.')"""
        if isinstance(ckpt, str):
            ckpt = torch.load(ckpt, map_location="cpu", weights_only=False)
        self.model.load_state_dict(ckpt["state_dict"])
        self.start_epoch = int(ckpt["epoch"])
        args = ckpt.get("args")
        if args is not None:
            get = (lambda k, d: args.get(k, d)) if isinstance(args, dict) else (lambda k, d: getattr(args, k, d))
            for k in ("akb", "aew", "apw", "kbmc", "kbmd", "pwm", "wrd", "wmf", "cmi", "dmi", "epsilon", "om", "epochs", "x_sigma", "br", "wd"):
                if get(k, None) is not None:
                    self.h[k] = get(k, None)
            self.base_lr = float(get("lr", self.base_lr))
            self.h["momentum"] = float(get("momentum", get("beta1", self.h["momentum"])))
            # the reference pickles args AFTER `ewm *= 5` was applied (:255-258): undo it to recover the base value
            ewm = float(get("ewm", self.base_ewm))
            boosted = (not self.m2) and self.dataset == "Cifar10" and self.annotated_ratio >= 0.05 and self.start_epoch > self.adjust_lr[0]
            self.base_ewm = ewm / 5.0 if (boosted and not isinstance(args, dict)) else (float(get("ewm_base", ewm)) if isinstance(args, dict) else ewm)
            self.h["ewm"] = self.base_ewm
        self.h["lr"] = self.base_lr
        opt = ckpt.get("optimizer")
        for ts in self.steps.values():
            ts.h.update(self.h)
        if opt is not None:
            if self.steps:
                for ts in self.steps.values():
                    ts.load_state_dict(opt)
            else:
                self._opt_state = opt
        return self.start_epoch

    def fit(self, loader_u, loader_l, epochs, on_epoch_end=None, checkpoint_path=None):
        """the epoch loop of main() (:222-258) without validation: train, checkpoint, schedule"""
        out = []
        for epoch in range(self.start_epoch, epochs):
            r = self.train_epoch(loader_u, loader_l, epoch)
            r["epoch"], r["lr"] = epoch, self.lr_at(epoch)
            out.append(r)
            if checkpoint_path:
                self.save_checkpoint(checkpoint_path, epoch)
            if on_epoch_end is not None:
                on_epoch_end(self, r)
        return out
