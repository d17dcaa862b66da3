"""Fused training-step engine: the loop body of main_shot_vae.train (:281-366) / main_M2_vae.train
(:258-307) as one explicit launch sequence on libshotvae -- no autograd, no per-op Python in the
steady state (the sequence is captured into a CUDA graph and replayed).

B200-first restructuring of the reference step (results unchanged):
  * the four network passes are batched two by two: [P1 labelled | P3 unlabelled] and then
    [P2 label-smoothed | P4 mixed-up] run as G=2 pass groups through the same launches (BatchNorm
    statistics stay per pass), halving the launch count and doubling the tiles per launch;
  * both backwards accumulate straight into the flat gradient arena; loss scalars, their signs and
    every schedule coefficient live in device memory, so nothing synchronises with the host;
  * SGD (momentum + weight decay) is one kernel over the flat arena; BatchNorm running statistics are
    updated once per step in the reference's pass order P1, P2, P3, P4;
  * with world_size > 1 the gradient arena is all-reduced in buckets on a side stream while the
    remaining backward runs (ddp.py).
"""
import contextlib
import math

import numpy as np
import os

import torch

from . import _abi
from ._abi import lib, check, ptr
from .plan import Ctx, pad16

TERM_NAMES = ("rec_l", "klc_l", "kld_l", "rec_u", "klc_u", "kld_u", "disc_post_l", "cont_post_l", "disc_post_u",
              "cont_post_u", "kl_inference")


def alpha_schedule(epoch, max_epoch, alpha_max):
    """main_shot_vae.py:518-520"""
    return alpha_max * math.exp(-5 * (1 - min(1, epoch / max_epoch)) ** 2)


def default_hyper(dataset="Cifar10", m2=False):
    """argparse defaults of main_shot_vae.py:30-106 with the per-dataset overrides (:139,:161-163;
    main_M2_vae.py:123,146-147)"""
    h = dict(epochs=600, akb=200, aew=400, apw=200, ewm=1e-3, kbmc=1e-3, kbmd=1e-3, pwm=1.0, wrd=1.0, wmf=0.4, cmi=0.0,
             dmi=2.3, epsilon=0.1, om=False, lr=0.1, momentum=0.9, wd=5e-4, x_sigma=1.0, br=True)
    if dataset == "Cifar100":
        h.update(akb=150, apw=400, dmi=4.6)
    if m2:
        h.update(cmi=200.0 if dataset == "Cifar10" else 1280.0)
    return h


class TrainStep:
    def __init__(self, model, batch, hyper=None, m2=False, use_graph=True, device_noise=True, skip_dead_decoders=False,
                 reducer=None):
        model._ensure_bound()
        self.model, self.net = model, model._net
        net = self.net
        self.B, self.m2 = int(batch), bool(m2)
        self.h = dict(default_hyper(m2=m2) if hyper is None else hyper)
        self.use_graph, self.device_noise = use_graph, device_noise
        self.skip_dead_decoders = skip_dead_decoders
        self.reducer = reducer
        self.cta_limit = getattr(reducer, "cta_limit", 0) if reducer is not None else 0
        # Data parallel: one CUDA graph per backward segment, the NCCL calls issued between them from the host.
        # SHOTVAE_DDP_GRAPH=single captures the all-reduces into ONE graph instead (works: NCCL supports stream capture) --
        # MEASURED at 2 GPUs: 5.005 vs 5.021 ms/step, i.e. the host-side gaps are not what the 2 % data-parallel overhead is made
        # of, and destroy_process_group() hung for minutes with the captured communicator still referenced -> opt-in only
        self.fwd_overlap = os.environ.get("SHOTVAE_FWD_OVERLAP", "1") != "0"
        self.ddp_single_graph = os.environ.get("SHOTVAE_DDP_GRAPH", "segments") == "single"
        # weight gradients overlap the dgrad / BatchNorm-backward chain on a second stream (SHOTVAE_SIDE=0 turns it off)
        if net.side is None and os.environ.get("SHOTVAE_SIDE", "1") != "0":
            net.side = torch.cuda.Stream(device=net.device)
        self.side2 = torch.cuda.Stream(device=net.device) if net.side is not None else None
        # The forward of [P1 | P3] and its sample -> decoder -> ELBO -> decoder-backward chain (the critical chain between the
        # encoder forwards and the encoder backward) run on a HIGH-priority stream: stream priorities survive CUDA-graph capture
        # (kernel-node priority), so its kernels get SMs before the dead decoder forwards and the decoder weight gradients that
        # run beside it.  MEASURED: C2 4.632 -> 4.589 ms/step (A/B x 3), C4 / C5 unchanged.  SHOTVAE_PRIO=0: default priority.
        self.side_a = torch.cuda.Stream(device=net.device, priority=-1) if (net.side is not None and os.environ.get("SHOTVAE_PRIO", "1") != "0") else None
        # last-block backward of [P2 | P4] early (see _part0): its own pair of streams
        self.bwd_split = net.side is not None and (not m2) and os.environ.get("SHOTVAE_BWD_SPLIT", "0") != "0"      # MEASURED: 4.93 vs 4.63 ms/step -- the window is not idle enough, off
        self.side3 = torch.cuda.Stream(device=net.device) if self.bwd_split else None
        self.side4 = torch.cuda.Stream(device=net.device) if self.bwd_split else None
        if net.side is not None and net.side_dec is None and os.environ.get("SHOTVAE_DEC_SIDE", "1") != "0":
            net.side_dec = torch.cuda.Stream(device=net.device)       # decoder weight gradients beside the decoder's dgrad chain
        dev, B, nd, D = net.device, self.B, net.nd, net.ldc
        self.dev = dev
        f32, i64 = torch.float32, torch.int64
        z = lambda *s, dtype=f32: torch.zeros(*s, dtype=dtype, device=dev)
        # static device inputs: views of ONE arena, so that the pipelined end-to-end path (step_async) moves a step's inputs
        # with one host -> device and one device -> device copy
        self._in_layout = [("img_l", (B, net.in_ch, 32, 32), f32), ("img_u", (B, net.in_ch, 32, 32), f32), ("label_l", (B,), i64),
                           ("label_u", (B,), i64), ("lam", (4,), f32),       # lam = {lam_l, 1-lam_l, lam_u, 1-lam_u}
                           ("idx_l", (B,), i64), ("idx_u", (B,), i64), ("s_lab", (B,), i64)]
        self._in_arena, iv = self._input_arena(lambda n: torch.zeros(n, dtype=torch.uint8, device=dev))
        self.img_l, self.img_u, self.label_l, self.label_u = iv["img_l"], iv["img_u"], iv["label_l"], iv["label_u"]
        self.lam, self.idx_l, self.idx_u, self.s_lab = iv["lam"], iv["idx_l"], iv["idx_u"], iv["s_lab"]
        self._pipe = None            # step_async state (staging slots, copy stream), built on first use
        self.eps, self.unif = z(4, B, D), z(2, B, nd)
        # device-resident scalars
        self.coef = z(16)
        self.sgd_hyper = z(8)
        self.terms = z(16)
        # device noise (sv_noise_fill): {seed, offset, ticket, pad}; the seed follows torch's CUDA seed (reading it consumes no
        # host RNG draw, so the reference's randperm / beta sequence is untouched) and the rank, streams of different
        # TrainStep objects start 2^40 counters apart
        TrainStep._instances = getattr(TrainStep, "_instances", 0) + 1
        rank = 0 if reducer is None else int(torch.distributed.get_rank())
        seed = (int(torch.cuda.initial_seed()) * 0x9E3779B97F4A7C15 + rank * 0xD1B54A32D192ED03) & 0x7FFFFFFFFFFFFFFF
        self.noise_state = torch.tensor([seed, TrainStep._instances << 40, 0, 0], dtype=i64, device=dev)
        self.kl_sum = z(1)           # sum of the per-step inference-KL monitor since it was last cleared (Train/KL_Inference, :331-339,376)
        self.kl_count = 0
        # mixup targets
        self.s_mu, self.s_sig, self.s_alpha = z(B, D), z(B, D), z(B, nd)
        self.m_mu, self.m_sig, self.m_alpha = z(B, D), z(B, D), z(B, nd)
        # SHOT step: the two forward launch sequences [P1|P3] and [P2|P4] write into the two halves of ONE
        # 4-group context, whose encoder/heads backward then runs as a single G=4 launch sequence
        if m2:
            self.ctxS = self.ctxA = Ctx(net, 2, B)
            self.ctxB = None
        else:
            self.ctxS = Ctx(net, 4, B)
            self.ctxA = Ctx(net, 2, B, parent=self.ctxS, g0=0)
            self.ctxB = Ctx(net, 2, B, parent=self.ctxS, g0=2)
        self._adopted = False
        self.graph = None
        self.launches_per_step = None
        self._epoch = None
        self._steps_done = 0         # optimizer steps taken (momentum first-step rule; restored by load_state_dict)
        self._calls = 0              # run_resident() calls on THIS object (eager warm-up, then graph capture)
        # pinned host staging for the end-to-end path
        pin = lambda *s, dtype=f32: torch.zeros(*s, dtype=dtype).pin_memory()
        self.h_img_l, self.h_img_u = pin(B, net.in_ch, 32, 32), pin(B, net.in_ch, 32, 32)
        self.h_label_l, self.h_label_u = pin(B, dtype=i64), pin(B, dtype=i64)
        self.h_lam, self.h_idx = pin(4), pin(3, B, dtype=i64)
        self.h_terms = pin(16)
        self.set_epoch(0)
        self.set_lr(self.h["lr"])
        net.zero_grads()

    def _input_arena(self, alloc):
        """one byte arena holding every per-step input in self._in_layout (16-byte aligned fields) -> (arena, {name: typed view})"""
        offs, off = [], 0
        for name, shape, dt in self._in_layout:
            n = int(np.prod(shape)) * torch.empty((), dtype=dt).element_size()
            offs.append((name, shape, dt, off, n))
            off += (n + 15) // 16 * 16
        arena = alloc(off)
        return arena, {name: arena[o:o + n].view(dt).view(shape) for name, shape, dt, o, n in offs}

    # ---- schedules ---------------------------------------------------------------------------
    def set_epoch(self, epoch):
        h = self.h
        cmi, dmi = alpha_schedule(epoch, h["akb"], h["cmi"]), alpha_schedule(epoch, h["akb"], h["dmi"])
        ew = alpha_schedule(epoch, h["aew"], h["ewm"])
        kbc, kbd = alpha_schedule(epoch, h["akb"], h["kbmc"]), alpha_schedule(epoch, h["akb"], h["kbmd"])
        pwm = alpha_schedule(epoch, h["apw"], h["pwm"])
        ucw = alpha_schedule(epoch, round(h["wmf"] * h["epochs"]), h["wrd"])
        self.sched = dict(cmi=cmi, dmi=dmi, ew=ew, kbc=kbc, kbd=kbd, pwm=pwm, ucw=ucw)
        c = torch.zeros(16, dtype=torch.float32)
        c[0] = ew
        c[1:6] = torch.tensor([ew, kbc, cmi, kbd, dmi])
        c[6:8] = torch.tensor([1.0, 0.0 if self.m2 else ew * kbc * pwm])
        c[8:10] = torch.tensor([ucw, ew * kbc * pwm])
        self.coef.copy_(c)
        self._epoch = epoch

    def set_lr(self, lr):
        world = 1 if self.reducer is None else self.reducer.world
        hy = torch.tensor([lr, self.h["momentum"], self.h["wd"], 1.0 / world, 1.0 if self._steps_done == 0 else 0.0, 0, 0, 0],
                          dtype=torch.float32)
        self.sgd_hyper.copy_(hy)
        self._lr = lr

    # ---- host-side draws (reference: np.random.beta / torch.randperm on the host) -----------------
    def draw_host(self):
        """lambda and pairing draws in the reference's order: beta(eps,eps), randperm (label smoothing),
        beta(2,2), randperm (mixup)."""
        B = self.B
        if self.m2:
            return None
        lam_l = float(np.random.beta(self.h["epsilon"], self.h["epsilon"])) if self.h["epsilon"] > 0 else 1.0
        idx_l = torch.randperm(B)
        lam_u = float(np.random.beta(2.0, 2.0))
        idx_u = torch.arange(B) if self.h["om"] else torch.randperm(B)
        return lam_l, idx_l, lam_u, idx_u

    def stage_draws(self, draws, label_l):
        """label_l: the labelled batch's labels, host (pinned staging) or device"""
        if draws is None:
            return
        lam_l, idx_l, lam_u, idx_u = draws
        self.h_lam.copy_(torch.tensor([lam_l, 1 - lam_l, lam_u, 1 - lam_u], dtype=torch.float64).float())
        self.h_idx[0].copy_(idx_l)
        self.h_idx[1].copy_(idx_u)
        self.lam.copy_(self.h_lam, non_blocking=True)
        self.idx_l.copy_(self.h_idx[0], non_blocking=True)
        if not self.h["om"]:
            self.idx_u.copy_(self.h_idx[1], non_blocking=True)
        if label_l.is_cuda:
            torch.index_select(label_l, 0, self.idx_l, out=self.s_lab)     # smoothed_disc_label = disc_label[index] (mixup.py:37)
        else:
            self.h_idx[2].copy_(label_l[idx_l])
            self.s_lab.copy_(self.h_idx[2], non_blocking=True)

    # ---- the launch sequence -----------------------------------------------------------------------
    def _losses_first(self, ctx, rec):
        """ELBO terms + gradients for the [labelled | unlabelled] group pair"""
        net, B, st = self.net, self.B, _abi.stream()
        nd, D, ch = net.nd, net.ldc, net.in_ch
        mu, ls, la = ctx.bufs["mu"], ctx.bufs["ls"], ctx.bufs["la"]
        cp = pad16(ch)
        g_rec = ctx.t("g.rec", (2 * B, 32, 32, cp))
        g_mu, g_ls, g_la = ctx.t("g.mu", (2 * B, D), torch.float32), ctx.t("g.ls", (2 * B, D), torch.float32), \
            ctx.t("g.la", (2 * B, nd), torch.float32)
        bce = 1 if self.h["br"] else 0
        for g, img in ((0, self.img_l), (1, self.img_u)):
            r = slice(g * B, (g + 1) * B)
            t = self.terms[3 * g:]
            if net.f32:
                # FP32 mode: fp32 gradient in the layout of x_hat (NHWC, `ch` channels), then padded to the decoder's 16
                g32 = ctx.t("g.rec.f32", (2 * B, 32, 32, ch), torch.float32)
                check(lib.sv_elbo_rec_fwd_bwd(ptr(img), ptr(rec[r]), 1, B, ch, 32 * 32, bce, float(self.h["x_sigma"]), ptr(self.coef),
                                              ptr(t), None, 0, ptr(g32[r]), st))
                check(lib.sv_pad_channels_f32(ptr(g32[r]), ptr(g_rec[r]), B * 32 * 32, ch, cp, st))
            else:
                check(lib.sv_elbo_rec_fwd_bwd(ptr(img), ptr(rec[r]), 1, B, ch, 32 * 32, bce, float(self.h["x_sigma"]), ptr(self.coef),
                                              ptr(t), ptr(g_rec[r]), cp, None, st))
            check(lib.sv_elbo_kl_fwd(ptr(mu[r]), ptr(ls[r]), ptr(la[r]), B, D, nd, ptr(t), st))
            check(lib.sv_elbo_kl_bwd(ptr(mu[r]), ptr(ls[r]), ptr(la[r]), ptr(t), ptr(self.coef[1:]), 0, B, D, nd, ptr(g_mu[r]),
                                     ptr(g_ls[r]), ptr(g_la[r]), 0, st))
        check(lib.sv_inference_kl(ptr(la[B:]), ptr(self.label_u), B, nd, ptr(self.terms[10:]), st))
        check(lib.sv_inference_kl(ptr(la[B:]), ptr(self.label_u), B, nd, ptr(self.kl_sum), st))      # epoch accumulator (+=): no per-step host sync
        return g_rec, g_mu, g_ls, g_la

    def _noise(self):
        if self.device_noise:
            check(lib.sv_noise_fill(ptr(self.eps), self.eps.numel(), ptr(self.unif), self.unif.numel(), ptr(self.noise_state),
                                    _abi.stream()))

    def _sequence(self, parts=(0, 1, 2)):
        """parts: 0 = forwards, losses, backward of [P2|P4] and the decoder backward of [P1|P3] (after it the
        decoder gradient bucket is final); 1 = rest of the backward -- or (1, i) = its i-th segment (one resolution block of
        the encoder, plan.Net.bwd_segments); 2 = SGD + BatchNorm running statistics.
        With several ranks the parts are separate CUDA graphs and the bucket all-reduces are issued between
        them on a side stream (NCCL is kept out of the captured graphs)."""
        # data parallel: the backward overlaps the bucket all-reduces -> its persistent grids leave NCCL's SMs alone (ddp.py)
        limit = self.cta_limit        # (fixed at construction: the per-launch argument records are built for this grid size)
        for part in parts:
            if part == 0:
                self._part0()
            elif part == 2:
                self._part2()
            elif part == "2a":
                self._part2(rng=(self._sgd_split, self.net.n_params), running=False)
            elif part == "2b":
                self._part2(rng=(0, self._sgd_split), running=True)
            else:
                if limit:
                    lib.sv_set_cta_limit(limit)
                try:
                    self._part1(seg=None if part == 1 else part[1])
                finally:
                    if limit:
                        lib.sv_set_cta_limit(0)

    def _ddp_plan(self):
        """[(parts of one CUDA graph, bucket to all-reduce after it)] for a data-parallel step"""
        segs = self.reducer.segments if self.reducer is not None else []
        if len(segs) < 2:
            return [((0,), "decoder"), ((1,), "encoder"), ((2,), None)]
        plan = [((0,), "decoder")] + [(((1, i),), name) for i, name in enumerate(segs)]
        # The optimizer in two launches: every range whose all-reduce is already behind the compute stream (all buckets but the
        # last, small, fully exposed one) is updated WHILE that last all-reduce is in flight, the rest after it.  The split
        # point is rounded up to the kernel's 16-byte granularity; `wait` names the bucket the part waits for.
        lo_last, hi_last = self.reducer.buckets[segs[-1]]
        self._sgd_split = min(self.net.n_params, (hi_last + 3) // 4 * 4) if lo_last == 0 else 0
        if self._sgd_split and os.environ.get("SHOTVAE_DDP_SGD_SPLIT", "1") != "0":
            return plan + [(("2a",), None, segs[-2]), (("2b",), None, None)]
        return plan + [((2,), None)]

    def _part0(self):
        net, B, st = self.net, self.B, _abi.stream()
        nd, D, ch = net.nd, net.ldc, net.in_ch
        cp = pad16(ch)
        A, Bc = self.ctxA, self.ctxB
        self.ctxS.reset()
        check(lib.sv_fill_zero(ptr(self.terms), self.terms.numel() * 4, st))
        net.pack_weights()
        self._noise()
        # Without --om the inputs of [P2 | P4] are mixtures of the INPUT images under host-drawn pairings: their encoder forward
        # depends on nothing [P1 | P3] computes (only the loss targets do), so it runs on a second stream beside the first one.
        # Every conv kernel is a persistent 148-CTA grid of 15-25 us with ~45 % of that in prologue / first-tile latency / tail
        # (profiles/r02_kernel_findings.md): two independent chains fill each other's gaps.  SHOTVAE_FWD_OVERLAP=0: one after the other.
        conc = (not self.m2) and (not self.h["om"]) and (not net.f32) and net.side is not None and self.side2 is not None and self.fwd_overlap
        if conc:
            ev0 = torch.cuda.Event()
            ev0.record(torch.cuda.current_stream())
            self.side2.wait_event(ev0)
            with torch.cuda.stream(self.side2):
                s2 = _abi.stream()
                xB = Bc.t("x_img", (2 * B, 32, 32, cp))
                check(lib.sv_mixup_lerp(ptr(self.img_l), None, None, None, ptr(self.idx_l), ptr(self.lam), B, ch, 32 * 32, D, nd, None,
                                        ptr(xB[:B]), cp, None, None, None, s2))
                check(lib.sv_mixup_lerp(ptr(self.img_u), None, None, None, ptr(self.idx_u), ptr(self.lam[2:]), B, ch, 32 * 32, D, nd, None,
                                        ptr(xB[B:]), cp, None, None, None, s2))
                featB = net.encoder_fwd(Bc, xB)
                mu2, ls2, la2 = net.heads_fwd(Bc, featB)
        # ---- forward of [P1 | P3].  With SHOTVAE_PRIO=1 and the two forwards overlapped it runs on the HIGH-priority stream that
        # also carries the decoder chain: [P1 | P3]'s forward then finishes first and its decoder chain -- the critical path to
        # the encoder backward -- overlaps the rest of [P2 | P4]'s forward instead of starting when both forwards end.
        hp = conc and self.side_a is not None
        evA = None
        if hp:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            self.side_a.wait_event(ev)
        with (torch.cuda.stream(self.side_a) if hp else contextlib.nullcontext()):
            xA = A.t("x_img", (2 * B, 32, 32, cp))
            check(net.fn("sv_pack_image")(ptr(self.img_l), ptr(xA[:B]), B, ch, 32 * 32, cp, _abi.stream()))
            check(net.fn("sv_pack_image")(ptr(self.img_u), ptr(xA[B:]), B, ch, 32 * 32, cp, _abi.stream()))
            feat = net.encoder_fwd(A, xA)
            mu, ls, la = net.heads_fwd(A, feat)
            if hp:
                evA = torch.cuda.Event()
                evA.record(self.side_a)
        # From here two independent chains run until part 1: (a) sample -> decoder forward -> ELBO terms -> decoder
        # backward of [P1 | P3], and (b) mixup -> encoder forward of [P2 | P4] -> posterior-matching terms.  (a) goes to
        # the side stream: its many small launches (1x1 ... 8x8 decoder layers, loss reductions) fill the gaps of (b)'s
        # large persistent convolutions.
        side = net.side if not self.m2 else None
        if side is not None and self.side_a is not None:
            side = self.side_a          # (experiment: the critical decoder chain on a high-priority stream, see __init__)
        main = torch.cuda.current_stream()

        def fork():
            if side is None:
                return contextlib.nullcontext()
            ev = torch.cuda.Event()
            ev.record(main)
            side.wait_event(ev)
            return torch.cuda.stream(side)

        with fork():
            net.sample_fwd(A, 0, 0, self.eps[0], label=self.label_l)
            lat = net.sample_fwd(A, 1, 2, self.eps[2], unif=self.unif[0])
            rec = net.decoder_fwd(A, lat)
            g_rec, g_mu, g_ls, g_la = self._losses_first(A, rec)
            if self.m2:
                # supervised cross-entropy on the labelled half (main_M2_vae.py:276-277)
                check(lib.sv_posterior_fwd_bwd(ptr(la[:B]), None, ptr(self.label_l), None, None, None, None, None, None,
                                               ptr(self.coef[6:]), B, D, nd, ptr(self.terms[6:]), ptr(g_la[:B]), None, None, 1,
                                               _abi.stream()))
            # backward of [P1 | P3]: decoder first (its gradients are final afterwards)
            self._g_lat = net.decoder_bwd(A, g_rec)
            self._g = (g_mu, g_ls, g_la)
        if conc:
            # the mixing targets of the posterior terms need [P1 | P3]'s latents; then join the second forward
            if evA is not None:
                main.wait_event(evA)
            check(lib.sv_mixup_lerp(None, ptr(mu[:B]), ptr(ls[:B]), ptr(la[:B]), ptr(self.idx_l), ptr(self.lam), B, ch, 32 * 32, D, nd,
                                    None, None, 0, ptr(self.s_mu), ptr(self.s_sig), ptr(self.s_alpha), st))
            check(lib.sv_mixup_lerp(None, ptr(mu[B:]), ptr(ls[B:]), ptr(la[B:]), ptr(self.idx_u), ptr(self.lam[2:]), B, ch, 32 * 32, D, nd,
                                    None, None, 0, ptr(self.m_mu), ptr(self.m_sig), ptr(self.m_alpha), st))
            main.wait_stream(self.side2)
        elif not self.m2:
            # ---- label smoothing (P1 -> P2 inputs) and optimal-interpolation mixup (P3 -> P4 inputs)
            xB = Bc.t("x_img", (2 * B, 32, 32, cp))
            # (FP32 mode: the mixed images are written as fp32 NCHW and then laid out NHWC like any input batch)
            xm = Bc.t("x_mix.f32", (2 * B, ch, 32, 32), torch.float32) if net.f32 else None
            o32 = (lambda r: ptr(xm[r])) if net.f32 else (lambda r: None)
            o16 = (lambda r: None) if net.f32 else (lambda r: ptr(xB[r]))
            lo, hi = slice(0, B), slice(B, 2 * B)
            check(lib.sv_mixup_lerp(ptr(self.img_l), ptr(mu[:B]), ptr(ls[:B]), ptr(la[:B]), ptr(self.idx_l), ptr(self.lam), B, ch,
                                    32 * 32, D, nd, o32(lo), o16(lo), cp, ptr(self.s_mu), ptr(self.s_sig), ptr(self.s_alpha), st))
            if self.h["om"]:
                check(lib.sv_pairwise_kl_second_nearest(ptr(mu[B:]), ptr(ls[B:]), B, D, ptr(self.idx_u), None, st))
            check(lib.sv_mixup_lerp(ptr(self.img_u), ptr(mu[B:]), ptr(ls[B:]), ptr(la[B:]), ptr(self.idx_u), ptr(self.lam[2:]), B, ch,
                                    32 * 32, D, nd, o32(hi), o16(hi), cp, ptr(self.m_mu), ptr(self.m_sig), ptr(self.m_alpha), st))
            if net.f32:
                check(lib.sv_pack_image_f32(ptr(xm), ptr(xB), 2 * B, ch, 32 * 32, cp, st))
            # ---- forward of [P2 | P4]
            featB = net.encoder_fwd(Bc, xB)
            mu2, ls2, la2 = net.heads_fwd(Bc, featB)
        if not self.m2:
            # (the decoder forwards of P2/P4 -- kept only for their BatchNorm running statistics -- run beside the
            # encoder backward in part 1: nothing in the step waits for them before the running-statistics update)
            g_mu2, g_ls2, g_la2 = Bc.t("g.mu", (2 * B, D), torch.float32), Bc.t("g.ls", (2 * B, D), torch.float32), \
                Bc.t("g.la", (2 * B, nd), torch.float32)
            check(lib.sv_posterior_fwd_bwd(ptr(la2[:B]), None, ptr(self.label_l), ptr(self.s_lab), ptr(self.lam), ptr(mu2[:B]),
                                           ptr(ls2[:B]), ptr(self.s_mu), ptr(self.s_sig), ptr(self.coef[6:]), B, D, nd,
                                           ptr(self.terms[6:]), ptr(g_la2[:B]), ptr(g_mu2[:B]), ptr(g_ls2[:B]), 0, st))
            check(lib.sv_posterior_fwd_bwd(ptr(la2[B:]), ptr(self.m_alpha), None, None, None, ptr(mu2[B:]), ptr(ls2[B:]),
                                           ptr(self.m_mu), ptr(self.m_sig), ptr(self.coef[8:]), B, D, nd, ptr(self.terms[8:]),
                                           ptr(g_la2[B:]), ptr(g_mu2[B:]), ptr(g_ls2[B:]), 0, st))
            # (the backward of [P2 | P4] -- heads + encoder only -- runs together with [P1 | P3] in part 1)
        # The heads + last-resolution-block backward of [P2 | P4] needs only the posterior-matching gradients just computed, not
        # the decoder chain of [P1 | P3]: it runs NOW, as a G = 2 launch sequence on its own streams, beside that chain (small
        # grids: 1x1 ... 16x16 decoder layers) instead of after it as half of the G = 4 backward.  Part 1 then runs the same block
        # for [P1 | P3] only and continues with both pass groups merged from the next block on.
        self._split_done = False
        if conc and self.bwd_split:
            ev = torch.cuda.Event()
            ev.record(main)
            self.side3.wait_event(ev)
            with torch.cuda.stream(self.side3):
                gfB = net.heads_bwd(Bc, g_mu2, g_ls2, g_la2, side=self.side4)
                net.encoder_bwd(Bc, gfB, seg=0, side=self.side4)
            self._split_done = True
        # The decoder forwards of P2 / P4 (kept only for their BatchNorm running statistics) need nothing but [P2 | P4]'s heads.
        # With the two encoder forwards overlapped, the sample -> decoder -> ELBO -> decoder-backward chain of [P1 | P3] is the
        # ONLY chain in flight from here to part 1 (timeline: ~650 us with one kernel at a time, profiles/r02_timeline_c2.txt, profiles/r02_kernel_findings.md section 8):
        # the dead forwards fill it instead of competing with the encoder backward.  SHOTVAE_DEAD_EARLY=0: beside part 1.
        self._dead_done = False
        if conc and not self.skip_dead_decoders and os.environ.get("SHOTVAE_DEAD_EARLY", "1") != "0":
            ev = torch.cuda.Event()
            ev.record(main)
            self.side2.wait_event(ev)
            with torch.cuda.stream(self.side2):
                self._dead_decoders()
            main.wait_stream(self.side2)
            self._dead_done = True
        if side is not None:
            main.wait_stream(side)
        if self._split_done:
            main.wait_stream(self.side3)

    def _dead_decoders(self):
        """decoder forwards of [P2 | P4]: the reference discards these reconstructions (main_shot_vae.py:311,356) but the
        forwards still update the decoder's BatchNorm running statistics"""
        net, Bc = self.net, self.ctxB
        net.sample_fwd(Bc, 0, 1, self.eps[1], label=self.label_l, label_mix=self.s_lab, lam_dev=self.lam)
        latB = net.sample_fwd(Bc, 1, 2, self.eps[3], unif=self.unif[1])
        net.decoder_fwd(Bc, latB)

    def _part1(self, seg=None):
        net, A, S = self.net, self.ctxA, self.ctxS
        if seg is not None and seg > 0:
            net.encoder_bwd(S, None, seg=seg)
            return
        g_mu, g_ls, g_la = self._g
        dead = (not self.m2) and (not self.skip_dead_decoders) and not getattr(self, "_dead_done", False)
        main = torch.cuda.current_stream()
        if dead and self.side2 is not None:
            ev = torch.cuda.Event()
            ev.record(main)
            self.side2.wait_event(ev)
            with torch.cuda.stream(self.side2):
                self._dead_decoders()
        elif dead:
            self._dead_decoders()
        net.sample_bwd(A, 0, self._g_lat, g_mu, g_ls, None, accumulate=1)
        net.sample_bwd(A, 1, self._g_lat, g_mu, g_ls, g_la, accumulate=1)
        if S is not A and not self._adopted:
            net.adopt_views(S)
            self._adopted = True
        D, nd = net.ldc, net.nd
        if getattr(self, "_split_done", False):
            # the last block of [P2 | P4] already ran in part 0: the same block for [P1 | P3], then both merged (the views' buffers
            # are the two halves of S's, so S continues from the full input-gradient buffer)
            net.encoder_bwd(A, net.heads_bwd(A, g_mu, g_ls, g_la), seg=0)
            S._bwd_state = (S.bufs[A._bwd_name], A._bwd_state[1])
            if seg is None:
                for i in range(1, len(net.bwd_segments())):
                    net.encoder_bwd(S, None, seg=i)
        else:
            # heads + encoder backward of ALL pass groups at once (g.mu / g.ls / g.la of the views are slices of S's)
            net.encoder_bwd(S, net.heads_bwd(S, S.t("g.mu", (S.NB, D), torch.float32), S.t("g.ls", (S.NB, D), torch.float32),
                                             S.t("g.la", (S.NB, nd), torch.float32)), seg=seg)
        if dead and self.side2 is not None:
            main.wait_stream(self.side2)

    def _part2(self, rng=None, running=True):
        net, A, Bc, st = self.net, self.ctxA, self.ctxB, _abi.stream()
        # ---- optimizer (rng: a [lo, hi) slice of the flat arena) + BatchNorm running statistics
        lo, hi = rng if rng is not None else (0, net.n_params)
        if hi > lo:
            check(lib.sv_sgd_step(ptr(net.params[lo:hi]), ptr(net.grads[lo:hi]), ptr(net.momentum[lo:hi]), ptr(self.sgd_hyper), hi - lo, st))
        if not running:
            return
        if self.m2:
            net.bn_running_update([(A, 0), (A, 1)])
        else:
            net.bn_running_update([(A, 0), (Bc, 0), (A, 1), (Bc, 1)])     # reference order P1, P2, P3, P4

    # ---- public API --------------------------------------------------------------------------------
    def _run_parts_eager(self):
        if self.reducer is None:
            self._sequence()
            return
        for ent in self._ddp_plan():
            parts, bucket = ent[0], ent[1]
            if bucket is None:
                self._ddp_wait(ent)
            self._sequence(parts)
            if bucket is not None:
                self.reducer.bucket_ready(bucket)      # overlaps with the next part of the backward

    def run_resident(self):
        """one optimizer step on the inputs currently resident in the static device buffers"""
        if self._steps_done == 1:
            self.sgd_hyper[4:5].zero_()       # momentum buffer is initialised; torch semantics from now on
        if not self.use_graph:
            self._run_parts_eager()
        elif self.graph is None and self._calls >= 2:
            n0 = _abi.launch_count()
            if self.reducer is not None and self.ddp_single_graph:
                # data parallel, ONE graph: the bucket all-reduces are captured with the step (NCCL supports stream capture);
                # they sit on the reducer's side stream, forked from / joined to the capture stream by bucket_ready / wait_all,
                # so a replay has no host-side gaps between the backward segments
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._run_parts_eager()
                graphs = [g]
            else:
                groups = [(0, 1, 2)] if self.reducer is None else [ent[0] for ent in self._ddp_plan()]
                graphs = []
                for parts in groups:
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        self._sequence(parts)
                    graphs.append(g)
            self.launches_per_step = _abi.launch_count() - n0
            self.graph = graphs
            self._replay()
        elif self.graph is not None:
            self._replay()
        else:
            n0 = _abi.launch_count()
            self._run_parts_eager()           # eager warm-up steps allocate every buffer
            self.launches_per_step = _abi.launch_count() - n0
        self._steps_done += 1
        self._calls += 1
        self.kl_count += 1
        self.net.param_epoch += 1             # the fused SGD moved the FP32 masters: the drop-in forward must repack

    def _replay(self):
        if self.reducer is None or len(self.graph) == 1:
            self.graph[0].replay()
            return
        for g, ent in zip(self.graph, self._ddp_plan()):
            bucket = ent[1]
            if bucket is None:
                self._ddp_wait(ent)
            g.replay()
            if bucket is not None:
                self.reducer.bucket_ready(bucket)

    def _ddp_wait(self, ent):
        """before an optimizer part of the data-parallel plan: wait for the bucket it names, or for all of them"""
        if len(ent) > 2 and ent[2] is not None:
            self.reducer.wait_bucket(ent[2])
        else:
            self.reducer.wait_all()

    def load_inputs(self, image_l, label_l, image_u, label_u, draws="auto"):
        """host -> device copy of one (labelled, unlabelled) batch pair through pinned staging buffers,
        plus the host RNG draws of this step.  Batches that already live on the device (lib.dataloader.DeviceLoader)
        are copied device to device and never touch the host."""
        # images that already sit in pinned host memory are copied straight from there (the caller must leave them
        # untouched until the step has been synchronised -- step() does that before it returns); anything else goes
        # through the pinned staging buffers
        for dst, staged, src in ((self.img_l, self.h_img_l, image_l), (self.img_u, self.h_img_u, image_u)):
            if src.is_cuda:
                dst.copy_(src)
            elif src.is_pinned() and src.dtype == dst.dtype and src.shape == dst.shape and src.is_contiguous():
                dst.copy_(src, non_blocking=True)
            else:
                staged.copy_(src)
                dst.copy_(staged, non_blocking=True)
        for dst, staged, src in ((self.label_l, self.h_label_l, label_l), (self.label_u, self.h_label_u, label_u)):
            if src.is_cuda:
                dst.copy_(src)
            else:
                staged.copy_(src)
                dst.copy_(staged, non_blocking=True)
        self.stage_draws(self.draw_host() if draws == "auto" else draws, self.label_l if label_l.is_cuda else self.h_label_l)

    def h2d_bytes(self):
        n = self.h_img_l.numel() * 4 * 2 + self.B * 8 * 2
        if not self.m2:
            n += 16 + self.B * 8 * (2 if self.h["om"] else 3)
        return n

    def step(self, image_l, label_l, image_u, label_u, draws="auto"):
        """end-to-end call: host batch in, loss terms (host floats) out"""
        self.load_inputs(image_l, label_l, image_u, label_u, draws)
        self.run_resident()
        return self.read_terms()

    # ---- pipelined end-to-end path -------------------------------------------------------------------------------
    def step_async(self, image_l, label_l, image_u, label_u, draws="auto"):
        """Pipelined end-to-end call (host batch in, loss terms out, one call per optimizer step like step()): this batch's
        inputs and host draws go through a pinned staging slot to a device staging slot on a COPY stream -- beside the previous
        step, which is still running -- and the step is enqueued behind one device-to-device copy.  Returns the loss terms of the
        PREVIOUS call (None on the first): the host never waits for the step it has just enqueued, and is never more than one
        step ahead.  drain() returns the terms of the last step.  (step() is the same work with a host synchronisation per step:
        copy, step, read.)"""
        if self._pipe is None:
            mk_h = lambda n: torch.zeros(n, dtype=torch.uint8).pin_memory()
            mk_d = lambda n: torch.zeros(n, dtype=torch.uint8, device=self.dev)
            self._pipe = dict(host=[self._input_arena(mk_h) for _ in range(2)], dev=[mk_d(self._in_arena.numel()) for _ in range(2)],
                              copy=torch.cuda.Stream(device=self.dev), k=0, pending=None,
                              h2d=[torch.cuda.Event() for _ in range(2)], free=[torch.cuda.Event() for _ in range(2)],
                              done=[torch.cuda.Event() for _ in range(2)], used=[False, False],
                              terms=[torch.zeros(16).pin_memory() for _ in range(2)])
        P = self._pipe
        slot = P["k"] & 1
        harena, hv = P["host"][slot]
        if P["used"][slot]:
            P["h2d"][slot].synchronize()            # the slot's previous host -> device copy (two calls ago) has left the host buffer
        hv["img_l"].copy_(image_l)
        hv["img_u"].copy_(image_u)
        hv["label_l"].copy_(label_l)
        hv["label_u"].copy_(label_u)
        d = self.draw_host() if draws == "auto" else draws
        if d is not None:
            lam_l, idx_l, lam_u, idx_u = d
            hv["lam"].copy_(torch.tensor([lam_l, 1 - lam_l, lam_u, 1 - lam_u], dtype=torch.float64).float())
            hv["idx_l"].copy_(idx_l)
            hv["idx_u"].copy_(idx_u)               # (--om: recomputed on the device inside the step)
            hv["s_lab"].copy_(hv["label_l"][idx_l])          # smoothed_disc_label = disc_label[index] (mixup.py:37)
        main = torch.cuda.current_stream()
        with torch.cuda.stream(P["copy"]):
            if P["used"][slot]:
                P["copy"].wait_event(P["free"][slot])        # the device slot was consumed by the step two calls ago
            P["dev"][slot].copy_(harena, non_blocking=True)
            P["h2d"][slot].record(P["copy"])
        main.wait_event(P["h2d"][slot])
        self._in_arena.copy_(P["dev"][slot], non_blocking=True)
        P["free"][slot].record(main)
        P["used"][slot] = True
        self.run_resident()
        P["terms"][slot].copy_(self.terms, non_blocking=True)
        P["done"][slot].record(main)
        prev, P["pending"] = P["pending"], slot
        P["k"] += 1
        return None if prev is None else self._collect(prev)

    def drain(self):
        """loss terms of the last step_async() call (waits for it); None when nothing is pending"""
        P = self._pipe
        if P is None or P["pending"] is None:
            return None
        slot, P["pending"] = P["pending"], None
        return self._collect(slot)

    def _collect(self, slot):
        P = self._pipe
        P["done"][slot].synchronize()
        return self._terms_dict(P["terms"][slot])

    def h2d_bytes_async(self):
        return int(self._in_arena.numel())

    def read_terms(self):
        self.h_terms.copy_(self.terms, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self._terms_dict(self.h_terms)

    def _terms_dict(self, h_terms):
        t = {k: float(h_terms[i]) for i, k in enumerate(TERM_NAMES)}
        s = self.sched
        for sfx in ("l", "u"):
            t["prior_" + sfx] = s["kbc"] * abs(t["klc_" + sfx] - s["cmi"]) + s["kbd"] * abs(t["kld_" + sfx] - s["dmi"])
        return t

    # ---- optimizer state (reference checkpoint field 'optimizer', main_shot_vae.py:209,241) ------------------
    def state_dict(self):
        """The fused optimizer's state in torch.optim.SGD's state_dict layout (parameters numbered in
        model.parameters() order, one 'momentum_buffer' each), so the reference's checkpoint dict
        {'epoch', 'args', 'state_dict', 'optimizer'} can carry it and torch.optim.SGD.load_state_dict accepts it."""
        net, h = self.net, self.h
        n = len(net.pnames)
        state = {}
        if self._steps_done > 0:
            for i, k in enumerate(net.pnames):
                o, cnt, shp = net.poff[k]
                state[i] = {"momentum_buffer": net.momentum[o:o + cnt].view(shp).clone()}
        group = dict(lr=self._lr, momentum=h["momentum"], dampening=0, weight_decay=h["wd"], nesterov=False, maximize=False,
                     foreach=None, differentiable=False, fused=None, params=list(range(n)))
        return {"state": state, "param_groups": [group], "shotvae": {"steps_done": self._steps_done}}

    def load_state_dict(self, sd):
        """accepts state_dict() above or a plain torch.optim.SGD state dict (a reference checkpoint's 'optimizer')"""
        net = self.net
        groups = sd["param_groups"]
        assert len(groups) == 1 and len(groups[0]["params"]) == len(net.pnames), "optimizer state does not match the model"
        state = sd.get("state", {})
        net.momentum.zero_()
        have = 0
        for i, k in enumerate(net.pnames):
            ent = state.get(i, state.get(str(i)))
            buf = None if ent is None else ent.get("momentum_buffer")
            if buf is not None:
                o, cnt, shp = net.poff[k]
                net.momentum[o:o + cnt].view(shp).copy_(buf.to(net.device, torch.float32))
                have += 1
        assert have in (0, len(net.pnames)), "partial momentum state (%d of %d buffers)" % (have, len(net.pnames))
        # torch semantics: a parameter without a momentum buffer takes buf = grad on its next step
        self._steps_done = int(sd.get("shotvae", {}).get("steps_done", 1 if have else 0))
        if have == 0:
            self._steps_done = 0
        self.h["momentum"], self.h["wd"] = float(groups[0]["momentum"]), float(groups[0]["weight_decay"])
        self.set_lr(float(groups[0]["lr"]))

    def set_noise(self, eps4, unif2):
        """parity mode: host-drawn noise in pass order (P1, P2, P3, P4) / (P3, P4)"""
        self.eps.copy_(eps4)
        self.unif.copy_(unif2)
