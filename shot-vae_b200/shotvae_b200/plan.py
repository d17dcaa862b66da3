"""Network plan + executor: the sequence of libshotvae kernels that realises the VAE forward /
backward passes of the SHOT-VAE hot path (reference shot_vae_model/{wideresnet,preactresnet,decoder,vae}.py).

Design:
  * parameters live in ONE flat FP32 arena (reference state_dict layouts: Conv OIHW, ConvT IOHW, BN
    weight/bias, Linear [out,in]); gradients and SGD momentum are sibling arenas.  nn.Parameters of the
    drop-in modules are views into it, the data-parallel all-reduce and the fused SGD run on the flat
    buffers.
  * activations are NHWC bf16; `G` independent passes (each with its own BatchNorm statistics) are
    batched into one launch sequence (NB = G*B images).
  * every conv / convT is an implicit GEMM described by a tap table (sv_igemm_fprop / sv_igemm_wgrad);
    stride-2 input gradients and transposed convolutions are decomposed into output-parity phases.
"""
import ctypes as C
import os
from ctypes import byref
import re
from collections import OrderedDict

import torch

from . import _abi
from ._abi import lib, check, ptr, IgemmArgs, WgradArgs, BnBwdTerm, taps_array

BN_EPS = 1e-5
BN_MOMENTUM = 0.1
# SHOTVAE_FUSE_BNBWD=0: BatchNorm-backward statistics by the separate sv_bn_bwd_reduce pass everywhere (A/B switch)
FUSE_BN_BWD = os.environ.get("SHOTVAE_FUSE_BNBWD", "1") != "0"
BWD_DECOUPLE = os.environ.get("SHOTVAE_BWD_DECOUPLE", "0") != "0"      # MEASURED: 4.69-4.71 vs 4.66-4.71 ms/step (C2), 24.2 vs 24.4 (C4) -- no gain, off
WG_WORKSPACE_FLOATS = 24 * 1024 * 1024          # cap of one weight tensor's partial-sum workspace
# queued reductions are issued once their partial sums exceed this (about half of the 126 MB L2: the reduction should still hit)
WG_FLUSH_BYTES = int(os.environ.get("SHOTVAE_WG_FLUSH_MB", "64")) * (1 << 20)


def pad16(c):
    return (c + 15) // 16 * 16


# ------------------------------------------------------------------------------------ topology
class Unit:
    def __init__(self, prefix, cin, cout, stride, shortcut):
        self.prefix, self.cin, self.cout, self.stride, self.shortcut = prefix, cin, cout, stride, shortcut


def encoder_topology(name):
    """Unit list of the supported encoders (wideresnet.py:68-99, preactresnet.py:85-117)."""
    if "wideresnet" in name:
        nums = re.findall(r"\d+", name)
        depth, width = int(nums[0]), int(nums[1])
        if (depth - 4) % 6 != 0:
            raise AssertionError("depth should be 6n+4")
        n_unit = (depth - 4) // 6
        units, cin = [], 16
        for bi, base in enumerate((16, 32, 64)):
            cout = int(base * width)
            for ui in range(n_unit):
                stride = 2 if (bi > 0 and ui == 0) else 1
                c_in = cin if ui == 0 else cout
                units.append(Unit("feature_extractor.encoder.wideblock%d.wide_block.wideunit%d" % (bi + 1, ui + 1),
                                  c_in, cout, stride, c_in != cout or stride != 1))
            cin = cout
        return dict(f0=16, units=units, slope=0.01, shortcut_slope=0.01, feat=cin)
    if name == "preactresnet18":
        units, cin = [], 64
        for bi, cout in enumerate((64, 128, 256, 512)):
            for ui in range(2):
                stride = 2 if (bi > 0 and ui == 0) else 1
                c_in = cin if ui == 0 else cout
                units.append(Unit("feature_extractor.encoder.block%d.preact_block.unit%d" % (bi + 1, ui + 1),
                                  c_in, cout, stride, c_in != cout or stride != 1))
            cin = cout
        # the PreAct shortcut is BN -> conv1x1 with NO activation: slope 1.0 makes act() the identity
        return dict(f0=64, units=units, slope=0.0, shortcut_slope=1.0, feat=512)
    raise NotImplementedError("{} not implemented".format(name))


DEC_CHANNELS = (1024, 512, 256, 128, 64)


def conv_taps(k, pad):
    return [(ky * k + kx, ky - pad, kx - pad) for ky in range(k) for kx in range(k)]


def dgrad_phase_taps(k, s, pad):
    """Input-gradient of a (k, stride s, pad) convolution == forward of the transposed convolution,
    split by output parity: {(py, px): [(tap_index, off_y, off_x)]}; out[s*j+py] += in[j+off] * w[tap]."""
    res = {}
    for py in range(s):
        for px in range(s):
            taps = []
            for ky in range(k):
                if (py + pad - ky) % s:
                    continue
                for kx in range(k):
                    if (px + pad - kx) % s:
                        continue
                    taps.append((ky * k + kx, (py + pad - ky) // s, (px + pad - kx) // s))
            res[(py, px)] = taps
    return res


def _valid_pairs(taps, NB, OH, OW, H, W, in_stride, exact):
    """number of (output row, tap) pairs that contribute a MAC per (n, c): all of them for the conv
    counting rule (padding taps count), only in-image ones for the transposed-conv
    'effective' rule"""
    if not exact:
        return float(NB) * OH * OW * len(taps)
    tot = 0
    for _, dy, dx in taps:
        ny = sum(1 for o in range(OH) if 0 <= o * in_stride + dy < H)
        nx = sum(1 for o in range(OW) if 0 <= o * in_stride + dx < W)
        tot += ny * nx
    return float(NB) * tot


def live_taps(taps, OH, OW, H, W, in_stride):
    """Drops taps that read outside the image for every output position (e.g. 12 of the 16 taps of the
    1x1 -> 2x2 decoder layer)."""
    def any_in(n_out, n_in, d):
        return any(0 <= o * in_stride + d < n_in for o in range(n_out))
    return [t for t in taps if any_in(OH, H, t[1]) and any_in(OW, W, t[2])]


# ------------------------------------------------------------------------------------ context
class Ctx:
    """Per-pass buffers (activations saved for backward, BN statistics, gradients).  Buffers are
    allocated on first use and then reused, so every later pass replays the same addresses
    (CUDA-graph capturable)."""

    def __init__(self, net, G, B, parent=None, g0=0):
        """parent/g0: this context is a VIEW of pass groups [g0, g0+G) of `parent` -- every buffer is a slice
        of the parent's buffer of the same name.  Two G=2 forward passes can then be followed by ONE G=4
        backward pass over the parent (engine.TrainStep)."""
        self.net, self.G, self.B, self.NB = net, G, B, G * B
        self.dev = net.device
        self.parent, self.g0 = parent, g0
        self.bufs = {}
        self.args = {}
        if parent is None:
            self.zero_arena = torch.zeros(3 * 1024 * 1024, dtype=torch.float32, device=self.dev)
        self.zero_used = 0
        self.bn = OrderedDict()   # bn name -> dict(mean, var, count, ...)
        self.tape = []

    def t(self, name, shape, dtype=None):
        """dtype None = the network's activation type (bf16; fp32 in the parity-grade mode)"""
        if dtype is None:
            dtype = self.net.adt
        b = self.bufs.get(name)
        if b is None:
            if self.parent is not None:
                if shape[0] == self.NB:        # per-image tensor: slice the images of this view's groups
                    full = self.parent.t(name, (self.parent.NB,) + tuple(shape[1:]), dtype)
                    b = full[self.g0 * self.B:(self.g0 + self.G) * self.B]
                else:                          # per-group tensor ([G][C] BatchNorm coefficients)
                    assert shape[0] == self.G, (name, shape)
                    full = self.parent.t(name, (self.parent.G,) + tuple(shape[1:]), dtype)
                    b = full[self.g0:self.g0 + self.G]
            else:
                b = torch.empty(shape, dtype=dtype, device=self.dev)
            self.bufs[name] = b
        return b

    def z(self, name, numel):
        """fp32 buffer that is zero at the start of every pass (one memset for all of them)."""
        b = self.bufs.get(name)
        if b is None and self.parent is not None:
            per = numel // self.G
            assert per * self.G == numel, (name, numel)
            full = self.parent.z(name, per * self.parent.G)
            b = full[self.g0 * per:(self.g0 + self.G) * per]
            self.bufs[name] = b
        if b is None:
            n = (numel + 3) // 4 * 4
            assert self.zero_used + n <= self.zero_arena.numel(), "zero arena exhausted"
            b = self.zero_arena[self.zero_used:self.zero_used + numel]
            self.zero_used += n
            self.bufs[name] = b
        return b

    def reset(self):
        if self.parent is not None:
            return              # the parent's arena covers the views
        n = self.zero_used if self.zero_used else self.zero_arena.numel()
        check(lib.sv_fill_zero(ptr(self.zero_arena), n * 4, _abi.stream()))


# ------------------------------------------------------------------------------------ the net
class Net:
    def __init__(self, encoder_name, nd, ldc, in_ch, named_params, named_buffers, device, temperature=0.67, impl=0,
                 precision="bf16"):
        """precision: "bf16" = production (bf16 NHWC activations and conv operands on the tcgen05 kernels, fp32 accumulation /
        statistics / losses / optimizer); "fp32" = parity-grade mode (every activation tensor and conv operand fp32, the
        FP32 kernels of csrc/igemm_f32.cu and the `_f32` elementwise entry points; same launch sequence)."""
        assert precision in ("bf16", "fp32"), precision
        self.precision, self.f32 = precision, precision == "fp32"
        self.adt = torch.float32 if self.f32 else torch.bfloat16
        self.encoder_name, self.nd, self.ldc, self.in_ch = encoder_name, nd, ldc, in_ch
        self.device = torch.device(device)
        self.topo = encoder_topology(encoder_name)
        self.temperature = float(temperature)
        self.impl = impl
        self.latent = ldc + nd
        # ---- flat arenas
        self.pnames = list(named_params.keys())
        self.poff, off = {}, 0
        for k, v in named_params.items():
            self.poff[k] = (off, v.numel(), tuple(v.shape))
            off += v.numel()
        self.n_params = off
        n_alloc = (off + 3) // 4 * 4
        self.params = torch.zeros(n_alloc, dtype=torch.float32, device=self.device)
        self.grads = torch.zeros(n_alloc, dtype=torch.float32, device=self.device)
        self.momentum = torch.zeros(n_alloc, dtype=torch.float32, device=self.device)
        for k, v in named_params.items():
            self.p(k).copy_(v.detach().to(self.device, torch.float32))
        self.boff, off = {}, 0
        for k, v in named_buffers.items():
            if v.dtype == torch.float32:
                self.boff[k] = (off, v.numel())
                off += v.numel()
        self.running = torch.zeros(max(off, 1), dtype=torch.float32, device=self.device)
        self.nbt_names = [k for k, v in named_buffers.items() if v.dtype != torch.float32]
        self.nbt = torch.zeros(max(len(self.nbt_names), 1), dtype=torch.int64, device=self.device)
        for k, v in named_buffers.items():
            self.b(k).copy_(v.detach().to(self.device))
        self._wg_pending, self._wg_pending_bytes = [], 0     # queued weight-gradient reductions (_wgrad / _wgrad_flush)
        self.param_epoch = 0      # advanced by whoever rewrites the FP32 masters behind torch's back (TrainStep's fused SGD)
        self.dry = False          # dry mode: build buffers / records only, launch nothing (Ctx views -> parent tape)
        self.eval_bn = False      # True: BatchNorm normalises with the running statistics (model.eval(), reference valid()/test())
        self.side = None          # optional torch.cuda.Stream for the weight gradients (set by the engine)
        self.side_dec = None      # optional second stream: the decoder's weight gradients (decoder_bwd)
        self.timing = None        # bench instrumentation: list of (kind, key, flops, ev0, ev1, algorithmic bytes) when enabled
        self._build_packs()

    def fn(self, name):
        """the C-ABI entry point for this network's activation type: `name` (bf16) or its `_f32` twin"""
        if not self.f32:
            return getattr(lib, name)
        return getattr(lib, {"sv_colsum_bf16": "sv_colsum_f32"}.get(name, name + "_f32"))

    # ---- views
    def p(self, name):
        o, n, shp = self.poff[name]
        return self.params[o:o + n].view(shp)

    def g(self, name):
        o, n, shp = self.poff[name]
        return self.grads[o:o + n].view(shp)

    def b(self, name):
        if name in self.boff:
            o, n = self.boff[name]
            return self.running[o:o + n]
        i = self.nbt_names.index(name)
        return self.nbt[i:i + 1].view(())

    # ---- weight packing recipes ------------------------------------------------------------
    def _add_pack(self, key, wname, N, C, taps, n_real, c_real, sn, sc, st, grid=None):
        """grid = (H, W) of the stride-1 row grid the pack is used on (None: strided gather).  The library
        is asked whether the halo-tile tcgen05 kernel covers the problem; if so the weights are packed in
        its 8-channel-plane layout."""
        T = len(taps)
        dst = torch.zeros(T, N, C, dtype=self.adt, device=self.device)
        layout = 2 if self.f32 else 0       # 2: fp32 [T][N][C] operands of the FP32 kernels
        if not self.f32 and grid is not None and self.impl in (0, 3) and os.environ.get("SHOTVAE_HALO", "1") != "0":      # (A/B switch: 0 = per-tap TMA kernel everywhere)
            a = IgemmArgs()
            a.A = a.Wt = a.out_bf16 = ptr(dst)
            a.NB, a.H, a.W, a.C, a.OH, a.OW, a.N, a.T = 128, grid[0], grid[1], C, grid[0], grid[1], N, T
            a.in_stride, a.out_stride, a.OHf, a.OWf, a.group_images = 1, 1, grid[0], grid[1], 128
            a.dy, a.dx = taps_array([t[1] for t in taps]), taps_array([t[2] for t in taps])
            a.w_layout = 1
            layout = 1 if lib.sv_igemm_fprop_supports(byref(a), 3) else 0
        self.packs[key] = dict(w=dst, taps=taps, wname=wname, N=N, C=C, T=T, n_real=n_real, c_real=c_real,
                               sn=sn, sc=sc, st=st, tidx=taps_array([t[0] for t in taps]), layout=layout)

    def _build_packs(self):
        self.packs = OrderedDict()
        topo = self.topo
        f0 = topo["f0"]
        cin_p = pad16(self.in_ch)
        # conv0: Conv2d(in_ch, f0, 3, 1, 1, bias) -- OIHW
        self._add_pack("conv0.f", "feature_extractor.encoder.pre_process.conv0.weight", f0, cin_p, conv_taps(3, 1),
                       f0, self.in_ch, self.in_ch * 9, 9, 1, grid=(32, 32))
        H = 32
        for ui, u in enumerate(topo["units"]):
            Ho = H // u.stride
            for cname, ci, co, k, s, pad, hin in (("conv1", u.cin, u.cout, 3, u.stride, 1, H),
                                                 ("conv2", u.cout, u.cout, 3, 1, 1, Ho),
                                                 ("sc", u.cin, u.cout, 1, u.stride, 0, H)):
                if cname == "sc" and not u.shortcut:
                    continue
                wname = u.prefix + (".i_block.conv.weight" if cname == "sc" else ".f_block.%s.weight" % cname)
                K = k * k
                hout = hin // s
                # 128 -> 128 channels at 8x8: the halo kernel has to split N into two 64-column halves (295 KB of weights) and
                # half of every 128-slot tile is padding; the per-tap kernel with whole-N tiles is faster there
                # (MEASURED at NB = 256: 16.6 vs 23.1 us plain, 19.1 vs 24.0 us with residual + statistics)
                big = ci >= 128 and co >= 128 and s == 1 and k == 3 and os.environ.get("SHOTVAE_B3", "tc") == "tc"
                self._add_pack("u%d.%s.f" % (ui, cname), wname, co, ci, conv_taps(k, pad), co, ci, ci * K, K, 1,
                               grid=(hin, hin) if (s == 1 and not big) else None)
                for (py, px), taps in dgrad_phase_taps(k, s, pad).items():
                    taps = live_taps(taps, hout, hout, hout, hout, 1)
                    if taps:
                        self._add_pack("u%d.%s.d%d%d" % (ui, cname, py, px), wname, ci, co, taps, ci, co, K, ci * K, 1,
                                       grid=None if big else (hout, hout))
            H = Ho
        # decoder: ConvTranspose2d weights are [Cin, Cout, k, k]
        cin, hin = DEC_CHANNELS[0], 1
        for li, cout in enumerate(DEC_CHANNELS[1:] + (self.in_ch,)):
            wname = "feature_reconstructor.decoder.%d.weight" % (3 * (li + 1))
            cout_p = pad16(cout)
            for (py, px), taps in dgrad_phase_taps(4, 2, 1).items():
                taps = live_taps(taps, hin, hin, hin, hin, 1)
                self._add_pack("d%d.f%d%d" % (li + 1, py, px), wname, cout_p, cin, taps, cout, cin, 16, cout * 16, 1,
                               grid=(hin, hin) if li < 4 else None)
            taps = live_taps(conv_taps(4, 1), hin, hin, 2 * hin, 2 * hin, 2)
            self._add_pack("d%d.d" % (li + 1), wname, cin, cout_p, taps, cin, cout, cout * 16, 16, 1)
            cin, hin = cout, hin * 2

    def pack_weights(self):
        """FP32 master weights -> bf16 operand layouts (once per optimizer step), one launch for all."""
        if getattr(self, "_pack_table", None) is None:
            n = len(self.packs)
            arr = (_abi.PackDesc * n)()
            for i, pk in enumerate(self.packs.values()):
                d = arr[i]
                d.src, d.dst = ptr(self.p(pk["wname"])), ptr(pk["w"])
                d.sn, d.sc, d.st = pk["sn"], pk["sc"], pk["st"]
                d.N, d.C, d.T, d.n_real, d.c_real, d.layout = pk["N"], pk["C"], pk["T"], pk["n_real"], pk["c_real"], pk["layout"]
                for k in range(pk["T"]):
                    d.tap[k] = pk["taps"][k][0]
            raw = bytes(arr)
            self._pack_table = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(self.device)
            self._pack_n = n
        check(lib.sv_pack_weights_batched(ptr(self._pack_table), self._pack_n, 592, _abi.stream()))

    # ---- low-level launch helpers --------------------------------------------------------------
    def _igemm(self, ctx, key, A, pack, NB, H, W, OH, OW, in_stride=1, out=None, outf=None, res=None, bias=None, stats=None,
               out_stride=1, off=(0, 0), OHf=None, OWf=None, n_valid=0, exact=False, batch=None, bnb=None):
        """batch: a list -> the launch is deferred; _igemm_flush(batch) issues all of them through sv_igemm_fprop_batch
        (independent problems, e.g. the four output-parity phases of a transposed convolution).
        bnb = dict(rec, y, slope): input-gradient launch whose epilogue also accumulates the BatchNorm-backward statistics of
        BatchNorm `rec` (input y) into `stats` ([2][G][N]: dbeta block, dgamma block); returns False (nothing launched, nothing
        changed) when the kernel that runs this shape has no such epilogue, so that the caller can fall back to
        sv_bn_bwd_reduce."""
        if self.f32 and bnb is not None:
            return False                     # the FP32 kernel has no fused-statistics epilogue: the caller runs sv_bn_bwd_reduce
        a = ctx.args.get(key)
        if a is None:
            pk = self.packs[pack]
            a = IgemmArgs()
            a.Wt = ptr(pk["w"])
            a.NB, a.H, a.W, a.C = NB, H, W, pk["C"]
            a.OH, a.OW, a.N, a.T = OH, OW, pk["N"], pk["T"]
            a.in_stride, a.out_stride, a.out_off_y, a.out_off_x = in_stride, out_stride, off[0], off[1]
            a.OHf = OHf if OHf is not None else OH * out_stride
            a.OWf = OWf if OWf is not None else OW * out_stride
            a.n_valid, a.group_images = n_valid, ctx.B
            a.dy, a.dx = taps_array([t[1] for t in pk["taps"]]), taps_array([t[2] for t in pk["taps"]])
            a.w_layout = pk["layout"]
            assert A.shape[-1] == pk["C"], (key, A.shape, pk["C"])
            ctx.args[key] = a
        # operand pointers are refreshed on every call (callers may hand in different tensors)
        a.A, a.out_bf16, a.out_f32, a.residual, a.bias, a.stats = ptr(A), ptr(out), ptr(outf), ptr(res), ptr(bias), ptr(stats)
        a.impl = 3 if a.w_layout == 1 else (self.impl if self.impl != 3 else 0)
        if self.f32:                         # every tensor is fp32: the output goes through out_f32 (row length N unless n_valid)
            a.out_bf16, a.out_f32, a.impl = None, ptr(out if out is not None else outf), 4
        if bnb is not None:
            rec = bnb["rec"]
            a.bn_y, a.bn_scale, a.bn_shift, a.bn_mean, a.bn_var = ptr(bnb["y"]), ptr(rec["scale"]), ptr(rec["shift"]), ptr(rec["mean"]), ptr(rec["var"])
            a.bn_slope, a.bn_eps = float(bnb["slope"]), BN_EPS
            ok = ctx.args.get(key + "#bnb")
            if ok is None:
                ok = ctx.args[key + "#bnb"] = bool(FUSE_BN_BWD and lib.sv_igemm_fprop_supports(C.byref(a), a.impl))
            if not ok:
                a.bn_y = None
                return False
        else:
            a.bn_y = None
        if self.dry:
            return True
        if batch is not None and self.timing is None:
            batch.append(a)
            return True
        if self.timing is None:
            check(lib.sv_igemm_fprop(C.byref(a), _abi.stream()))
            return True
        pk = self.packs[pack]
        flops = 2.0 * pk["n_real"] * pk["c_real"] * _valid_pairs(pk["taps"], NB, OH, OW, H, W, in_stride, exact)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(lib.sv_igemm_fprop(C.byref(a), _abi.stream()))
        e1.record()
        esz = lambda t: 0 if t is None else t.numel() * t.element_size()
        nbytes = esz(A) // (in_stride * in_stride) + esz(pk["w"]) + esz(res) + NB * OH * OW * pk["N"] * ((2 if out is not None else 0) +
                                                                                                  (4 if outf is not None else 0))
        self.timing.append(("igemm_fprop", key, flops, e0, e1, nbytes))
        return True

    def _igemm_flush(self, batch):
        if not batch:
            return
        arr = (IgemmArgs * len(batch))(*batch)
        check(lib.sv_igemm_fprop_batch(arr, len(batch), _abi.stream()))
        del batch[:]

    def _wgrad(self, ctx, key, A, Gr, taps, NB, H, W, Cc, OH, OW, N, in_stride, wname, n_real, c_real, sn, sc, st, exact=False):
        """weight-gradient GEMM into this tensor's OWN partial-sum workspace; the sum over the partial slices into the gradient
        arena is queued and issued by _wgrad_flush() for several tensors in one launch (a shared workspace forced one small
        reduction launch behind every weight gradient: 33 launches, 0.5 ms per C2 step, ~80 % of it launch latency)."""
        ent = ctx.args.get(key)
        if ent is None:
            T = len(taps)
            bmn = 128 if N % 128 == 0 else 64 if N % 64 == 0 else 32 if N % 32 == 0 else 16
            tiles = ((T * Cc + 127) // 128) * ((N + bmn - 1) // bmn)
            M = NB * OH * OW
            if self.f32:                     # 64 x 64 output tiles of the FP32 kernel
                tiles = ((T * Cc + 63) // 64) * ((N + 63) // 64)
            splits = max(1, min((296 + tiles - 1) // tiles, max(1, M // 256)))
            while splits > 1 and splits * N * T * Cc > WG_WORKSPACE_FLOATS:
                splits -= 1
            assert splits * N * T * Cc <= WG_WORKSPACE_FLOATS, "wgrad workspace too small for %s" % key
            a = WgradArgs()
            a.NB, a.H, a.W, a.C, a.OH, a.OW, a.N, a.T = NB, H, W, Cc, OH, OW, N, T
            a.in_stride, a.splits = in_stride, splits
            a.dy, a.dx = taps_array([t[1] for t in taps]), taps_array([t[2] for t in taps])
            a.impl = 4 if self.f32 else (1 if self.impl == 1 else 0)
            a.A, a.Gr = ptr(A), ptr(Gr)
            a.partial = ptr(A)               # (any non-null pointer: the workspace is sized from the answer)
            tc_splits = 0 if self.f32 else lib.sv_igemm_wgrad_splits(byref(a))     # > 0: the tcgen05 kernel runs it, one slice per CTA
            if tc_splits > 0:
                splits = a.splits = tc_splits
                assert splits * N * T * Cc <= WG_WORKSPACE_FLOATS, "wgrad workspace too small for %s" % key
            ws = torch.empty(splits * N * T * Cc, dtype=torch.float32, device=self.device)
            a.partial = ptr(ws)
            d = _abi.ReduceDesc()
            d.partial, d.grad = ptr(ws), ptr(self.g(wname))
            d.sn, d.sc, d.st = sn, sc, st
            d.splits, d.N, d.C, d.T, d.n_real, d.c_real = splits, N, Cc, T, n_real, c_real
            for i, t in enumerate(taps):
                d.tap_index[i] = t[0]
            ent = (a, d, ws, splits, T)
            ctx.args[key] = ent
        a, d, ws, splits, T = ent
        a.A, a.Gr = ptr(A), ptr(Gr)
        s = _abi.stream()
        nbytes_red = splits * N * T * Cc * 4 + n_real * c_real * T * 8
        if self.timing is not None:
            flops = 2.0 * n_real * c_real * _valid_pairs(taps, NB, OH, OW, H, W, in_stride, exact)
            e0, e1 = (torch.cuda.Event(enable_timing=True) for _ in range(2))
            e0.record()
            check(lib.sv_igemm_wgrad(C.byref(a), s))
            e1.record()
            nbytes = A.numel() * 2 // (in_stride * in_stride) + Gr.numel() * 2 + splits * N * T * Cc * 4
            self.timing.append(("igemm_wgrad", key, flops, e0, e1, nbytes))
        else:
            check(lib.sv_igemm_wgrad(C.byref(a), s))
        self._wg_pending.append((key, d, nbytes_red))
        self._wg_pending_bytes += nbytes_red
        if self._wg_pending_bytes > WG_FLUSH_BYTES:
            self._wgrad_flush()

    def _wgrad_flush(self):
        """sum the queued partial-sum workspaces into the gradient arena: one launch (current stream = the stream the queued
        weight gradients were issued on)"""
        pend = self._wg_pending
        if not pend:
            return
        arr = (_abi.ReduceDesc * len(pend))(*[d for _, d, _ in pend])
        s = _abi.stream()
        if self.timing is not None:
            e0, e1 = (torch.cuda.Event(enable_timing=True) for _ in range(2))
            e0.record()
            check(lib.sv_wgrad_reduce_batched(arr, len(pend), s))
            e1.record()
            self.timing.append(("wgrad_reduce", "+".join(k for k, _, _ in pend), 0.0, e0, e1, sum(b for _, _, b in pend)))
        else:
            check(lib.sv_wgrad_reduce_batched(arr, len(pend), s))
        del pend[:]
        self._wg_pending_bytes = 0

    def _bn_fwd(self, ctx, bn_name, key, stats, Cc, count):
        """finalize batch statistics -> (scale, shift, mean, var), all [G][C]"""
        G = ctx.G
        rec = dict(mean=ctx.t(key + ".mean", (G, Cc), torch.float32), var=ctx.t(key + ".var", (G, Cc), torch.float32),
                   scale=ctx.t(key + ".scale", (G, Cc), torch.float32), shift=ctx.t(key + ".shift", (G, Cc), torch.float32),
                   count=count, C=Cc, name=bn_name)
        ctx.bn[bn_name] = rec
        if self.dry:
            return rec
        if self.eval_bn:
            self._bn_eval_coeffs(bn_name, rec)
            return rec
        check(lib.sv_bn_finalize(ptr(stats), ptr(self.p(bn_name + ".weight")), ptr(self.p(bn_name + ".bias")), float(count),
                                 BN_EPS, G, Cc, Cc, ptr(rec["mean"]), ptr(rec["var"]), ptr(rec["scale"]), ptr(rec["shift"]),
                                 _abi.stream()))
        ctx.bn[bn_name] = rec
        return rec

    def _timed(self, kind, key, nbytes, launch):
        """bench instrumentation: CUDA events around one launch (only when self.timing is a list)"""
        if self.timing is None:
            launch()
            return
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        launch()
        e1.record()
        self.timing.append((kind, key, 0.0, e0, e1, nbytes))

    def _bn_fwd_act(self, ctx, bn_name, key, stats, Cc, count, y, a, slope):
        """BatchNorm finalize + apply + activation in one launch"""
        G = ctx.G
        rec = dict(mean=ctx.t(key + ".mean", (G, Cc), torch.float32), var=ctx.t(key + ".var", (G, Cc), torch.float32),
                   scale=ctx.t(key + ".scale", (G, Cc), torch.float32), shift=ctx.t(key + ".shift", (G, Cc), torch.float32),
                   count=count, C=Cc, name=bn_name)
        ctx.bn[bn_name] = rec
        if self.dry:
            return rec
        if self.eval_bn:
            self._bn_eval_coeffs(bn_name, rec)
            self._bn_act(ctx, y, a, rec, slope, count)
            return rec
        self._timed("bn_fwd_act", key, 4 * y.numel(), lambda: check(self.fn("sv_bn_finalize_act_fwd")(
            ptr(y), ptr(a), ptr(stats), ptr(self.p(bn_name + ".weight")), ptr(self.p(bn_name + ".bias")), float(count), BN_EPS,
            float(slope), count, G, Cc, ptr(rec["mean"]), ptr(rec["var"]), ptr(rec["scale"]), ptr(rec["shift"]), _abi.stream())))
        ctx.bn[bn_name] = rec
        return rec

    def _bn_eval_coeffs(self, bn_name, rec):
        """eval mode (nn.BatchNorm2d with training=False): scale = gamma / sqrt(running_var + eps), shift = beta -
        running_mean * scale, the same for every pass group.  Tiny per-channel tensors, outside the training hot path."""
        Cr = self.p(bn_name + ".weight").numel()
        rm, rv = self.b(bn_name + ".running_mean"), self.b(bn_name + ".running_var")
        sc = self.p(bn_name + ".weight") * torch.rsqrt(rv + BN_EPS)
        for k, v in (("scale", sc), ("shift", self.p(bn_name + ".bias") - rm * sc), ("mean", rm), ("var", rv)):
            rec[k].zero_()
            rec[k][:, :Cr] = v

    def _bn_act(self, ctx, y, a, rec, slope, rows_per_group):
        check(self.fn("sv_bn_act_fwd")(ptr(y), ptr(a), ptr(rec["scale"]), ptr(rec["shift"]), float(slope), rows_per_group, ctx.G,
                                rec["C"], _abi.stream()))

    def _bn_bwd_stats(self, ctx, key, i, Cc):
        """(dbeta, dgamma) accumulators [G][C] of term i of BatchNorm `key`: one zero-arena buffer [2][G][C], the layout the
        fused input-gradient epilogue writes"""
        st = ctx.z("%s.bst%d" % (key, i), 2 * ctx.G * Cc)
        return st[:ctx.G * Cc], st[ctx.G * Cc:]

    def _bn_bwd(self, ctx, key, terms, y, addend, g_y, rows_per_group, HW):
        """terms: list of dict(rec, g_a | g_feat, slope).  Runs the dgamma/dbeta reductions and the apply."""
        G = ctx.G
        Cc = terms[0]["rec"]["C"]
        arr = (BnBwdTerm * len(terms))()
        s = _abi.stream()
        for i, t in enumerate(terms):
            rec = t["rec"]
            db, dg = self._bn_bwd_stats(ctx, key, i, Cc)
            gb = 2 * t["g_a"].numel() if t.get("g_a") is not None else 0
            if not t.get("fused"):           # (fused: the input-gradient conv's epilogue has already accumulated db / dg)
                self._timed("bn_bwd_reduce", key, gb + 2 * y.numel(), lambda: check(self.fn("sv_bn_bwd_reduce")(
                    ptr(t.get("g_a")), ptr(t.get("g_feat")), ptr(y), ptr(rec["scale"]), ptr(rec["shift"]), ptr(rec["mean"]), ptr(rec["var"]),
                    BN_EPS, float(t["slope"]), rows_per_group, HW, G, Cc, ptr(dg), ptr(db), s)))
            arr[i].g_a, arr[i].g_feat = ptr(t.get("g_a")), ptr(t.get("g_feat"))
            arr[i].scale, arr[i].shift, arr[i].mean, arr[i].var = ptr(rec["scale"]), ptr(rec["shift"]), ptr(rec["mean"]), ptr(rec["var"])
            arr[i].dgamma, arr[i].dbeta = ptr(dg), ptr(db)
            arr[i].grad_gamma, arr[i].grad_beta = ptr(self.g(rec["name"] + ".weight")), ptr(self.g(rec["name"] + ".bias"))
            arr[i].slope, arr[i].c_real = float(t["slope"]), Cc
        nb = 2 * y.numel() * (2 + sum(1 for t in terms if t.get("g_a") is not None) + (1 if addend is not None else 0))
        self._timed("bn_bwd_apply", key, nb, lambda: check(self.fn("sv_bn_bwd_apply")(arr, len(terms), ptr(y), ptr(addend), ptr(g_y), BN_EPS,
                                                                               rows_per_group, HW, G, Cc, s)))

    # ---- encoder -------------------------------------------------------------------------------
    def encoder_fwd(self, ctx, x_img):
        """x_img: bf16 NHWC [NB, 32, 32, pad16(in_ch)] -> feat fp32 [NB, feat]"""
        topo, NB, G, B = self.topo, ctx.NB, ctx.G, ctx.B
        slope, sslope = topo["slope"], topo["shortcut_slope"]
        H, f0 = 32, topo["f0"]
        ctx.x_img = x_img
        h = ctx.t("h0", (NB, H, H, f0))
        st = ctx.z("st.h0", G * 2 * f0)
        self._igemm(ctx, "conv0", x_img, "conv0.f", NB, H, H, H, H, out=h, stats=st,
                    bias=self.p("feature_extractor.encoder.pre_process.conv0.bias"))
        ctx.tape = []
        for ui, u in enumerate(topo["units"]):
            k = "u%d" % ui
            Ho = H // u.stride
            rows_in, rows_out = B * H * H, B * Ho * Ho
            a1 = ctx.t(k + ".a1", (NB, H, H, u.cin))
            bn1 = self._bn_fwd_act(ctx, u.prefix + ".f_block.norm1", k + ".bn1", st, u.cin, rows_in, h, a1, slope)
            y1 = ctx.t(k + ".y1", (NB, Ho, Ho, u.cout))
            st1 = ctx.z(k + ".st1", G * 2 * u.cout)
            self._igemm(ctx, k + ".conv1", a1, k + ".conv1.f", NB, H, H, Ho, Ho, in_stride=u.stride, out=y1, stats=st1)
            a2 = ctx.t(k + ".a2", (NB, Ho, Ho, u.cout))
            bn2 = self._bn_fwd_act(ctx, u.prefix + ".f_block.norm2", k + ".bn2", st1, u.cout, rows_out, y1, a2, slope)
            bns = a_s = None
            res = h
            if u.shortcut:
                a_s = ctx.t(k + ".as", (NB, H, H, u.cin))
                bns = self._bn_fwd_act(ctx, u.prefix + ".i_block.norm", k + ".bns", st, u.cin, rows_in, h, a_s, sslope)
                res = ctx.t(k + ".s", (NB, Ho, Ho, u.cout))
                self._igemm(ctx, k + ".sc", a_s, k + ".sc.f", NB, H, H, Ho, Ho, in_stride=u.stride, out=res)
            hn = ctx.t(k + ".h", (NB, Ho, Ho, u.cout))
            stn = ctx.z(k + ".sth", G * 2 * u.cout)
            self._igemm(ctx, k + ".conv2", a2, k + ".conv2.f", NB, Ho, Ho, Ho, Ho, out=hn, stats=stn, res=res)
            ctx.tape.append(dict(u=u, k=k, h_in=h, H=H, Ho=Ho, bn1=bn1, a1=a1, y1=y1, bn2=bn2, a2=a2, bns=bns, a_s=a_s))
            h, st, H = hn, stn, Ho
        Cf = topo["feat"]
        bnT = self._bn_fwd(ctx, "feature_extractor.encoder.transition.norm", "bnT", st, Cf, B * H * H)
        feat = ctx.t("feat", (NB, Cf), torch.float32)
        ctx.enc_out = dict(h=h, H=H, bnT=bnT)
        if not self.dry:
            check(self.fn("sv_bn_act_gap_fwd")(ptr(h), ptr(feat), ptr(bnT["scale"]), ptr(bnT["shift"]), float(slope), NB, H * H, Cf, B,
                                        _abi.stream()))
        return feat

    def bwd_segments(self):
        """unit index ranges [(lo, hi), ...] of the encoder's resolution blocks in BACKWARD order (last block first): the
        points at which a data-parallel run hands a finished gradient range to the all-reduce (ddp.GradReducer)"""
        units = self.topo["units"]
        starts = [i for i, u in enumerate(units) if i == 0 or u.stride == 2] + [len(units)]
        return [(starts[i], starts[i + 1]) for i in range(len(starts) - 2, -1, -1)]

    def segment_param_ranges(self):
        """gradient-arena ranges [(lo, hi), ...] that are final after each backward segment of bwd_segments(): the first
        one also carries the transition BatchNorm and the heads, the last one conv0"""
        split = self.poff["feature_reconstructor.decoder.0.weight"][0]
        units = self.topo["units"]
        first = lambda ui: min(o for k, (o, _, _) in self.poff.items() if k.startswith(units[ui].prefix + "."))
        segs, out, hi = self.bwd_segments(), [], split
        for i, (lo_u, _) in enumerate(segs):
            lo = 0 if i == len(segs) - 1 else first(lo_u)
            out.append((lo, hi))
            hi = lo
        return out

    def encoder_bwd(self, ctx, g_feat, seg=None, side="net"):
        """g_feat: fp32 [NB, feat].  Accumulates every encoder parameter gradient into the grad arena.
        seg = None: the whole backward; seg = i: only segment i of bwd_segments() (segments must be run in order; the state
        between them lives in ctx; every segment ends with the side stream joined, so it can be its own CUDA graph).
        side: the stream for the weight gradients ("net" = self.side)."""
        topo, NB, G, B = self.topo, ctx.NB, ctx.G, ctx.B
        slope, sslope = topo["slope"], topo["shortcut_slope"]
        segs = self.bwd_segments()
        if seg is None or seg == 0:
            eo = ctx.enc_out
            H, Cf = eo["H"], topo["feat"]
            g_h = ctx.t("g.h.%d.%d" % (H, Cf), (NB, H, H, Cf))
            self._bn_bwd(ctx, "bnT", [dict(rec=eo["bnT"], g_feat=g_feat, slope=slope)], eo["h"], None, g_h, B * H * H, H * H)
            flip = 0
        else:
            g_h, flip = ctx._bwd_state
        lo_u, hi_u = (0, len(ctx.tape)) if seg is None else segs[seg]
        last = seg is None or seg == len(segs) - 1
        # Weight gradients run on a side stream (when the engine provides one): they only need (activation, output
        # gradient), so they overlap the dgrad -> BatchNorm-backward chain of the main stream -- tensor-core bound
        # kernels next to HBM-bound ones.  Buffers shared between units by shape (g.y1.*, the two g.h flip buffers)
        # are protected by `side_done`: the main stream waits for the previous unit's weight gradients before it
        # overwrites what they read.
        side = (self.side if side == "net" else side) if not self.dry else None
        main = torch.cuda.current_stream() if side is not None else None
        side_done = None

        def ready():
            """marks the point on the main stream after which a weight gradient's inputs are complete"""
            if side is None:
                return None
            ev = torch.cuda.Event()
            ev.record(main)
            return ev

        def on_side(fn, ev):
            # issued AFTER the main stream's next input-gradient conv (so that conv gets the SMs first and the
            # weight gradient then shares them with the BatchNorm kernels), but dependent only on `ev`
            if side is None:
                fn()
                return
            side.wait_event(ev)
            with torch.cuda.stream(side):
                fn()

        def mark_side():
            if side is None:
                return None
            ev = torch.cuda.Event()
            ev.record(side)
            return ev

        for rec in reversed(ctx.tape[lo_u:hi_u]):
            u, k, Hin, Ho = rec["u"], rec["k"], rec["H"], rec["Ho"]
            rows_in, rows_out = B * Hin * Hin, B * Ho * Ho
            g_out = g_h
            K9 = 9
            # conv2: input gradient (main), weight gradient (side)
            ev = ready()
            g_a2 = ctx.t("g.a2.%d.%d" % (Ho, u.cout), (NB, Ho, Ho, u.cout))
            f2 = self._igemm(ctx, k + ".conv2.d", g_out, k + ".conv2.d00", NB, Ho, Ho, Ho, Ho, out=g_a2,
                             stats=ctx.z(k + ".bn2.bst0", 2 * G * u.cout), bnb=dict(rec=rec["bn2"], y=rec["y1"], slope=slope))
            if not f2:
                self._igemm(ctx, k + ".conv2.d", g_out, k + ".conv2.d00", NB, Ho, Ho, Ho, Ho, out=g_a2)
            on_side(lambda: self._wgrad(ctx, k + ".conv2.w", rec["a2"], g_out, conv_taps(3, 1), NB, Ho, Ho, u.cout, Ho, Ho, u.cout, 1,
                                        u.prefix + ".f_block.conv2.weight", u.cout, u.cout, u.cout * K9, K9, 1), ev)
            # g.y1 and the unit's input gradient are PER-UNIT buffers (round 2 late): with buffers shared by shape the main
            # stream had to wait for the previous unit's weight gradients before overwriting what they read, which locked the
            # dgrad -> BatchNorm-backward chain to the weight-gradient stream unit by unit (gaps of 25-50 us on the critical
            # chain wherever the weight gradient was the slower one).  SHOTVAE_BWD_DECOUPLE=0: shared buffers + the wait.
            g_y1 = ctx.t(("g.y1.%s" % k) if BWD_DECOUPLE else "g.y1.%d.%d" % (Ho, u.cout), (NB, Ho, Ho, u.cout))
            if side_done is not None and not BWD_DECOUPLE:
                main.wait_event(side_done)      # the previous unit's conv1 weight gradient has read g.y1 / its g_out
            self._bn_bwd(ctx, k + ".bn2", [dict(rec=rec["bn2"], g_a=g_a2, slope=slope, fused=f2)], rec["y1"], None, g_y1, rows_out, Ho * Ho)

            # conv1 (+ projection shortcut)
            def wgrads1():
                self._wgrad(ctx, k + ".conv1.w", rec["a1"], g_y1, conv_taps(3, 1), NB, Hin, Hin, u.cin, Ho, Ho, u.cout, u.stride,
                            u.prefix + ".f_block.conv1.weight", u.cout, u.cin, u.cin * K9, K9, 1)
                if u.shortcut:
                    self._wgrad(ctx, k + ".sc.w", rec["a_s"], g_out, conv_taps(1, 0), NB, Hin, Hin, u.cin, Ho, Ho, u.cout, u.stride,
                                u.prefix + ".i_block.conv.weight", u.cout, u.cin, u.cin, 1, 1)
            ev = ready()
            g_a1 = ctx.t("g.a1.%d.%d" % (Hin, u.cin), (NB, Hin, Hin, u.cin))
            f1 = self._dgrad(ctx, k + ".conv1", g_y1, g_a1, NB, Ho, Hin, u.stride,
                             bnb=dict(rec=rec["bn1"], y=rec["h_in"], slope=slope, stats=ctx.z(k + ".bn1.bst0", 2 * G * u.cin)))
            on_side(wgrads1, ev)
            side_done = mark_side()
            terms = [dict(rec=rec["bn1"], g_a=g_a1, slope=slope, fused=f1)]
            addend = g_out
            if u.shortcut:
                g_as = ctx.t("g.as.%d.%d" % (Hin, u.cin), (NB, Hin, Hin, u.cin))
                fs = self._dgrad(ctx, k + ".sc", g_out, g_as, NB, Ho, Hin, u.stride,
                                 bnb=dict(rec=rec["bns"], y=rec["h_in"], slope=sslope, stats=ctx.z(k + ".bn1.bst1", 2 * G * u.cin)))
                terms.append(dict(rec=rec["bns"], g_a=g_as, slope=sslope, fused=fs))
                addend = None
            flip ^= 1
            g_prev = ctx.t(("g.h.%s" % k) if BWD_DECOUPLE else "g.h.%d.%d.%d" % (Hin, u.cin, flip), (NB, Hin, Hin, u.cin))
            self._bn_bwd(ctx, k + ".bn1", terms, rec["h_in"], addend, g_prev, rows_in, Hin * Hin)
            g_h = g_prev
            ctx._bwd_name = ("g.h.%s" % k) if BWD_DECOUPLE else "g.h.%d.%d.%d" % (Hin, u.cin, flip)
        if side is not None:
            with torch.cuda.stream(side):
                self._wgrad_flush()          # the segment's gradients are complete when the side stream is joined
            main.wait_stream(side)
        else:
            self._wgrad_flush()
        ctx._bwd_state = (g_h, flip)
        if not last:
            return
        # conv0: weight + bias gradient (no input gradient: the input is data)
        f0 = topo["f0"]
        cin_p = pad16(self.in_ch)
        self._wgrad(ctx, "conv0.w", ctx.x_img, g_h, conv_taps(3, 1), NB, 32, 32, cin_p, 32, 32, f0, 1,
                    "feature_extractor.encoder.pre_process.conv0.weight", f0, self.in_ch, self.in_ch * 9, 9, 1)
        self._wgrad_flush()
        check(self.fn("sv_colsum_bf16")(ptr(g_h), ptr(self.g("feature_extractor.encoder.pre_process.conv0.bias")), NB * 32 * 32, f0, f0,
                                        _abi.stream()))

    def _dgrad(self, ctx, key, g_out, g_in, NB, Ho, Hin, stride, bnb=None):
        """input gradient of a conv by output-parity phases (stride 1: a single phase).  bnb (stride 1 only): fuse the
        BatchNorm-backward statistics of the BatchNorm in front of this conv into the launch; returns whether that happened."""
        s = stride
        phases = [(py, px) for py in range(s) for px in range(s) if ("%s.d%d%d" % (key, py, px)) in self.packs]
        if s == 1 and bnb is not None and len(phases) == 1:
            k0 = "%s.d00" % key
            if self._igemm(ctx, k0, g_out, k0, NB, Ho, Ho, Ho, Ho, out=g_in, OHf=Hin, OWf=Hin, stats=bnb["stats"], bnb=bnb):
                return True
        if len(phases) < s * s:
            check(lib.sv_fill_zero(ptr(g_in), g_in.numel() * g_in.element_size(), _abi.stream()))
        fused = False
        for py, px in phases:
            self._igemm(ctx, "%s.d%d%d" % (key, py, px), g_out, "%s.d%d%d" % (key, py, px), NB, Ho, Ho, Ho, Ho, out=g_in,
                        out_stride=s, off=(py, px), OHf=Hin, OWf=Hin)
        return fused

    # ---- heads + sample ------------------------------------------------------------------------
    HEADS = (("continuous_inference.mean", "mu"), ("continuous_inference.log_sigma", "ls"), ("disc_latent_inference", "logits"))

    def _heads_desc(self, outs=None, grads=None):
        """sv_heads record of the three inference heads (one launch for all of them in each direction)"""
        h = _abi.Heads()
        h.n = len(self.HEADS)
        for i, (hname, short) in enumerate(self.HEADS):
            h.N[i] = self.nd if short == "logits" else self.ldc
            h.W[i], h.bias[i] = ptr(self.p(hname + ".fc.weight")), ptr(self.p(hname + ".fc.bias"))
            h.dW[i], h.dbias[i] = ptr(self.g(hname + ".fc.weight")), ptr(self.g(hname + ".fc.bias"))
            if outs is not None:
                h.out[i] = ptr(outs[i])
            if grads is not None:
                h.g[i] = ptr(grads[i])
        return h

    def heads_fwd(self, ctx, feat):
        NB, Cf = ctx.NB, self.topo["feat"]
        s = _abi.stream()
        outs = {}
        for hname, short in self.HEADS:
            outs[short] = ctx.t(short, (NB, self.nd if short == "logits" else self.ldc), torch.float32)
        la = ctx.t("la", (NB, self.nd), torch.float32)
        if not self.dry:
            h = self._heads_desc(outs=[outs[short] for _, short in self.HEADS])
            check(lib.sv_heads_fwd(ptr(feat), Cf, byref(h), NB, Cf, s))
            check(lib.sv_log_softmax_fwd(ptr(outs["logits"]), ptr(la), NB, self.nd, s))
        ctx.feat = feat
        return outs["mu"], outs["ls"], la

    def heads_bwd(self, ctx, g_mu, g_ls, g_la, side="net"):
        """returns g_feat fp32 [NB, feat]; accumulates head parameter gradients (side: as in encoder_bwd)"""
        NB, Cf = ctx.NB, self.topo["feat"]
        s = _abi.stream()
        g_logits = ctx.t("g.logits", (NB, self.nd), torch.float32)
        check(lib.sv_log_softmax_bwd(ptr(g_la), ptr(ctx.bufs["la"]), ptr(g_logits), NB, self.nd, s))
        g_feat = ctx.t("g.feat", (NB, Cf), torch.float32)
        h = self._heads_desc(grads=[g_mu, g_ls, g_logits])

        def weight_grads():
            check(lib.sv_heads_bwd_weight(byref(h), ptr(ctx.feat), Cf, NB, Cf, _abi.stream()))

        # the head weight gradients are not needed by the backward chain: side stream (joined at the end of encoder_bwd,
        # which always follows)
        side = (self.side if side == "net" else side) if not self.dry else None
        if side is not None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            side.wait_event(ev)
            with torch.cuda.stream(side):
                weight_grads()
        else:
            weight_grads()
        check(lib.sv_heads_bwd_input(byref(h), ptr(g_feat), Cf, NB, Cf, s))
        return g_feat

    def sample_fwd(self, ctx, group, mode, eps, unif=None, label=None, label_mix=None, lam_dev=None):
        """Sample.forward (vae.py:23-56) for one pass group; eps/unif are device tensors [B, .]"""
        B, D, nd = ctx.B, self.ldc, self.nd
        lat = ctx.t("latent", (ctx.NB, self.latent), torch.float32)
        r0 = group * B
        mu, ls, la = ctx.bufs["mu"], ctx.bufs["ls"], ctx.bufs["la"]
        check(lib.sv_sample_fwd(ptr(mu[r0:]), ptr(ls[r0:]), ptr(la[r0:]), ptr(eps), ptr(unif), ptr(label), ptr(label_mix),
                                ptr(lam_dev), mode, self.temperature, B, D, nd, ptr(lat[r0:]), self.latent, _abi.stream()))
        ctx.sample = getattr(ctx, "sample", {})
        ctx.sample[group] = dict(mode=mode, eps=eps)
        return lat

    def sample_bwd(self, ctx, group, g_latent, g_mu, g_ls, g_la, accumulate=1):
        B, D, nd = ctx.B, self.ldc, self.nd
        r0 = group * B
        sm = ctx.sample[group]
        lat = ctx.bufs["latent"]
        check(lib.sv_sample_bwd(ptr(g_latent[r0:]), self.latent, ptr(ctx.bufs["ls"][r0:]), ptr(sm["eps"]), ptr(lat[r0:]),
                                sm["mode"], self.temperature, B, D, nd, ptr(g_mu[r0:]), ptr(g_ls[r0:]),
                                ptr(g_la[r0:]) if g_la is not None else None, accumulate, _abi.stream()))

    # ---- decoder -------------------------------------------------------------------------------
    def decoder_fwd(self, ctx, latent):
        """latent fp32 [NB, ldc+nd] -> reconstruction logits fp32 NHWC [NB, 32, 32, in_ch]"""
        NB, G, B = ctx.NB, ctx.G, ctx.B
        s = _abi.stream()
        c0 = DEC_CHANNELS[0]
        y = ctx.t("d0.y", (NB, 1, 1, c0))
        st = ctx.z("d0.st", G * 2 * c0)
        check(lib.sv_linear_fwd(ptr(latent), self.latent, ptr(self.p("feature_reconstructor.decoder.0.weight")), c0, 1, None,
                                ptr(y) if self.f32 else None, None if self.f32 else ptr(y), c0, ptr(st), B, NB, c0, self.latent, s))
        ctx.dec = []
        ctx.dec_latent = latent
        cin, hin = c0, 1
        for li in range(5):
            a = ctx.t("d%d.a" % li, (NB, hin, hin, cin))
            bn = self._bn_fwd_act(ctx, "feature_reconstructor.decoder.%d" % (3 * li + 1), "d%d.bn" % li, st, cin, B * hin * hin, y, a, 0.0)
            ctx.dec.append(dict(y=y, a=a, bn=bn, hin=hin, cin=cin))
            last = li == 4
            cout = self.in_ch if last else DEC_CHANNELS[li + 1]
            ho = hin * 2
            if last:
                yn, stn = ctx.t("rec", (NB, ho, ho, cout), torch.float32), None
            else:
                yn, stn = ctx.t("d%d.y" % (li + 1), (NB, ho, ho, cout)), ctx.z("d%d.st" % (li + 1), G * 2 * cout)
            phases = []
            for py in range(2):
                for px in range(2):
                    self._igemm(ctx, "d%d.f%d%d" % (li + 1, py, px), a, "d%d.f%d%d" % (li + 1, py, px), NB, hin, hin, hin, hin,
                                out=None if last else yn, outf=yn if last else None, stats=stn, out_stride=2, off=(py, px),
                                OHf=ho, OWf=ho, n_valid=cout if last else 0, exact=True, batch=phases)
            self._igemm_flush(phases)     # the four output-parity phases: one grid where the kernel allows it
            y, st, cin, hin = yn, stn, cout, ho
        return y

    def decoder_bwd(self, ctx, g_rec):
        """g_rec: bf16 NHWC [NB, 32, 32, pad16(in_ch)] -> g_latent fp32 [NB, latent]"""
        NB, G, B = ctx.NB, ctx.G, ctx.B
        s = _abi.stream()
        g = g_rec
        cout, cout_p = self.in_ch, pad16(self.in_ch)
        # The weight gradients (and their batched reductions) only need (g, a) of their layer: with a second stream from the
        # engine (side_dec) they run beside the dgrad -> BatchNorm-backward chain instead of inside it.  Timeline before: this
        # function is the only chain in flight for ~430 us of the step, ~200 us of it weight gradients + reductions.
        sd = getattr(self, "side_dec", None) if (not self.dry and self.side is not None and self.timing is None) else None
        cur = torch.cuda.current_stream() if sd is not None else None

        def on_sd(fn):
            if sd is None:
                fn()
                return
            ev = torch.cuda.Event()
            ev.record(cur)
            sd.wait_event(ev)
            with torch.cuda.stream(sd):
                fn()

        for li in range(4, -1, -1):
            d = ctx.dec[li]
            hin, cin, ho = d["hin"], d["cin"], d["hin"] * 2
            wname = "feature_reconstructor.decoder.%d.weight" % (3 * (li + 1))
            taps = self.packs["d%d.d" % (li + 1)]["taps"]
            # ConvT weight gradient: rows = coarse input pixels, Gr = a_in (N = cin), A = g_out (C = cout)
            on_sd(lambda li=li, g=g, d=d, taps=taps, ho=ho, hin=hin, cin=cin, cout=cout, cout_p=cout_p, wname=wname:
                  self._wgrad(ctx, "d%d.w" % (li + 1), g, d["a"], taps, NB, ho, ho, cout_p, hin, hin, cin, 2, wname, cin, cout,
                              cout * 16, 16, 1, exact=True))
            g_a = ctx.t("g.d%d.a" % li, (NB, hin, hin, cin))
            self._igemm(ctx, "d%d.d" % (li + 1), g, "d%d.d" % (li + 1), NB, ho, ho, hin, hin, in_stride=2, out=g_a, exact=True)
            g_y = ctx.t("g.d%d.y" % li, (NB, hin, hin, cin))
            self._bn_bwd(ctx, "d%d.bn" % li, [dict(rec=d["bn"], g_a=g_a, slope=0.0)], d["y"], None, g_y, B * hin * hin, hin * hin)
            g, cout, cout_p = g_y, cin, cin
        c0 = DEC_CHANNELS[0]
        w0 = "feature_reconstructor.decoder.0.weight"
        g32, g16 = (ptr(g), None) if self.f32 else (None, ptr(g))

        def stem_weight_grad():
            self._wgrad_flush()
            check(lib.sv_linear_bwd_weight(g32, g16, c0, ptr(ctx.dec_latent), self.latent, ptr(self.g(w0)), c0, 1, None, NB, c0,
                                           self.latent, _abi.stream()))

        on_sd(stem_weight_grad)
        g_lat = ctx.t("g.latent", (NB, self.latent), torch.float32)
        check(lib.sv_linear_bwd_input(g32, g16, c0, ptr(self.p(w0)), c0, 1, ptr(g_lat), self.latent, 0, NB, c0, self.latent, s))
        if sd is not None:
            cur.wait_stream(sd)
        return g_lat

    # ---- BatchNorm running statistics ----------------------------------------------------------
    def bn_running_update(self, passes):
        """passes: list of (ctx, group) in the order the reference would have executed the forwards
        (P1, P2, P3, P4 for a SHOT step).  Only BNs that ran in a pass are updated for it.  One launch
        updates every BatchNorm (descriptor table built once per pass configuration)."""
        key = tuple((id(c), gi) for c, gi in passes)
        cache = getattr(self, "_run_tables", None)
        if cache is None:
            cache = self._run_tables = {}
        ent = cache.get(key)
        if ent is None:
            names = []
            for ctx, _ in passes:
                for n in ctx.bn:
                    if n not in names:
                        names.append(n)
            arr = (_abi.RunDesc * len(names))()
            max_c = 0
            for i, n in enumerate(names):
                ps = [(c, gi) for c, gi in passes if n in c.bn]
                assert len(ps) <= 4
                rec0 = ps[0][0].bn[n]
                d = arr[i]
                for k, (c, gi) in enumerate(ps):
                    d.mean[k] = c.bn[n]["mean"][gi].data_ptr()
                    d.var[k] = c.bn[n]["var"][gi].data_ptr()
                d.running_mean, d.running_var = ptr(self.b(n + ".running_mean")), ptr(self.b(n + ".running_var"))
                d.nbt = ptr(self.b(n + ".num_batches_tracked"))
                d.count, d.npass, d.C = float(rec0["count"]), len(ps), rec0["C"]
                max_c = max(max_c, rec0["C"])
            table = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(self.device)
            ent = cache[key] = (table, len(names), max_c)
        table, n_bn, max_c = ent
        check(lib.sv_bn_running_update_batched(ptr(table), n_bn, max_c, BN_MOMENTUM, _abi.stream()))

    def adopt_views(self, parent):
        """After its views ran their forward passes: rebuild the parent's tape / BatchNorm records over the
        FULL buffers (no launches), so that encoder_bwd / heads_bwd can run once over all pass groups."""
        self.dry = True
        try:
            cp = pad16(self.in_ch)
            feat = self.encoder_fwd(parent, parent.t("x_img", (parent.NB, 32, 32, cp)))
            self.heads_fwd(parent, feat)
        finally:
            self.dry = False

    def zero_grads(self):
        self.grads.zero_()
