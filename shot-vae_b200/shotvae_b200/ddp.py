"""Data-parallel gradient exchange: one process per GPU, the flat FP32 gradient arena all-reduced with
NCCL (NVLink 5 / NVSwitch) in buckets on a side stream while the rest of the backward still runs.

This replaces the reference's per-sub-block nn.DataParallel wrappers (wideresnet.py:78-94,
vae.py:108-133, decoder.py:63-64): instead of scatter / replicate / gather through GPU 0 around every
block of every forward, each rank owns a replica and a batch shard, BatchNorm statistics stay per
replica (DataParallel's semantics), and the only exchange is one sum of the gradient arena per step
(the SGD kernel divides by world_size).  Buckets follow the order in which the last backward
finalises gradients: the decoder range (88 % of the bytes) first, then the encoder block by block (last resolution block +
heads first), each all-reduce running beside the next block's backward.

Replicas start identical: construction broadcasts rank 0's parameters, BatchNorm buffers and momentum
(nn.DataParallel gets that for free from its single master copy; torch DDP does the same broadcast).
The per-rank noise streams must differ -- seed the CUDA generator per rank (bench.py does).
"""
import os

import torch
import torch.distributed as dist

DECODER_HEAD = "feature_reconstructor.decoder.0.weight"


class GradReducer:
    def __init__(self, net, group=None, broadcast=True):
        """net: plan.Net (or any object with the same flat arenas: params / grads / momentum / running / nbt,
        poff, n_params).  CPU arenas (gloo) are reduced synchronously -- that mode exists for the world_size-2
        CPU tests of this class; the product path is CUDA + NCCL."""
        self.net, self.group = net, group
        self.world = dist.get_world_size(group)
        self.cuda = net.grads.is_cuda
        self.stream = torch.cuda.Stream(device=net.grads.device) if self.cuda else None
        split = net.poff[DECODER_HEAD][0]
        # the two-bucket split relies on the decoder owning the TAIL of the arena (state_dict order of the reference
        # module tree: feature_extractor, heads, feature_reconstructor)
        tail = [k for k, (o, _, _) in net.poff.items() if o >= split]
        assert tail and all(k.startswith("feature_reconstructor.") for k in tail) and \
            all(o >= split for k, (o, _, _) in net.poff.items() if k.startswith("feature_reconstructor.")), \
            "decoder parameters are not the contiguous tail of the parameter arena"
        self.buckets = {"encoder": (0, split), "decoder": (split, net.n_params)}
        # the encoder range by backward segment (plan.Net.bwd_segments: last resolution block + transition + heads first, the
        # first block + conv0 last), so that each piece is reduced while the next segment of the backward still runs and only
        # the small first-block range (0.3 MB for WRN-28-2) is left exposed in front of the optimizer
        self.segments = []
        if hasattr(net, "segment_param_ranges"):
            for i, rng in enumerate(net.segment_param_ranges()):
                self.buckets["enc%d" % i] = rng
                self.segments.append("enc%d" % i)
            assert sum(e - s for s, e in (self.buckets[k] for k in self.segments)) == split, "encoder segments do not tile the encoder range"
        self.bytes_per_step = net.n_params * 4
        self.events = {}             # bucket name -> event recorded behind its all-reduce (wait_bucket)
        # SMs left to NCCL while a bucket is in flight.  The tcgen05 conv / weight-gradient kernels are persistent grids with
        # statically strided tiles, one CTA per SM: if a collective holds k SMs when such a grid starts, its last k CTAs form a
        # second wave and the kernel takes twice as long (MEASURED at 8 GPUs: 24 NVLS channels, 4.81 -> 5.04 ms/step).  With
        # NCCL_MAX_CTAS=k in the environment (bench.py / launch.py set it before the communicator exists) NCCL never uses more
        # than k CTAs, and the engine launches the backward's persistent grids with SMs - k CTAs (sv_set_cta_limit) so that both
        # always fit.  0 = no reservation.
        self.reserved_ctas = int(os.environ.get("NCCL_MAX_CTAS", "0") or 0) if self.cuda else 0
        self.cta_limit = 0
        if self.reserved_ctas > 0:
            sms = torch.cuda.get_device_properties(net.grads.device).multi_processor_count
            self.cta_limit = max(sms - self.reserved_ctas, sms // 2)
        if broadcast:
            self.broadcast_state()

    def broadcast_state(self, src=0):
        """rank `src`'s parameters, BatchNorm running statistics / counters and momentum to every rank"""
        for name in ("params", "running", "nbt", "momentum"):
            t = getattr(self.net, name, None)
            if t is not None and t.numel():
                dist.broadcast(t, src=src, group=self.group)
        if hasattr(self.net, "param_epoch"):
            self.net.param_epoch += 1          # masters changed: bf16 operand copies must be re-derived

    def state_checksum(self):
        """debug aid: (max - min) over ranks of a parameter checksum; 0.0 when the replicas are identical"""
        s = self.net.params.double().sum().reshape(1)
        lo, hi = s.clone(), s.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=self.group)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=self.group)
        return float(hi - lo)

    def bucket_ready(self, name):
        """called on the compute stream right after the last kernel that writes this gradient range"""
        s, e = self.buckets[name]
        if not self.cuda:
            dist.all_reduce(self.net.grads[s:e], op=dist.ReduceOp.SUM, group=self.group)
            return
        cur = torch.cuda.current_stream()
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            dist.all_reduce(self.net.grads[s:e], op=dist.ReduceOp.SUM, group=self.group)
            ev = self.events.get(name)
            if ev is None:
                ev = self.events[name] = torch.cuda.Event()
            ev.record(self.stream)

    def wait_bucket(self, name):
        """the compute stream waits for this bucket's all-reduce (and, the reducer stream being ordered, every earlier one)"""
        if self.cuda and name in self.events:
            torch.cuda.current_stream().wait_event(self.events[name])

    def wait_all(self):
        if self.cuda:
            torch.cuda.current_stream().wait_stream(self.stream)
