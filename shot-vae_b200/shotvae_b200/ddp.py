"""Data-parallel gradient exchange: one process per GPU, the flat FP32 gradient arena all-reduced with
NCCL (NVLink 5 / NVSwitch) in buckets on a side stream while the rest of the backward still runs.

This replaces the reference's per-sub-block nn.DataParallel wrappers (wideresnet.py:78-94,
vae.py:108-133, decoder.py:63-64): instead of scatter / replicate / gather through GPU 0 around every
block of every forward, each rank owns a replica and a batch shard, BatchNorm statistics stay per
replica (DataParallel's semantics), and the only exchange is one sum of the gradient arena per step
(the SGD kernel divides by world_size).  Buckets follow the order in which the last backward
finalises gradients: the decoder range (88 % of the bytes) first, then encoder + heads.
"""
import torch
import torch.distributed as dist


class GradReducer:
    def __init__(self, net, group=None):
        self.net, self.group = net, group
        self.world = dist.get_world_size(group)
        self.stream = torch.cuda.Stream()
        split = net.poff["feature_reconstructor.decoder.0.weight"][0]
        self.buckets = {"encoder": (0, split), "decoder": (split, net.n_params)}
        self.bytes_per_step = net.n_params * 4

    def bucket_ready(self, name):
        """called on the compute stream right after the last kernel that writes this gradient range"""
        s, e = self.buckets[name]
        cur = torch.cuda.current_stream()
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            dist.all_reduce(self.net.grads[s:e], op=dist.ReduceOp.SUM, group=self.group)

    def wait_all(self):
        torch.cuda.current_stream().wait_stream(self.stream)
