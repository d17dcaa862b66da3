"""Drop-in `lib.dataloader` (reference lib/dataloader.py) with the input pipeline moved onto the device.

The reference feeds train() from torchvision datasets through DataLoader worker processes: per sample
Pad(4, reflect) -> RandomHorizontalFlip -> RandomCrop(32) -> ToTensor on the host (:42-70), batches pinned and copied
with .cuda() every step.  At the B200 step rate (>20 k images/s per GPU) those workers are the cap, so here the uint8
dataset lives in HBM (CIFAR-10 train = 150 MB) and one libshotvae kernel (sv_augment_batch) gathers, pads, flips,
crops and converts a whole batch straight into the FP32 NCHW tensor the step consumes.

Kept from the reference API: `cifar10_dataset / cifar100_dataset / svhn_dataset / mnist_dataset(dataset_base_path,
train_flag)` and the SSL / SL per-class samplers `get_ssl_sampler / get_cifar10_ssl_sampler / get_cifar100_ssl_sampler /
get_cifar10_sl_sampler / get_cifar100_sl_sampler` (same arguments, same torch.randperm consumption per class, so the
same seed selects the same labelled subset; they return torch SubsetRandomSamplers as the reference's do).
New: `DeviceLoader(dataset, batch_size, sampler=...)` replaces `torch.utils.data.DataLoader(...)` in main()
(main_shot_vae.py:127-135); it yields (image, label) CUDA tensors, which the reference's train()/valid()/test()
consume unchanged (`.float().cuda()` on them is a no-op).
"""
import numpy as np
import torch
from torch.utils.data.sampler import SubsetRandomSampler

__all__ = ["DeviceImageDataset", "DeviceLoader", "mnist_dataset", "svhn_dataset", "cifar10_dataset", "cifar100_dataset",
           "get_ssl_sampler", "get_cifar10_ssl_sampler", "get_cifar100_ssl_sampler", "get_cifar10_sl_sampler",
           "get_cifar100_sl_sampler"]


class DeviceImageDataset:
    """uint8 images resident on the GPU + the transform description of the reference's torchvision pipeline.

    data: uint8 array / tensor [N, H, W, C] (hwc=True: CIFAR, MNIST as [N, 28, 28, 1]) or [N, C, H, W] (hwc=False: SVHN);
    targets: integer labels.  train_flag selects the augmenting transform (pad 4 reflect, flip, crop 32) or plain
    ToTensor, `pad_always` reproduces mnist_dataset, which pads and crops in both modes (:6-11)."""

    def __init__(self, data, targets, train_flag=True, hwc=True, flip=True, pad=4, out_size=32, pad_always=False, device="cuda"):
        data = torch.as_tensor(np.ascontiguousarray(data)) if not torch.is_tensor(data) else data.contiguous()
        assert data.dtype == torch.uint8 and data.dim() == 4
        self.hwc = bool(hwc)
        self.data = data.to(device)
        if hwc:
            self.n, self.src_h, self.src_w, self.ch = data.shape
        else:
            self.n, self.ch, self.src_h, self.src_w = data.shape
        self.targets = [int(t) for t in (targets.tolist() if hasattr(targets, "tolist") else targets)]
        self.labels = self.targets                       # torchvision's SVHN calls them .labels (main_shot_vae.py:168)
        self.targets_dev = torch.tensor(self.targets, dtype=torch.int64, device=self.data.device)
        self.train_flag = bool(train_flag)
        self.pad = pad if (train_flag or pad_always) else 0
        self.flip = bool(flip and train_flag)
        self.out_size = out_size
        self.range = self.src_h + 2 * self.pad - out_size   # crop offsets are drawn from [0, range]
        assert self.range >= 0 and self.src_w + 2 * self.pad - out_size == self.range

    def __len__(self):
        return self.n

    def draw_params(self, B, generator=None):
        """per-sample (crop row, crop column, flip) on the device generator; None when the transform is deterministic"""
        if self.range == 0 and not self.flip:
            return None
        dev = self.data.device
        p = torch.zeros(B, 3, dtype=torch.int32, device=dev)
        if self.range:
            p[:, :2] = torch.randint(0, self.range + 1, (B, 2), device=dev, generator=generator, dtype=torch.int32)
        if self.flip:
            p[:, 2] = (torch.rand(B, device=dev, generator=generator) < 0.5).to(torch.int32)
        return p

    def batch(self, index, params="draw", out=None, generator=None):
        """index: int64 CUDA tensor [B] -> (fp32 NCHW [B, C, out, out] in [0, 1], int64 labels [B])"""
        from shotvae_b200 import _abi
        from shotvae_b200._abi import lib, check, ptr
        index = index.to(self.data.device, torch.int64).contiguous()
        B = index.numel()
        if isinstance(params, str):
            params = self.draw_params(B, generator)
        if params is not None:
            params = params.to(self.data.device, torch.int32).contiguous()
        if out is None:
            out = torch.empty(B, self.ch, self.out_size, self.out_size, dtype=torch.float32, device=self.data.device)
        check(lib.sv_augment_batch(ptr(self.data), ptr(index), ptr(params), B, self.ch, self.src_h, self.src_w, self.pad,
                                   self.out_size, self.out_size, 1 if self.hwc else 0, ptr(out), _abi.stream()))
        return out, self.targets_dev[index]

    def __getitem__(self, i):
        img, lab = self.batch(torch.tensor([int(i)], dtype=torch.int64))
        return img[0], int(lab[0])


class DeviceLoader:
    """Batch iterator over a DeviceImageDataset: the stand-in for `DataLoader(dataset, batch_size=..., sampler=...)`.
    `sampler` is any iterable of dataset indices re-iterated every epoch (a SubsetRandomSampler draws a fresh host
    permutation each time, as in the reference); None = sequential.  The last batch may be short (drop_last=False is
    the reference's DataLoader default, main_shot_vae.py:127-135)."""

    def __init__(self, dataset, batch_size=1, sampler=None, drop_last=False, generator=None, **_ignored_dataloader_kwargs):
        self.dataset, self.batch_size, self.sampler, self.drop_last, self.generator = dataset, int(batch_size), sampler, drop_last, generator

    def __len__(self):
        n = len(self.sampler) if self.sampler is not None else len(self.dataset)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        order = torch.as_tensor(list(self.sampler) if self.sampler is not None else range(len(self.dataset)), dtype=torch.int64)
        order = order.to(self.dataset.data.device)            # one small H2D copy per epoch
        for s in range(0, order.numel(), self.batch_size):
            idx = order[s:s + self.batch_size]
            if self.drop_last and idx.numel() < self.batch_size:
                return
            yield self.dataset.batch(idx, generator=self.generator)


def _torchvision_arrays(kind, root, train_flag):
    """raw uint8 arrays + labels of a torchvision dataset (downloaded / cached exactly as the reference does)"""
    from torchvision import datasets
    if kind == "mnist":
        d = datasets.MNIST(root=root, train=train_flag, download=True)
        return d.data.unsqueeze(-1).numpy(), d.targets, True
    if kind == "svhn":
        d = datasets.SVHN(root=root, split="train" if train_flag else "test", download=True)
        return d.data, d.labels, False                   # SVHN is stored [N, 3, 32, 32]
    cls = datasets.CIFAR10 if kind == "cifar10" else datasets.CIFAR100
    d = cls(root=root, train=train_flag, download=True)
    return d.data, d.targets, True


def mnist_dataset(dataset_base_path, train_flag=True):
    data, targets, hwc = _torchvision_arrays("mnist", dataset_base_path, train_flag)
    return DeviceImageDataset(data, targets, train_flag, hwc, pad_always=True)       # (:6-16) pads + crops in both modes


def svhn_dataset(dataset_base_path, train_flag=True):
    data, targets, hwc = _torchvision_arrays("svhn", dataset_base_path, train_flag)
    return DeviceImageDataset(data, targets, train_flag, hwc)


def cifar100_dataset(dataset_base_path, train_flag=True):
    data, targets, hwc = _torchvision_arrays("cifar100", dataset_base_path, train_flag)
    return DeviceImageDataset(data, targets, train_flag, hwc)


def cifar10_dataset(dataset_base_path, train_flag=True):
    data, targets, hwc = _torchvision_arrays("cifar10", dataset_base_path, train_flag)
    return DeviceImageDataset(data, targets, train_flag, hwc)


def _per_class_split(labels, num_classes, cuts):
    """For every class in order: shuffle its sample positions with torch.randperm (the reference's draw, one per
    class, :87-88,127-128) and cut the shuffled list at `cuts` -> one index list per (start, stop) range."""
    labels = torch.as_tensor(labels)
    parts = [[] for _ in cuts]
    for c in range(num_classes):
        loc = torch.nonzero(labels == c).view(-1)
        loc = loc[torch.randperm(loc.size(0))]
        for dst, (a, b) in zip(parts, cuts):
            dst.extend(loc[a:b].tolist())
    return parts


def get_ssl_sampler(labels, valid_num_per_class, annotated_num_per_class, num_classes):
    """-> (sampler_valid, sampler_train_l, sampler_train_u); the unlabelled set includes the labelled one (:131-133)"""
    v, a = valid_num_per_class, annotated_num_per_class
    valid, train_l, train_u = _per_class_split(labels, num_classes, [(0, v), (v, v + a), (v, None)])
    return SubsetRandomSampler(valid), SubsetRandomSampler(train_l), SubsetRandomSampler(train_u)


def get_cifar10_ssl_sampler(labels, valid_num_per_class, annotated_num_per_class, num_classes):
    return get_ssl_sampler(labels, valid_num_per_class, annotated_num_per_class, num_classes)


def get_cifar100_ssl_sampler(labels, valid_num_per_class, annotated_num_per_class, num_classes=100):
    return get_ssl_sampler(labels, valid_num_per_class, annotated_num_per_class, num_classes)


def get_cifar10_sl_sampler(labels, valid_num_per_class, num_classes):
    """-> (sampler_valid, sampler_train) (:73-92)"""
    valid, train = _per_class_split(labels, num_classes, [(0, valid_num_per_class), (valid_num_per_class, None)])
    return SubsetRandomSampler(valid), SubsetRandomSampler(train)


def get_cifar100_sl_sampler(labels, valid_num_per_class, num_classes=100):
    return get_cifar10_sl_sampler(labels, valid_num_per_class, num_classes)
