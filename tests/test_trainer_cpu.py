"""Host logic of the epoch driver (shotvae_b200/train.py) against the reference's own scheduling code executed with
torch's MultiStepLR exactly as main_shot_vae.main() drives it (:198-199,222-258): warm-up, milestones, `ewm x 5`."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _trainer_stub(dataset="Cifar10", m2=False, annotated_ratio=0.1, adjust_lr=(400, 500, 550)):
    from shotvae_b200.train import Trainer
    t = Trainer.__new__(Trainer)          # schedules only: no model, no device
    t.base_lr, t.base_ewm, t.adjust_lr, t.dataset, t.m2, t.annotated_ratio = 0.1, 1e-3, tuple(adjust_lr), dataset, m2, annotated_ratio
    return t


def test_lr_and_ewm_schedules_follow_the_reference_main_loop():
    import warnings
    from torch.optim.lr_scheduler import MultiStepLR
    adjust_lr, epochs, lr0 = [4, 7, 9], 12, 0.1
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.SGD([p], lr=lr0, momentum=0.9, weight_decay=5e-4)
    sched = MultiStepLR(opt, milestones=adjust_lr)
    ewm, want_lr, want_ewm = 1e-3, [], []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for epoch in range(epochs):                       # main_shot_vae.py:222-258, verbatim control flow
            if epoch == 0:
                for g in opt.param_groups:
                    g["lr"] = lr0 * 0.2
            want_lr.append(opt.param_groups[0]["lr"])      # the rate train() runs this epoch at
            want_ewm.append(ewm)
            opt.step()
            sched.step(epoch)
            if epoch == 0:
                for g in opt.param_groups:
                    g["lr"] = lr0
            if epoch == adjust_lr[0]:
                ewm = ewm * 5
    t = _trainer_stub(adjust_lr=adjust_lr)
    for e in range(epochs):
        assert abs(t.lr_at(e) - want_lr[e]) < 1e-12, (e, t.lr_at(e), want_lr[e])
        assert abs(t.ewm_at(e) - want_ewm[e]) < 1e-15, (e, t.ewm_at(e), want_ewm[e])
    # the x5 rule is Cifar10-with-enough-labels only (:255-257); M2 has no such rule
    assert _trainer_stub(dataset="Cifar100", adjust_lr=adjust_lr).ewm_at(11) == 1e-3
    assert _trainer_stub(annotated_ratio=0.01, adjust_lr=adjust_lr).ewm_at(11) == 1e-3
    assert _trainer_stub(m2=True, adjust_lr=adjust_lr).ewm_at(11) == 1e-3


def test_launcher_builds_the_one_process_per_gpu_command():
    from shotvae_b200.launch import build_command
    cmd = build_command(8, "train.py", ["--epochs", "3"], port=29512)
    assert cmd[:3] == [sys.executable, "-m", "torch.distributed.run"]
    assert "--nproc-per-node" in cmd and cmd[cmd.index("--nproc-per-node") + 1] == "8"
    assert cmd[cmd.index("--master-addr") + 1] == "127.0.0.1" and cmd[cmd.index("--master-port") + 1] == "29512"
    assert cmd[-3:] == ["train.py", "--epochs", "3"]
    assert build_command(1, "train.py", [])[:2] == [sys.executable, "train.py"]       # one GPU: no launcher needed
