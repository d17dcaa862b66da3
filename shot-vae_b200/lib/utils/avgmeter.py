"""Drop-in `lib.utils.avgmeter` (reference lib/utils/avgmeter.py): running average used by the reference's train() /
valid() / test() for its timing and loss print-outs.  Host-side bookkeeping, no kernel involved."""


class AverageMeter(object):
    """last value, running sum, count and mean of the values fed to update(val, n)"""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count
