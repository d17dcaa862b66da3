"""WideResNet encoder -- parameter holder with the reference's module tree (wideresnet.py:8-114).

The nn.Conv2d / nn.BatchNorm2d objects below are never called: they only own parameters (so that
state_dict keys, shapes and the default initialisation match the reference constructor exactly);
the forward/backward arithmetic is executed by shotvae_b200.plan.Net through libshotvae."""
import re

from torch import nn


def _bn_act_conv(seq, tag_norm, tag_act, tag_conv, cin, cout, k, stride, act):
    seq.add_module(tag_norm, nn.BatchNorm2d(cin))
    if act is not None:
        seq.add_module(tag_act, act)
    seq.add_module(tag_conv, nn.Conv2d(cin, cout, kernel_size=k, stride=stride, padding=k // 2, bias=False))


class _ResidualUnit(nn.Module):
    """pre-activation unit: f_block = BN-act-conv3x3(stride)-[dropout]-BN-act-conv3x3, optional i_block
    = BN-[act]-conv1x1(stride) when the shape changes."""

    def __init__(self, cin, cout, stride, drop_rate, make_act, shortcut_act):
        super().__init__()
        f = nn.Sequential()
        _bn_act_conv(f, "norm1", "relu1", "conv1", cin, cout, 3, stride, make_act())
        f.add_module("dropout", nn.Dropout(drop_rate))
        _bn_act_conv(f, "norm2", "relu2", "conv2", cout, cout, 3, 1, make_act())
        self.f_block = f
        if cin != cout or stride != 1:
            i = nn.Sequential()
            _bn_act_conv(i, "norm", "relu", "conv", cin, cout, 1, stride, make_act() if shortcut_act else None)
            self.i_block = i


class _Stem(nn.Sequential):
    def __init__(self, cin, cout, small_input):
        super().__init__()
        if not small_input:
            raise NotImplementedError("libshotvae implements the small_input (32x32, 3x3 stem) encoders only")
        self.add_module("conv0", nn.Conv2d(cin, cout, kernel_size=3, stride=1, padding=1, bias=True))


class _Stage(nn.Module):
    def __init__(self, attr, unit_fmt, cin, cout, depth, down, drop_rate, make_act, shortcut_act):
        super().__init__()
        seq = nn.Sequential()
        for i in range(depth):
            seq.add_module(unit_fmt % (i + 1), _ResidualUnit(cin if i == 0 else cout, cout, 2 if (down and i == 0) else 1,
                                                            drop_rate, make_act, shortcut_act))
        setattr(self, attr, seq)


class WideResNet(nn.Module):
    def __init__(self, num_input_channels=1, num_init_features=16, depth=28, width=2, data_parallel=True,
                 small_input=False, drop_rate=0.0):
        super().__init__()
        assert (depth - 4) % 6 == 0, 'depth should be 6n+4'
        if drop_rate:
            raise NotImplementedError("drop_rate != 0 is not implemented by libshotvae")
        n = (depth - 4) // 6
        widths = [int(v * width) for v in (16, 32, 64)]
        act = lambda: nn.LeakyReLU(inplace=True)
        enc = nn.Sequential()
        enc.add_module("pre_process", _Stem(num_input_channels, num_init_features, small_input))
        cin = num_init_features
        for i, w in enumerate(widths):
            enc.add_module("wideblock%d" % (i + 1), _Stage("wide_block", "wideunit%d", cin, w, n, i > 0, drop_rate, act, True))
            cin = w
        tr = nn.Sequential()
        tr.add_module("norm", nn.BatchNorm2d(cin))
        tr.add_module("relu", act())
        enc.add_module("transition", tr)
        self.encoder = enc
        self.num_feature_channel = cin
        self.plan_name = "wideresnet-%d-%d" % (depth, width)


def get_wide_resnet(name, drop_rate, input_channels=1, small_input=False, data_parallel=True):
    """name: wideresnet-<depth>-<width>, e.g. wideresnet-28-2"""
    depth, width = [int(v) for v in re.findall(r'\d+', name)]
    return WideResNet(depth=depth, width=width, drop_rate=drop_rate, num_input_channels=input_channels,
                      data_parallel=data_parallel, small_input=small_input)
