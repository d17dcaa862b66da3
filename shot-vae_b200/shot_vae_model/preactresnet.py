"""PreActResNet encoder holder (reference preactresnet.py:19-133); basic-block variants only
(preactresnet18 is the configured one; the bottleneck nets are outside the hot path)."""
from torch import nn

from .wideresnet import _Stem, _Stage

preactresnet_dict = {
    "preactresnet18": {"expansion": 1, "block_config": [2, 2, 2, 2]},
    "preactresnet34": {"expansion": 1, "block_config": [3, 4, 6, 3]},
}


class PreActResNet(nn.Module):
    def __init__(self, expansion, block_config, num_input_channels=1, num_init_features=64, data_parallel=True,
                 small_input=False, drop_rate=0.0, plan_name="preactresnet18"):
        super().__init__()
        if expansion != 1:
            raise NotImplementedError("bottleneck PreActResNets are not implemented by libshotvae")
        if drop_rate:
            raise NotImplementedError("drop_rate != 0 is not implemented by libshotvae")
        act = lambda: nn.ReLU(inplace=True)
        enc = nn.Sequential()
        enc.add_module("pre_process", _Stem(num_input_channels, num_init_features, small_input))
        cin, cout = num_init_features, num_init_features
        for i, depth in enumerate(block_config):
            enc.add_module("block%d" % (i + 1), _Stage("preact_block", "unit%d", cin, cout, depth, i != 0, drop_rate, act, False))
            cin, cout = cout, cout * 2
        tr = nn.Sequential()
        tr.add_module("norm", nn.BatchNorm2d(cin))
        tr.add_module("relu", act())
        enc.add_module("transition", tr)
        self.encoder = enc
        self.num_feature_channel = cin
        self.plan_name = plan_name


def get_preact_resnet(name, drop_rate, input_channels=1, small_input=False, data_parallel=True):
    if name not in preactresnet_dict:
        raise NotImplementedError("{} not implemented".format(name))
    cfg = preactresnet_dict[name]
    return PreActResNet(num_input_channels=input_channels, expansion=cfg["expansion"], block_config=cfg["block_config"],
                        drop_rate=drop_rate, data_parallel=data_parallel, small_input=small_input, plan_name=name)
