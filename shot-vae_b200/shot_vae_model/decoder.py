"""Transposed-conv decoder holder (reference decoder.py:4-69): latent -> 1024 -> 512 -> 256 -> 128 ->
64 -> C, BatchNorm + ReLU between, no output activation."""
from torch import nn


class Decoder(nn.Module):
    def __init__(self, latent_dim=100, num_feature=64, num_channel=1, data_parallel=True, kernel_size=(5, 6)):
        super().__init__()
        layers, cin = [], latent_dim
        for i, mult in enumerate((16, 8, 4, 2, 1)):
            cout = num_feature * mult
            layers.append(nn.ConvTranspose2d(cin, cout, kernel_size, 1, 0, bias=False) if i == 0
                          else nn.ConvTranspose2d(cin, cout, 4, 2, 1, bias=False))
            layers += [nn.BatchNorm2d(cout), nn.ReLU(True)]
            cin = cout
        layers.append(nn.ConvTranspose2d(cin, num_channel, 4, 2, 1, bias=False))
        self.decoder = nn.Sequential(*layers)
        self.kernel_size = tuple(kernel_size)
