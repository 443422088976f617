"""Stand-in for the `lpips` pip package (not in the image).  LPIPS itself is a fixed input of the hot path
(BASELINE.json north_star); this deterministic frozen conv stack only keeps `--lpips` runs executable.
Perceptual-loss VALUES are not comparable with the real LPIPS-VGG (parity unpinned)."""
import torch
import torch.nn as nn


class LPIPS(nn.Module):
    def __init__(self, net='vgg', **_):
        super().__init__()
        g = torch.Generator().manual_seed(0)
        chans = [3, 16, 32, 64]
        self.convs = nn.ModuleList([nn.Conv2d(a, b, 3, stride=2, padding=1) for a, b in zip(chans[:-1], chans[1:])])
        for c in self.convs:
            with torch.no_grad():
                c.weight.copy_(torch.randn(c.weight.shape, generator=g) * (2.0 / (c.in_channels * 9)) ** 0.5)
                c.bias.zero_()
        for p in self.parameters():
            p.requires_grad_(False)

    def forward(self, x, y):
        d = 0
        for c in self.convs:
            x, y = torch.relu(c(x)), torch.relu(c(y))
            nx = x / (x.norm(dim=1, keepdim=True) + 1e-10); ny = y / (y.norm(dim=1, keepdim=True) + 1e-10)
            d = d + ((nx - ny) ** 2).sum(1, keepdim=True).mean((2, 3), keepdim=True)
        return d
