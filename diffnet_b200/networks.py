"""The network stays ordinary PyTorch (north_star): this is the reference's 2-D U-Net
(``DiffNet/networks/unets.py:13-81``: 5 stride-2 encoders 32-64-128-256-256, 4 decoders with
skip concatenation, nearest x2 + 4x4 conv + sigmoid head; 4,163,585 parameters for
``UNet(2, 1)``) restated as a table-driven module, and a 3-D sibling with the same shape for
the 64^3 parametric config (the reference uses ``wgan3d.GoodGenerator`` there, also ~4.16 M
parameters; any inputs -> u network works with the FEM loss).
Neither is on the measured FEM hot path; they exist so the train-step benchmark has the
reference's parameter count to all-reduce."""
from __future__ import annotations

import re

import torch
from torch import nn


def _from_reference_key(k: str) -> str:
    """``down3.model.0.weight`` / ``up2.model.1.weight`` / ``final.2.bias`` (DiffNet/networks/unets.py:49-66)
    -> ``enc.2.0.weight`` / ``dec.1.1.weight`` / ``head.2.bias``; other keys unchanged."""
    m = re.match(r"(.*?)(down|up)(\d+)\.model\.(.*)$", k)
    if m:
        return f"{m.group(1)}{'enc' if m.group(2) == 'down' else 'dec'}.{int(m.group(3)) - 1}.{m.group(4)}"
    return re.sub(r"(^|\.)final\.", r"\1head.", k)


def _to_reference_key(k: str) -> str:
    m = re.match(r"(.*?)(enc|dec)\.(\d+)\.(.*)$", k)
    if m:
        return f"{m.group(1)}{'down' if m.group(2) == 'enc' else 'up'}{int(m.group(3)) + 1}.model.{m.group(4)}"
    return re.sub(r"(^|\.)head\.", r"\1final.", k)

_ENC = ((32, False, 0.0), (64, True, 0.0), (128, True, 0.0), (256, True, 0.5), (256, True, 0.5))
_DEC = ((256, 256, 0.5), (512, 128, 0.5), (256, 64, 0.0), (128, 32, 0.0))   # (in, out, dropout)


def _down(nd, cin, cout, norm, drop):
    conv = nn.Conv2d if nd == 2 else nn.Conv3d
    inorm = nn.InstanceNorm2d if nd == 2 else nn.InstanceNorm3d
    layers = [conv(cin, cout, 4, 2, 1, bias=False)]
    if norm:
        layers.append(inorm(cout))
    layers.append(nn.LeakyReLU(0.2))
    if drop:
        layers.append(nn.Dropout(drop))
    return nn.Sequential(*layers)


def _up(nd, cin, cout, drop):
    convt = nn.ConvTranspose2d if nd == 2 else nn.ConvTranspose3d
    inorm = nn.InstanceNorm2d if nd == 2 else nn.InstanceNorm3d
    layers = [convt(cin, cout, 4, 2, 1, bias=False), inorm(cout), nn.ReLU(inplace=True)]
    if drop:
        layers.append(nn.Dropout(drop))
    return nn.Sequential(*layers)


class UNet(nn.Module):
    """inputs (B, in_channels, *spatial) -> u (B, out_channels, *spatial) in (0, 1);
    spatial sizes must be multiples of 32.  ``nd=3`` builds the 3-D sibling (channels / 2 so the
    parameter count stays near the 2-D one)."""

    def __init__(self, in_channels=3, out_channels=1, nd=2, dropout=True):
        super().__init__()
        self.nd = nd
        scale = 1 if nd == 2 else 2
        enc, c = [], in_channels
        for cout, norm, drop in _ENC:
            cout //= scale
            enc.append(_down(nd, c, cout, norm, drop if dropout else 0.0))
            c = cout
        self.enc = nn.ModuleList(enc)
        self.dec = nn.ModuleList([_up(nd, cin // scale, cout // scale, drop if dropout else 0.0)
                                  for cin, cout, drop in _DEC])
        conv = nn.Conv2d if nd == 2 else nn.Conv3d
        pad = nn.ZeroPad2d((1, 0, 1, 0)) if nd == 2 else nn.ConstantPad3d((1, 0, 1, 0, 1, 0), 0.0)
        self.head = nn.Sequential(nn.Upsample(scale_factor=2), pad, conv(64 // scale, out_channels, 4, padding=1),
                                  nn.Sigmoid())
        # weights saved by the reference's UNet (down1..down5 / up1..up4 / final) load as they are
        self._register_load_state_dict_pre_hook(self._accept_reference_keys)

    @staticmethod
    def _accept_reference_keys(state_dict, prefix, *args):
        for k in [k for k in state_dict if k.startswith(prefix)]:
            nk = prefix + _from_reference_key(k[len(prefix):])
            if nk != k:
                state_dict[nk] = state_dict.pop(k)

    def reference_state_dict(self):
        """state_dict under the reference's parameter names (for its ``UNet.load_state_dict``)."""
        return {_to_reference_key(k): v for k, v in self.state_dict().items()}

    def forward(self, x):
        skips = []
        for e in self.enc:
            x = e(x)
            skips.append(x)
        skips.pop()
        for d in self.dec:
            x = torch.cat((d(x), skips.pop()), 1)
        return self.head(x)
