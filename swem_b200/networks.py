"""Torch networks that surround the SWEM memory: key/value encoders, projections, decoder.

The north star keeps these as ordinary ``torch.nn`` modules ("ResNet encoders and decoder remain
torch modules"); only the EM + readout path between them is CUDA.  They are written from the
architecture description so that a checkpoint of the reference loads unchanged -- every
``state_dict`` key and shape matches:

* key encoder   : reference ``methods/basic_modules/networks.py:132-170`` (torchvision ResNet-50/18
                  trunk up to ``layer3``; ``layer1`` is registered as ``res2``; convs have no bias)
* value encoder : ``networks.py:56-129`` on top of ``mod_resnet.py:40-153`` (ResNet-18 trunk whose
                  convs DO carry a bias, 3+extra input planes) + CBAM fuser (``attentions.py:72-84``)
* key projection: ``networks.py:173-182``;  decoder: ``networks.py:186-216``

Nothing in this file is on the hot path; it has no dependency on the CUDA extension.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


# --------------------------------------------------------------------------------------------
# ResNet trunks (stem + 3 stages), parameterised so that one builder serves both encoders
# --------------------------------------------------------------------------------------------
class _Shortcut(nn.Sequential):
    """1x1 strided projection + BN, registered as ``downsample.0`` / ``downsample.1``."""

    def __init__(self, cin: int, cout: int, stride: int, bias: bool):
        super().__init__(nn.Conv2d(cin, cout, 1, stride=stride, bias=bias), nn.BatchNorm2d(cout))


class _Basic(nn.Module):
    """Two 3x3 convs (ResNet-18/34 unit)."""
    widen = 1

    def __init__(self, cin: int, width: int, stride: int, bias: bool):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, width, 3, stride=stride, padding=1, bias=bias)
        self.bn1 = nn.BatchNorm2d(width)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(width, width, 3, padding=1, bias=bias)
        self.bn2 = nn.BatchNorm2d(width)
        cout = width * self.widen
        self.downsample = _Shortcut(cin, cout, stride, bias) if (stride != 1 or cin != cout) else None

    def forward(self, x):
        y = self.relu(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y))
        skip = x if self.downsample is None else self.downsample(x)
        return self.relu(y + skip)


class _Bottle(nn.Module):
    """1x1 -> 3x3 (strided) -> 1x1 x4 (ResNet-50 unit, stride on the 3x3 like torchvision)."""
    widen = 4

    def __init__(self, cin: int, width: int, stride: int, bias: bool):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, width, 1, bias=bias)
        self.bn1 = nn.BatchNorm2d(width)
        self.conv2 = nn.Conv2d(width, width, 3, stride=stride, padding=1, bias=bias)
        self.bn2 = nn.BatchNorm2d(width)
        self.conv3 = nn.Conv2d(width, width * 4, 1, bias=bias)
        self.bn3 = nn.BatchNorm2d(width * 4)
        self.relu = nn.ReLU(inplace=True)
        cout = width * self.widen
        self.downsample = _Shortcut(cin, cout, stride, bias) if (stride != 1 or cin != cout) else None

    def forward(self, x):
        y = self.relu(self.bn1(self.conv1(x)))
        y = self.relu(self.bn2(self.conv2(y)))
        y = self.bn3(self.conv3(y))
        skip = x if self.downsample is None else self.downsample(x)
        return self.relu(y + skip)


def _stage(unit, cin: int, width: int, depth: int, stride: int, bias: bool) -> nn.Sequential:
    blocks = [unit(cin, width, stride, bias)]
    blocks += [unit(width * unit.widen, width, 1, bias) for _ in range(depth - 1)]
    return nn.Sequential(*blocks)


_TRUNKS = {
    # name: (unit, depths of the three stages used)
    'resnet18': (_Basic, (2, 2, 2)),
    'resnet50': (_Bottle, (3, 4, 6)),
}


def _he_init_(module: nn.Module) -> None:
    """He-normal convs / unit BN, the init both reference ResNets start from before weights load."""
    for m in module.modules():
        if isinstance(m, nn.Conv2d):
            fan = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
            nn.init.normal_(m.weight, 0.0, math.sqrt(2.0 / fan))
            if m.bias is not None:
                nn.init.zeros_(m.bias)
        elif isinstance(m, nn.BatchNorm2d):
            nn.init.ones_(m.weight)
            nn.init.zeros_(m.bias)


class _Normalise(nn.Module):
    """Holds the ImageNet ``mean`` / ``std`` buffers under the reference's names."""

    def __init__(self):
        super().__init__()
        self.register_buffer('mean', torch.tensor(IMAGENET_MEAN).view(1, 3, 1, 1))
        self.register_buffer('std', torch.tensor(IMAGENET_STD).view(1, 3, 1, 1))


class KeyEncoder(_Normalise):
    """Image -> (f16, f8, f4).  ``num_features`` = channels of (f16, f8, f4)."""

    def __init__(self, backbone_name: str = 'resnet50'):
        super().__init__()
        if backbone_name not in _TRUNKS:
            raise KeyError('The backbone {} is not supported yet.'.format(backbone_name))
        unit, depths = _TRUNKS[backbone_name]
        w = unit.widen
        self.num_features = [256 * w, 128 * w, 64 * w]
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, stride=2, padding=1)
        self.res2 = _stage(unit, 64, 64, depths[0], 1, False)            # 1/4
        self.layer2 = _stage(unit, 64 * w, 128, depths[1], 2, False)     # 1/8
        self.layer3 = _stage(unit, 128 * w, 256, depths[2], 2, False)    # 1/16
        _he_init_(self)

    def forward(self, f):
        x = (f - self.mean) / self.std
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        f4 = self.res2(x)
        f8 = self.layer2(f4)
        f16 = self.layer3(f8)
        return f16, f8, f4


# --------------------------------------------------------------------------------------------
# Plain residual block / CBAM / fuser used by the value encoder and decoder
# --------------------------------------------------------------------------------------------
class ResBlock(nn.Module):
    """Pre-activation residual pair of 3x3 convs, with a 3x3 projection when widths differ."""

    def __init__(self, indim: int, outdim: int | None = None):
        super().__init__()
        outdim = indim if outdim is None else outdim
        self.downsample = None if indim == outdim else nn.Conv2d(indim, outdim, 3, padding=1)
        self.conv1 = nn.Conv2d(indim, outdim, 3, padding=1)
        self.conv2 = nn.Conv2d(outdim, outdim, 3, padding=1)

    def forward(self, x):
        r = self.conv2(F.relu(self.conv1(F.relu(x))))
        if self.downsample is not None:
            x = self.downsample(x)
        return x + r


class _Flatten(nn.Module):
    def forward(self, x):
        return x.flatten(1)


class _ChannelGate(nn.Module):
    def __init__(self, channels: int, reduction: int = 16):
        super().__init__()
        self.mlp = nn.Sequential(_Flatten(), nn.Linear(channels, channels // reduction), nn.ReLU(),
                                 nn.Linear(channels // reduction, channels))

    def forward(self, x):
        att = self.mlp(x.mean(dim=(2, 3))) + self.mlp(x.amax(dim=(2, 3)))
        return x * torch.sigmoid(att)[:, :, None, None]


class _ConvOnly(nn.Module):
    def __init__(self, cin, cout, k):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, padding=(k - 1) // 2)

    def forward(self, x):
        return self.conv(x)


class _SpatialGate(nn.Module):
    def __init__(self):
        super().__init__()
        self.spatial = _ConvOnly(2, 1, 7)

    def forward(self, x):
        pooled = torch.stack([x.amax(dim=1), x.mean(dim=1)], dim=1)
        return x * torch.sigmoid(self.spatial(pooled))


class CBAM(nn.Module):
    """Channel gate (avg+max pooled MLP) followed by a 7x7 spatial gate."""

    def __init__(self, channels: int):
        super().__init__()
        self.ChannelGate = _ChannelGate(channels)
        self.SpatialGate = _SpatialGate()

    def forward(self, x):
        return self.SpatialGate(self.ChannelGate(x))


class FeatureFusionBlock(nn.Module):
    def __init__(self, indim: int, outdim: int):
        super().__init__()
        self.block1 = ResBlock(indim, outdim)
        self.attention = CBAM(outdim)
        self.block2 = ResBlock(outdim, outdim)

    def forward(self, x, f16):
        x = self.block1(torch.cat([x, f16], dim=1))
        return self.block2(x + self.attention(x))


class _ValueTrunk(_Normalise):
    """ResNet-18 trunk with biased convs and ``3 + extra`` input planes + the CBAM fuser."""

    def __init__(self, in_dim: int, extra: int):
        super().__init__()
        self.conv1 = nn.Conv2d(3 + extra, 64, 7, stride=2, padding=3)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, stride=2, padding=1)
        self.layer1 = _stage(_Basic, 64, 64, 2, 1, True)
        self.layer2 = _stage(_Basic, 64, 128, 2, 2, True)
        self.layer3 = _stage(_Basic, 128, 256, 2, 2, True)
        _he_init_(self)
        self.fuser = FeatureFusionBlock(in_dim + 256, 512)

    def _encode(self, image, key_f16, planes):
        x = torch.cat([(image - self.mean) / self.std] + planes, dim=1)
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        x = self.layer3(self.layer2(self.layer1(x)))
        return self.fuser(x, key_f16)


class ValueEncoderSO(_ValueTrunk):
    """Single-object value encoder: image + object mask."""

    def __init__(self, in_dim: int = 1024):
        super().__init__(in_dim, extra=1)

    def forward(self, image, key_f16, mask):
        return self._encode(image, key_f16, [mask])


class ValueEncoder(_ValueTrunk):
    """Multi-object value encoder: image + object mask + mask of the other objects."""

    def __init__(self, in_dim: int = 1024):
        super().__init__(in_dim, extra=2)

    def forward(self, image, key_f16, mask, other_masks):
        return self._encode(image, key_f16, [mask, other_masks])


class KeyProjection(nn.Module):
    def __init__(self, indim: int, keydim: int):
        super().__init__()
        self.key_proj = nn.Conv2d(indim, keydim, 3, padding=1)
        nn.init.orthogonal_(self.key_proj.weight.data)
        nn.init.zeros_(self.key_proj.bias.data)

    def forward(self, x):
        return self.key_proj(x)


class UpsampleBlock(nn.Module):
    def __init__(self, skip_c: int, up_c: int, out_c: int):
        super().__init__()
        self.skip_conv = nn.Conv2d(skip_c, up_c, 3, padding=1)
        self.out_conv = ResBlock(up_c, out_c)

    def forward(self, skip_f, up_f):
        x = self.skip_conv(skip_f)
        x = x + F.interpolate(up_f, size=x.shape[-2:], mode='bilinear', align_corners=False)
        return self.out_conv(x)


class Decoder(nn.Module):
    """Object context (1/16) + skip features (1/8, 1/4) -> one logit plane at ``osize``."""

    def __init__(self, inplanes, mdim: int = 256):
        super().__init__()
        self.compress = ResBlock(inplanes[0], 512)
        self.up_16_8 = UpsampleBlock(inplanes[1], 512, mdim)
        self.up_8_4 = UpsampleBlock(inplanes[2], 256, mdim)
        self.pred = nn.Conv2d(mdim, 1, 3, padding=1)

    def lowres_logits(self, f16, f8, f4):
        """The logit plane at 1/4 resolution, i.e. ``forward`` without its final up-sampling."""
        x = self.up_8_4(f4, self.up_16_8(f8, self.compress(f16)))
        return self.pred(F.relu(x))

    def forward(self, f16, f8, f4, osize):
        return F.interpolate(self.lowres_logits(f16, f8, f4), size=osize, mode='bilinear', align_corners=False)


class FeatureFusionLayer(nn.Module):
    """GLU fusion of [mem_out | qv | S] -> object context (reference ``modules.py:13-26``).

    Stays a cuDNN conv pair for now (SURVEY section 8f rank 1 lists it as the next widening step).
    """

    def __init__(self, indim: int, outdim: int):
        super().__init__()
        self.layer_f = nn.Conv2d(indim, outdim, 3, padding=1)
        self.layer_a = nn.Conv2d(indim, outdim, 3, padding=1)
        for conv in (self.layer_f, self.layer_a):
            nn.init.orthogonal_(conv.weight.data)
            nn.init.zeros_(conv.bias.data)

    def forward(self, x):
        return self.layer_f(x) * torch.sigmoid(self.layer_a(x))
