"""ResNet-50 v1.5 trunk with the reference's module / state_dict surface (model/backbone/resnet.py:89-252,296-306).

The modules below are parameter containers (same attribute names => same state_dict keys as the reference's
``nn.Sequential(*list(resnet50().children())[:-2])``, MPNCOV.py:28-29); ``forward`` runs the fused CUDA pipeline.
"""
import torch.nn as nn

from .. import ops_resnet
from ..registry import BACKBONE


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, kernel_size=1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, kernel_size=3, stride=stride, padding=1, bias=False)   # v1.5: stride on 3x3
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, kernel_size=1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride


def _make_layer(inplanes, planes, blocks, stride):
    downsample = None
    if stride != 1 or inplanes != planes * 4:
        downsample = nn.Sequential(nn.Conv2d(inplanes, planes * 4, kernel_size=1, stride=stride, bias=False),
                                   nn.BatchNorm2d(planes * 4))
    layers = [Bottleneck(inplanes, planes, stride, downsample)]
    layers += [Bottleneck(planes * 4, planes) for _ in range(1, blocks)]
    return nn.Sequential(*layers)


class ResNetTrunk(nn.Sequential):
    """children()[:-2] of the reference ResNet: conv1, bn1, relu, maxpool, layer1..layer4 (indices 0..7)."""

    def __init__(self, layers=(3, 4, 6, 3)):
        mods = [nn.Conv2d(3, 64, kernel_size=7, stride=2, padding=3, bias=False), nn.BatchNorm2d(64), nn.ReLU(inplace=True),
                nn.MaxPool2d(kernel_size=3, stride=2, padding=1)]
        inplanes = 64
        for planes, n, stride in zip((64, 128, 256, 512), layers, (1, 2, 2, 2)):
            mods.append(_make_layer(inplanes, planes, n, stride))
            inplanes = planes * 4
        super().__init__(*mods)
        self.out_channels = inplanes
        for m in self.modules():                                   # resnet.py:191-196
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        self.__dict__['_plan'] = ops_resnet.TrunkPlan(self)

    def forward(self, x):
        return ops_resnet.resnet_trunk(x, self)


@BACKBONE.register
def resnet101(pretrained=False, progress=True, **kwargs):
    """Reference signature (resnet.py:309-319); the trunk of OSMENet (OSME.py:55-56).  Offline => reference initialisers
    unless $HAWKEYE_RESNET101_PTH points at torchvision's checkpoint."""
    import os
    import torch
    trunk = ResNetTrunk((3, 4, 23, 3))
    path = os.environ.get('HAWKEYE_RESNET101_PTH')
    if pretrained and path and os.path.exists(path):
        sd = torch.load(path, map_location='cpu')
        names = ['conv1', 'bn1', 'relu', 'maxpool', 'layer1', 'layer2', 'layer3', 'layer4']
        trunk.load_state_dict({str(names.index(k.split('.')[0])) + k[len(k.split('.')[0]):]: v for k, v in sd.items()
                               if k.split('.')[0] in names})
    elif pretrained and os.environ.get('HAWKEYE_ALLOW_RANDOM_INIT', '0') != '1':
        import logging
        logging.getLogger('hawkeye_b200').warning('resnet101(pretrained=True): no checkpoint at $HAWKEYE_RESNET101_PTH (%r) — the '
                                                  'trunk keeps the reference\'s RANDOM initialisation', path)
    return trunk


@BACKBONE.register
def resnet50(pretrained=False, progress=True, **kwargs):
    """Reference signature (resnet.py:296-306); offline => reference initialisers unless $HAWKEYE_RESNET50_PTH is set."""
    import os
    import torch
    trunk = ResNetTrunk((3, 4, 6, 3))
    path = os.environ.get('HAWKEYE_RESNET50_PTH')
    if pretrained and path and os.path.exists(path):
        sd = torch.load(path, map_location='cpu')
        names = ['conv1', 'bn1', 'relu', 'maxpool', 'layer1', 'layer2', 'layer3', 'layer4']
        remap = {}
        for k, v in sd.items():
            head = k.split('.')[0]
            if head in names:
                remap[str(names.index(head)) + k[len(head):]] = v
        trunk.load_state_dict(remap)           # strict: every trunk tensor must be present (fc.* is not part of the trunk)
    elif pretrained and os.environ.get('HAWKEYE_ALLOW_RANDOM_INIT', '0') != '1':
        import logging
        logging.getLogger('hawkeye_b200').warning(
            'resnet50(pretrained=True): no checkpoint at $HAWKEYE_RESNET50_PTH (%r) — the trunk keeps the reference\'s '
            'RANDOM initialisation.  Point HAWKEYE_RESNET50_PTH at torchvision\'s resnet50 .pth, or set '
            'HAWKEYE_ALLOW_RANDOM_INIT=1 (benchmarks / parity tests) to silence this.', path)
    return trunk
