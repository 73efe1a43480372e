"""ResNet-18/34/50/101/152 as layer graphs, in the two stem/head styles the
reference benchmarks use.

* ``style="imagenet"``: torchvision layout — 7x7/2 stem, 3x3/2 max-pool,
  bottleneck stride on the 3x3, adaptive avg-pool, 1000-way FC
  (reference uses ``torchvision.models.resnet*`` in
  ``benchmark/imagenet/imagenet_pytorch.py:19-30`` and a sequential re-write in
  ``benchmark/imagenet/gpipemodels/resnet/resnet.py``).
* ``style="cifar"``: pytorch-cifar layout — 3x3/1 stem, no max-pool,
  ``AvgPool2d(4)``, ``Linear(512*exp, 10)``, projection shortcut whenever
  stride != 1 or the width changes (reference
  ``benchmark/cifar10/pytorchcifargitmodels/resnet.py:70-102``,
  ``benchmark/mnist/models/mnistresnet.py:68`` for the 1-channel variant).
"""
from __future__ import annotations

import torch.nn as nn

from .graph import Add, Flatten, GraphBuilder, INPUT, LayerGraphModel, init_kaiming

_DEPTHS = {
    "resnet18": ("basic", [2, 2, 2, 2]),
    "resnet34": ("basic", [3, 4, 6, 3]),
    "resnet50": ("bottleneck", [3, 4, 6, 3]),
    "resnet101": ("bottleneck", [3, 4, 23, 3]),
    "resnet152": ("bottleneck", [3, 8, 36, 3]),
}


def _conv(cin, cout, k, s, p):
    return nn.Conv2d(cin, cout, kernel_size=k, stride=s, padding=p, bias=False)


def _basic_block(g: GraphBuilder, pfx: str, x: str, cin: int, planes: int, stride: int,
                 project: bool) -> str:
    g.block(pfx)
    h = g.add(f"{pfx}.conv1", _conv(cin, planes, 3, stride, 1), x)
    h = g.add(f"{pfx}.bn1", nn.BatchNorm2d(planes), h)
    h = g.add(f"{pfx}.relu1", nn.ReLU(inplace=False), h)
    h = g.add(f"{pfx}.conv2", _conv(planes, planes, 3, 1, 1), h)
    h = g.add(f"{pfx}.bn2", nn.BatchNorm2d(planes), h)
    sc = x
    if project:
        sc = g.add(f"{pfx}.downsample.conv", _conv(cin, planes, 1, stride, 0), x)
        sc = g.add(f"{pfx}.downsample.bn", nn.BatchNorm2d(planes), sc)
    h = g.add(f"{pfx}.add", Add(), h, sc)
    return g.add(f"{pfx}.relu2", nn.ReLU(inplace=False), h)


def _bottleneck(g: GraphBuilder, pfx: str, x: str, cin: int, planes: int, stride: int,
                project: bool) -> str:
    cout = planes * 4
    g.block(pfx)
    h = g.add(f"{pfx}.conv1", _conv(cin, planes, 1, 1, 0), x)
    h = g.add(f"{pfx}.bn1", nn.BatchNorm2d(planes), h)
    h = g.add(f"{pfx}.relu1", nn.ReLU(inplace=False), h)
    h = g.add(f"{pfx}.conv2", _conv(planes, planes, 3, stride, 1), h)
    h = g.add(f"{pfx}.bn2", nn.BatchNorm2d(planes), h)
    h = g.add(f"{pfx}.relu2", nn.ReLU(inplace=False), h)
    h = g.add(f"{pfx}.conv3", _conv(planes, cout, 1, 1, 0), h)
    h = g.add(f"{pfx}.bn3", nn.BatchNorm2d(cout), h)
    sc = x
    if project:
        sc = g.add(f"{pfx}.downsample.conv", _conv(cin, cout, 1, stride, 0), x)
        sc = g.add(f"{pfx}.downsample.bn", nn.BatchNorm2d(cout), sc)
    h = g.add(f"{pfx}.add", Add(), h, sc)
    return g.add(f"{pfx}.relu3", nn.ReLU(inplace=False), h)


def build_resnet(arch: str, style: str = "imagenet", in_channels: int = 3,
                 num_classes: int = 1000, input_hw: int = 224) -> LayerGraphModel:
    kind, depths = _DEPTHS[arch]
    exp = 1 if kind == "basic" else 4
    blk = _basic_block if kind == "basic" else _bottleneck
    g = GraphBuilder()
    g.block("stem")
    if style == "imagenet":
        x = g.add("conv1", _conv(in_channels, 64, 7, 2, 3), INPUT)
        x = g.add("bn1", nn.BatchNorm2d(64), x)
        x = g.add("relu", nn.ReLU(inplace=False), x)
        x = g.add("maxpool", nn.MaxPool2d(kernel_size=3, stride=2, padding=1), x)
    else:
        x = g.add("conv1", _conv(in_channels, 64, 3, 1, 1), INPUT)
        x = g.add("bn1", nn.BatchNorm2d(64), x)
        x = g.add("relu", nn.ReLU(inplace=False), x)
    cin = 64
    for li, (planes, n) in enumerate(zip([64, 128, 256, 512], depths)):
        for bi in range(n):
            stride = (1 if li == 0 else 2) if bi == 0 else 1
            project = stride != 1 or cin != planes * exp
            x = blk(g, f"layer{li + 1}.{bi}", x, cin, planes, stride, project)
            cin = planes * exp
    g.block("head")
    if style == "imagenet":
        x = g.add("avgpool", nn.AdaptiveAvgPool2d((1, 1)), x)
    else:
        x = g.add("avgpool", nn.AvgPool2d(4), x)
    x = g.add("flatten", Flatten(), x)
    g.add("fc", nn.Linear(512 * exp, num_classes), x)
    model = LayerGraphModel(g.nodes, name=arch, input_shape=(in_channels, input_hw, input_hw),
                            num_classes=num_classes)
    if style == "imagenet":
        init_kaiming(model)
    return model
