"""MobileNetV2 as a layer graph.

* ``style="imagenet"``: torchvision ``mobilenet_v2`` layout (ReLU6, residual
  only when stride==1 and cin==cout, dropout 0.2 head) — the reference's plain
  ImageNet dialect is ``torchvision.models.mobilenet_v2``
  (``benchmark/imagenet/imagenet_pytorch.py:29``).
* ``style="cifar"``: pytorch-cifar layout — stride-1 stem, stage-2 stride 1,
  ReLU, every block expands with a 1x1, stride-1 blocks always add a shortcut
  (projected with 1x1+BN when the width changes), ``AvgPool2d(4)`` (reference
  ``benchmark/cifar10/pytorchcifargitmodels/mobilenetv2.py:12-80``; MNIST =
  1 input channel, ``benchmark/mnist/models/mnistmobilenetv2.py:61``).
"""
from __future__ import annotations

import torch.nn as nn

from .graph import Add, Flatten, GraphBuilder, INPUT, LayerGraphModel, init_kaiming


def _conv(cin, cout, k, s, p, groups=1):
    return nn.Conv2d(cin, cout, kernel_size=k, stride=s, padding=p, groups=groups, bias=False)


def _build_imagenet(in_channels, num_classes, input_hw):
    setting = [(1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2),
               (6, 96, 3, 1), (6, 160, 3, 2), (6, 320, 1, 1)]
    g = GraphBuilder()
    g.block("stem")
    x = g.add("features.0.conv", _conv(in_channels, 32, 3, 2, 1), INPUT)
    x = g.add("features.0.bn", nn.BatchNorm2d(32), x)
    x = g.add("features.0.relu", nn.ReLU6(inplace=False), x)
    cin, idx = 32, 1
    for t, c, n, s in setting:
        for i in range(n):
            stride = s if i == 0 else 1
            pfx = f"features.{idx}"
            g.block(pfx)
            hidden = cin * t
            h = x
            if t != 1:
                h = g.add(f"{pfx}.expand.conv", _conv(cin, hidden, 1, 1, 0), h)
                h = g.add(f"{pfx}.expand.bn", nn.BatchNorm2d(hidden), h)
                h = g.add(f"{pfx}.expand.relu", nn.ReLU6(inplace=False), h)
            h = g.add(f"{pfx}.dw.conv", _conv(hidden, hidden, 3, stride, 1, groups=hidden), h)
            h = g.add(f"{pfx}.dw.bn", nn.BatchNorm2d(hidden), h)
            h = g.add(f"{pfx}.dw.relu", nn.ReLU6(inplace=False), h)
            h = g.add(f"{pfx}.project.conv", _conv(hidden, c, 1, 1, 0), h)
            h = g.add(f"{pfx}.project.bn", nn.BatchNorm2d(c), h)
            if stride == 1 and cin == c:
                h = g.add(f"{pfx}.add", Add(), h, x)
            x, cin, idx = h, c, idx + 1
    g.block("head")
    x = g.add(f"features.{idx}.conv", _conv(cin, 1280, 1, 1, 0), x)
    x = g.add(f"features.{idx}.bn", nn.BatchNorm2d(1280), x)
    x = g.add(f"features.{idx}.relu", nn.ReLU6(inplace=False), x)
    x = g.add("avgpool", nn.AdaptiveAvgPool2d((1, 1)), x)
    x = g.add("flatten", Flatten(), x)
    x = g.add("classifier.drop", nn.Dropout(0.2), x)
    g.add("classifier.fc", nn.Linear(1280, num_classes), x)
    m = LayerGraphModel(g.nodes, name="mobilenetv2", input_shape=(in_channels, input_hw, input_hw),
                        num_classes=num_classes)
    init_kaiming(m)
    return m


def _build_cifar(in_channels, num_classes, input_hw):
    cfg = [(1, 16, 1, 1), (6, 24, 2, 1), (6, 32, 3, 2), (6, 64, 4, 2),
           (6, 96, 3, 1), (6, 160, 3, 2), (6, 320, 1, 1)]
    g = GraphBuilder()
    g.block("stem")
    x = g.add("conv1", _conv(in_channels, 32, 3, 1, 1), INPUT)
    x = g.add("bn1", nn.BatchNorm2d(32), x)
    x = g.add("relu1", nn.ReLU(inplace=False), x)
    cin, idx = 32, 0
    for t, c, n, s in cfg:
        for i in range(n):
            stride = s if i == 0 else 1
            pfx = f"layers.{idx}"
            g.block(pfx)
            planes = t * cin
            h = g.add(f"{pfx}.conv1", _conv(cin, planes, 1, 1, 0), x)
            h = g.add(f"{pfx}.bn1", nn.BatchNorm2d(planes), h)
            h = g.add(f"{pfx}.relu1", nn.ReLU(inplace=False), h)
            h = g.add(f"{pfx}.conv2", _conv(planes, planes, 3, stride, 1, groups=planes), h)
            h = g.add(f"{pfx}.bn2", nn.BatchNorm2d(planes), h)
            h = g.add(f"{pfx}.relu2", nn.ReLU(inplace=False), h)
            h = g.add(f"{pfx}.conv3", _conv(planes, c, 1, 1, 0), h)
            h = g.add(f"{pfx}.bn3", nn.BatchNorm2d(c), h)
            if stride == 1:
                sc = x
                if cin != c:
                    sc = g.add(f"{pfx}.shortcut.conv", _conv(cin, c, 1, 1, 0), x)
                    sc = g.add(f"{pfx}.shortcut.bn", nn.BatchNorm2d(c), sc)
                h = g.add(f"{pfx}.add", Add(), h, sc)
            x, cin, idx = h, c, idx + 1
    g.block("head")
    x = g.add("conv2", _conv(320, 1280, 1, 1, 0), x)
    x = g.add("bn2", nn.BatchNorm2d(1280), x)
    x = g.add("relu2", nn.ReLU(inplace=False), x)
    x = g.add("avgpool", nn.AvgPool2d(4), x)
    x = g.add("flatten", Flatten(), x)
    g.add("linear", nn.Linear(1280, num_classes), x)
    return LayerGraphModel(g.nodes, name="mobilenetv2",
                           input_shape=(in_channels, input_hw, input_hw), num_classes=num_classes)


def build_mobilenetv2(arch: str = "mobilenetv2", style: str = "imagenet", in_channels: int = 3,
                      num_classes: int = 1000, input_hw: int = 224) -> LayerGraphModel:
    if style == "imagenet":
        return _build_imagenet(in_channels, num_classes, input_hw)
    return _build_cifar(in_channels, num_classes, input_hw)
