"""VGG-11/13/16/19 (no BatchNorm) as layer graphs.

* ``style="imagenet"``: torchvision layout with ``AdaptiveAvgPool2d(7)`` and
  the 25088-4096-4096-1000 classifier with dropout (reference
  ``benchmark/imagenet/gpipemodels/vgg/vgg.py:17-51``).
* ``style="cifar"``: pytorch-cifar layout, conv3x3(+bias)+ReLU, classifier
  ``Linear(512, 10)`` (reference ``benchmark/cifar10/pytorchcifargitmodels/vgg.py:15-38``).
* ``style="mnist"``: cifar layout with 1 input channel and the last max-pool
  removed because 28/2^5 < 1 (reference ``benchmark/mnist/models/mnistvgg.py:5-19``).
"""
from __future__ import annotations

import torch.nn as nn

from .graph import Flatten, GraphBuilder, INPUT, LayerGraphModel, init_kaiming

_CFG = {
    "vgg11": [64, "M", 128, "M", 256, 256, "M", 512, 512, "M", 512, 512, "M"],
    "vgg13": [64, 64, "M", 128, 128, "M", 256, 256, "M", 512, 512, "M", 512, 512, "M"],
    "vgg16": [64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M",
              512, 512, 512, "M"],
    "vgg19": [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M",
              512, 512, 512, 512, "M"],
}


def build_vgg(arch: str, style: str = "imagenet", in_channels: int = 3,
              num_classes: int = 1000, input_hw: int = 224) -> LayerGraphModel:
    cfg = list(_CFG[arch])
    if style == "mnist":
        cfg = cfg[:-1]
    g = GraphBuilder()
    x = INPUT
    cin = in_channels
    ci = pi = 0
    for v in cfg:
        if v == "M":
            pi += 1
            g.block(f"features.pool{pi}")
            x = g.add(f"features.pool{pi}", nn.MaxPool2d(kernel_size=2, stride=2), x)
        else:
            ci += 1
            g.block(f"features.conv{ci}")
            x = g.add(f"features.conv{ci}", nn.Conv2d(cin, v, kernel_size=3, padding=1), x)
            x = g.add(f"features.relu{ci}", nn.ReLU(inplace=False), x)
            cin = v
    g.block("head")
    if style == "imagenet":
        x = g.add("avgpool", nn.AdaptiveAvgPool2d((7, 7)), x)
        x = g.add("flatten", Flatten(), x)
        x = g.add("classifier.fc1", nn.Linear(512 * 7 * 7, 4096), x)
        x = g.add("classifier.relu1", nn.ReLU(inplace=False), x)
        x = g.add("classifier.drop1", nn.Dropout(), x)
        x = g.add("classifier.fc2", nn.Linear(4096, 4096), x)
        x = g.add("classifier.relu2", nn.ReLU(inplace=False), x)
        x = g.add("classifier.drop2", nn.Dropout(), x)
        g.add("classifier.fc3", nn.Linear(4096, num_classes), x)
    else:
        x = g.add("avgpool", nn.AvgPool2d(kernel_size=1, stride=1), x)
        x = g.add("flatten", Flatten(), x)
        g.add("classifier", nn.Linear(512, num_classes), x)
    model = LayerGraphModel(g.nodes, name=arch, input_shape=(in_channels, input_hw, input_hw),
                            num_classes=num_classes)
    if style == "imagenet":
        init_kaiming(model)
    return model
