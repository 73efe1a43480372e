"""Layer summary table (shapes, parameter counts) — the ``torchsummary.summary`` role
(``pipedream-fork/profiler/torchmodules/torchsummary/torchsummary.py:30-105``) and the
``benchmark/network_summary.py`` job, computed from a meta-device dry run of the layer graph
(no forward hooks, no real compute)."""
from __future__ import annotations

from typing import Tuple

import torch

from ..models import build_model
from ..models.graph import LayerGraphModel
from .layer_profiler import _node_output_shapes


def summary(model: LayerGraphModel, batch: int = -1) -> Tuple[str, int, int]:
    shapes = _node_output_shapes(model, max(batch, 1))
    lines = ["-" * 64, "{:>24}  {:>22} {:>14}".format("Layer (type)", "Output Shape", "Param #"),
             "=" * 64]
    total = trainable = 0
    for i, n in enumerate(model.nodes):
        mod = model.node_module(i)
        params = sum(p.numel() for p in mod.parameters())
        train = sum(p.numel() for p in mod.parameters() if p.requires_grad)
        total += params
        trainable += train
        shp = list(shapes[n.name])
        shp[0] = batch
        lines.append("{:>24}  {:>22} {:>14,}".format(
            "%s-%d" % (type(mod).__name__, i + 1), str(shp), params))
    act = sum(torch.Size(s).numel() for s in shapes.values()) / max(batch, 1) * 4 * 2 / 1e6
    lines += ["=" * 64, "Total params: {:,}".format(total), "Trainable params: {:,}".format(trainable),
              "Non-trainable params: {:,}".format(total - trainable), "-" * 64,
              "Forward/backward pass size (MB): %.2f" % act,
              "Params size (MB): %.2f" % (total * 4 / 1e6), "-" * 64]
    return "\n".join(lines), total, trainable


def network_summary(archs=None, datasets=("mnist", "cifar10", "imagenet")) -> None:
    """Print the summary of every model x dataset (reference ``benchmark/network_summary.py``)."""
    for ds in datasets:
        for arch in (archs or ["resnet18", "resnet50", "resnet152", "vgg11", "vgg16", "mobilenetv2"]):
            m = build_model(arch, ds)
            text, total, _ = summary(m)
            print("%s - %s (input %s)" % (ds, arch, "x".join(map(str, m.input_shape))))
            print(text)
            print()


if __name__ == "__main__":
    network_summary()
