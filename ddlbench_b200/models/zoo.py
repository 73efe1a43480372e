"""Model zoo front door.

``build_model(arch, dataset)`` returns the single layer-graph definition of a
network; the reference keeps one factory per (dataset, dialect):
``mnist_<arch>()`` (``benchmark/mnist/models``), ``cifar10_<arch>()``
(``benchmark/cifar10/pytorchcifargitmodels``), ``inet_<arch>()``
(``benchmark/imagenet/gpipemodels``) and ``torchvision.models.<arch>()``.  Those
factory names are re-exported here for drop-in use.

Unlike the reference harness scripts, which instantiate *every* architecture at
import time (``benchmark/imagenet/imagenet_pytorch.py:19-30``), nothing is built
until asked for.
"""
from __future__ import annotations

from typing import Callable, Dict, Tuple

from .graph import LayerGraphModel
from .mobilenetv2 import build_mobilenetv2
from .resnet import build_resnet
from .vgg import build_vgg

# dataset -> (channels, height/width, classes, model style)
DATASETS: Dict[str, Tuple[int, int, int, str]] = {
    "mnist": (1, 28, 10, "cifar"),
    "cifar10": (3, 32, 10, "cifar"),
    "imagenet": (3, 224, 1000, "imagenet"),
    "highres": (3, 512, 1000, "imagenet"),
}

RESNETS = ("resnet18", "resnet34", "resnet50", "resnet101", "resnet152")
VGGS = ("vgg11", "vgg13", "vgg16", "vgg19")
ARCHS = RESNETS + VGGS + ("mobilenetv2",)
# the subset `run.sh -m all` iterates over (reference run/run/run_template.sh:60-65)
BENCH_ARCHS = ("resnet18", "resnet50", "resnet152", "vgg11", "vgg16", "mobilenetv2")


def input_shape(dataset: str) -> Tuple[int, int, int]:
    c, hw, _, _ = DATASETS[dataset]
    return (c, hw, hw)


def num_classes(dataset: str) -> int:
    return DATASETS[dataset][2]


def build_model(arch: str, dataset: str = "imagenet") -> LayerGraphModel:
    arch = arch.lower().replace("mobilenet_v2", "mobilenetv2")
    if dataset not in DATASETS:
        raise ValueError(f"unknown dataset {dataset!r}; choose from {sorted(DATASETS)}")
    c, hw, ncls, style = DATASETS[dataset]
    if arch in RESNETS:
        m = build_resnet(arch, style=style, in_channels=c, num_classes=ncls, input_hw=hw)
    elif arch in VGGS:
        vstyle = "mnist" if dataset == "mnist" else style
        m = build_vgg(arch, style=vstyle, in_channels=c, num_classes=ncls, input_hw=hw)
    elif arch == "mobilenetv2":
        m = build_mobilenetv2(arch, style=style, in_channels=c, num_classes=ncls, input_hw=hw)
    else:
        from .extra import EXTRA_BUILDERS
        if arch not in EXTRA_BUILDERS:
            raise ValueError(f"unknown architecture {arch!r}; choose from "
                             f"{ARCHS + tuple(EXTRA_BUILDERS)}")
        if style != "imagenet":
            raise ValueError(f"{arch} is defined for ImageNet-shaped inputs only")
        m = EXTRA_BUILDERS[arch](arch, in_channels=c, num_classes=ncls, input_hw=hw)
    m.dataset = dataset
    return m


def _factory(prefix: str, dataset: str, arch: str) -> Callable[..., LayerGraphModel]:
    def f(pretrained: bool = False, **kwargs) -> LayerGraphModel:
        if pretrained:
            print("We dont have a pretrained model for %s" % dataset)
        return build_model(arch, dataset)
    f.__name__ = f"{prefix}_{arch}"
    f.__doc__ = f"{arch} for {dataset}-shaped inputs (reference factory name ``{prefix}_{arch}``)."
    return f


# reference-compatible factory names: mnist_resnet18, cifar10_vgg16, inet_mobilenetv2, highres_resnet50 ...
for _prefix, _ds in (("mnist", "mnist"), ("cifar10", "cifar10"), ("inet", "imagenet"),
                     ("highres", "highres")):
    for _arch in ARCHS:
        globals()[f"{_prefix}_{_arch}"] = _factory(_prefix, _ds, _arch)

__all__ = ["build_model", "input_shape", "num_classes", "DATASETS", "ARCHS", "BENCH_ARCHS"] + [
    f"{p}_{a}" for p in ("mnist", "cifar10", "inet", "highres") for a in ARCHS]
