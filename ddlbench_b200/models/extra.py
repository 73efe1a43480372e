"""The wider image-classification zoo the upstream PipeDream profiler ships next to the
DDLBench models (``pipedream-fork/profiler/image_classification/models/{densenet,inception,
mobilenet,nasnet+layers,resnext,squeezenet}.py``; not selectable from the reference's
``run.sh``, only from the profiler's ``-a`` flag).  Each is one layer-graph definition, so all
of them get the plain / sequential / profile-graph dialects, GPipe and PipeDream partitioning
and the fused engine for free; channel joins are explicit ``Concat`` nodes.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.nn as nn

from .graph import Add, Flatten, GraphBuilder, INPUT, LayerGraphModel, init_kaiming


class Concat(nn.Module):
    """n-ary channel concatenation node."""

    def __init__(self, dim: int = 1):
        super().__init__()
        self.dim = dim

    def forward(self, *xs):
        return torch.cat(xs, self.dim)

    def __repr__(self):
        return "Concat(%d)" % self.dim


def _conv(cin, cout, k, s=1, p=0, groups=1, bias=False):
    return nn.Conv2d(cin, cout, kernel_size=k, stride=s, padding=p, groups=groups, bias=bias)


def _cbr(g: GraphBuilder, pfx: str, x: str, cin: int, cout: int, k, s=1, p=0, groups=1,
         eps: float = 1e-5) -> str:
    x = g.add(f"{pfx}.conv", _conv(cin, cout, k, s, p, groups), x)
    x = g.add(f"{pfx}.bn", nn.BatchNorm2d(cout, eps=eps), x)
    return g.add(f"{pfx}.relu", nn.ReLU(inplace=False), x)


def _finish(g: GraphBuilder, name: str, in_channels: int, input_hw: int, num_classes: int,
            init: bool = True) -> LayerGraphModel:
    m = LayerGraphModel(g.nodes, name=name, input_shape=(in_channels, input_hw, input_hw),
                        num_classes=num_classes)
    if init:
        init_kaiming(m)
    return m


# --------------------------------------------------------------------------------- DenseNet
_DENSENETS = {"densenet121": (32, (6, 12, 24, 16), 64), "densenet169": (32, (6, 12, 32, 32), 64),
              "densenet201": (32, (6, 12, 48, 32), 64), "densenet161": (48, (6, 12, 36, 24), 96)}


def build_densenet(arch="densenet121", in_channels=3, num_classes=1000, input_hw=224, bn_size=4):
    growth, blocks, feat = _DENSENETS[arch]
    g = GraphBuilder().block("stem")
    x = _cbr(g, "features.stem", INPUT, in_channels, feat, 7, 2, 3)
    x = g.add("features.pool0", nn.MaxPool2d(3, 2, 1), x)
    c = feat
    for bi, n in enumerate(blocks):
        for li in range(n):
            pfx = f"features.denseblock{bi + 1}.layer{li + 1}"
            g.block(pfx)
            h = g.add(f"{pfx}.norm1", nn.BatchNorm2d(c), x)
            h = g.add(f"{pfx}.relu1", nn.ReLU(inplace=False), h)
            h = g.add(f"{pfx}.conv1", _conv(c, bn_size * growth, 1), h)
            h = g.add(f"{pfx}.norm2", nn.BatchNorm2d(bn_size * growth), h)
            h = g.add(f"{pfx}.relu2", nn.ReLU(inplace=False), h)
            h = g.add(f"{pfx}.conv2", _conv(bn_size * growth, growth, 3, 1, 1), h)
            x = g.add(f"{pfx}.cat", Concat(), x, h)
            c += growth
        if bi != len(blocks) - 1:
            pfx = f"features.transition{bi + 1}"
            g.block(pfx)
            x = g.add(f"{pfx}.norm", nn.BatchNorm2d(c), x)
            x = g.add(f"{pfx}.relu", nn.ReLU(inplace=False), x)
            x = g.add(f"{pfx}.conv", _conv(c, c // 2, 1), x)
            x = g.add(f"{pfx}.pool", nn.AvgPool2d(2, 2), x)
            c //= 2
    g.block("head")
    x = g.add("features.norm5", nn.BatchNorm2d(c), x)
    x = g.add("features.relu5", nn.ReLU(inplace=False), x)
    x = g.add("avgpool", nn.AdaptiveAvgPool2d((1, 1)), x)
    x = g.add("flatten", Flatten(), x)
    g.add("classifier", nn.Linear(c, num_classes), x)
    return _finish(g, arch, in_channels, input_hw, num_classes)


# -------------------------------------------------------------------------------- SqueezeNet
def _fire(g, pfx, x, cin, sq, e1, e3):
    g.block(pfx)
    s = g.add(f"{pfx}.squeeze", _conv(cin, sq, 1, bias=True), x)
    s = g.add(f"{pfx}.squeeze_act", nn.ReLU(inplace=False), s)
    a = g.add(f"{pfx}.expand1x1", _conv(sq, e1, 1, bias=True), s)
    a = g.add(f"{pfx}.expand1x1_act", nn.ReLU(inplace=False), a)
    b = g.add(f"{pfx}.expand3x3", _conv(sq, e3, 3, 1, 1, bias=True), s)
    b = g.add(f"{pfx}.expand3x3_act", nn.ReLU(inplace=False), b)
    return g.add(f"{pfx}.cat", Concat(), a, b), e1 + e3


def build_squeezenet(arch="squeezenet1_0", in_channels=3, num_classes=1000, input_hw=224):
    g = GraphBuilder().block("stem")
    if arch == "squeezenet1_0":
        x = g.add("features.0", _conv(in_channels, 96, 7, 2, bias=True), INPUT)
        c = 96
        plan = ["P", (16, 64, 64), (16, 64, 64), (32, 128, 128), "P", (32, 128, 128),
                (48, 192, 192), (48, 192, 192), (64, 256, 256), "P", (64, 256, 256)]
    elif arch == "squeezenet1_1":
        x = g.add("features.0", _conv(in_channels, 64, 3, 2, bias=True), INPUT)
        c = 64
        plan = ["P", (16, 64, 64), (16, 64, 64), "P", (32, 128, 128), (32, 128, 128), "P",
                (48, 192, 192), (48, 192, 192), (64, 256, 256), (64, 256, 256)]
    else:
        raise ValueError(arch)
    x = g.add("features.1", nn.ReLU(inplace=False), x)
    for i, item in enumerate(plan):
        if item == "P":
            x = g.add(f"features.pool{i}", nn.MaxPool2d(3, 2, ceil_mode=True), x)
        else:
            x, c = _fire(g, f"features.fire{i}", x, c, *item)
    g.block("head")
    x = g.add("classifier.drop", nn.Dropout(0.5), x)
    x = g.add("classifier.conv", _conv(c, num_classes, 1, bias=True), x)
    x = g.add("classifier.relu", nn.ReLU(inplace=False), x)
    x = g.add("classifier.pool", nn.AdaptiveAvgPool2d((1, 1)), x)
    g.add("flatten", Flatten(), x)
    return _finish(g, arch, in_channels, input_hw, num_classes)


# ------------------------------------------------------------------------------ MobileNet v1
def build_mobilenet(arch="mobilenet", in_channels=3, num_classes=1000, input_hw=224):
    g = GraphBuilder().block("stem")
    x = _cbr(g, "model.0", INPUT, in_channels, 32, 3, 2, 1)
    c = 32
    cfg = [(64, 1), (128, 2), (128, 1), (256, 2), (256, 1), (512, 2)] + [(512, 1)] * 5 + \
          [(1024, 2), (1024, 1)]
    for i, (co, s) in enumerate(cfg):
        pfx = f"model.{i + 1}"
        g.block(pfx)
        x = _cbr(g, f"{pfx}.dw", x, c, c, 3, s, 1, groups=c)
        x = _cbr(g, f"{pfx}.pw", x, c, co, 1)
        c = co
    g.block("head")
    x = g.add("avgpool", nn.AdaptiveAvgPool2d((1, 1)), x)
    x = g.add("flatten", Flatten(), x)
    g.add("fc", nn.Linear(c, num_classes), x)
    return _finish(g, arch, in_channels, input_hw, num_classes)


# ----------------------------------------------------------------------------------- ResNeXt
_RESNEXTS = {"resnext50": ((3, 4, 6, 3), 32, 4), "resnext101": ((3, 4, 23, 3), 32, 4),
             "resnext152": ((3, 8, 36, 3), 32, 4)}


def build_resnext(arch="resnext50", in_channels=3, num_classes=1000, input_hw=224):
    depths, card, base = _RESNEXTS[arch]
    g = GraphBuilder().block("stem")
    x = _cbr(g, "stem", INPUT, in_channels, 64, 7, 2, 3)
    x = g.add("maxpool", nn.MaxPool2d(3, 2, 1), x)
    cin = 64
    for li, (planes, n) in enumerate(zip((64, 128, 256, 512), depths)):
        width = planes * base // 64 * card
        cout = planes * 4
        for bi in range(n):
            stride = 2 if (bi == 0 and li > 0) else 1
            pfx = f"layer{li + 1}.{bi}"
            g.block(pfx)
            h = _cbr(g, f"{pfx}.reduce", x, cin, width, 1)
            h = _cbr(g, f"{pfx}.group", h, width, width, 3, stride, 1, groups=card)
            h = g.add(f"{pfx}.expand.conv", _conv(width, cout, 1), h)
            h = g.add(f"{pfx}.expand.bn", nn.BatchNorm2d(cout), h)
            sc = x
            if stride != 1 or cin != cout:
                sc = g.add(f"{pfx}.downsample.conv", _conv(cin, cout, 1, stride), x)
                sc = g.add(f"{pfx}.downsample.bn", nn.BatchNorm2d(cout), sc)
            h = g.add(f"{pfx}.add", Add(), h, sc)
            x = g.add(f"{pfx}.relu", nn.ReLU(inplace=False), h)
            cin = cout
    g.block("head")
    x = g.add("avgpool", nn.AdaptiveAvgPool2d((1, 1)), x)
    x = g.add("flatten", Flatten(), x)
    g.add("fc", nn.Linear(cin, num_classes), x)
    return _finish(g, arch, in_channels, input_hw, num_classes)


# ------------------------------------------------------------------------------ Inception v3
def _branch(g, pfx, x, cin, specs) -> Tuple[str, int]:
    """Chain of conv-bn-relu: specs = [(cout, k, s, p), ...]."""
    c = cin
    for i, (co, k, s, p) in enumerate(specs):
        x = _cbr(g, f"{pfx}.{i}", x, c, co, k, s, p, eps=1e-3)
        c = co
    return x, c


def _inception(g, pfx, x, cin, branches, pool=None) -> Tuple[str, int]:
    """branches: list of spec chains; pool: None | ("avg", proj) | ("max",)."""
    g.block(pfx)
    outs, c = [], 0
    for bi, specs in enumerate(branches):
        if specs and isinstance(specs[0], list):            # split tail (InceptionE)
            stem, tails = specs[0], specs[1:]
            h, ch = _branch(g, f"{pfx}.b{bi}", x, cin, stem)
            for ti, t in enumerate(tails):
                o, co = _branch(g, f"{pfx}.b{bi}t{ti}", h, ch, [t])
                outs.append(o)
                c += co
        else:
            o, co = _branch(g, f"{pfx}.b{bi}", x, cin, specs)
            outs.append(o)
            c += co
    if pool is not None:
        if pool[0] == "avg":
            h = g.add(f"{pfx}.pool", nn.AvgPool2d(3, 1, 1), x)
            o, co = _branch(g, f"{pfx}.bpool", h, cin, [(pool[1], 1, 1, 0)])
        else:
            o, co = g.add(f"{pfx}.pool", nn.MaxPool2d(3, 2), x), cin
        outs.append(o)
        c += co
    return g.add(f"{pfx}.cat", Concat(), *outs), c


def build_inception_v3(arch="inception_v3", in_channels=3, num_classes=1000, input_hw=299):
    g = GraphBuilder().block("stem")
    x, c = _branch(g, "stem", INPUT, in_channels, [(32, 3, 2, 0), (32, 3, 1, 0), (64, 3, 1, 1)])
    x = g.add("stem.pool1", nn.MaxPool2d(3, 2), x)
    x, c = _branch(g, "stem2", x, c, [(80, 1, 1, 0), (192, 3, 1, 0)])
    x = g.add("stem.pool2", nn.MaxPool2d(3, 2), x)
    for i, pf in enumerate((32, 64, 64)):                                   # InceptionA x3
        x, c = _inception(g, f"Mixed_5{'bcd'[i]}", x, c,
                          [[(64, 1, 1, 0)], [(48, 1, 1, 0), (64, 5, 1, 2)],
                           [(64, 1, 1, 0), (96, 3, 1, 1), (96, 3, 1, 1)]], ("avg", pf))
    x, c = _inception(g, "Mixed_6a", x, c,                                    # InceptionB
                      [[(384, 3, 2, 0)], [(64, 1, 1, 0), (96, 3, 1, 1), (96, 3, 2, 0)]], ("max",))
    for i, c7 in enumerate((128, 160, 160, 192)):                            # InceptionC x4
        x, c = _inception(g, f"Mixed_6{'bcde'[i]}", x, c,
                          [[(192, 1, 1, 0)],
                           [(c7, 1, 1, 0), (c7, (1, 7), 1, (0, 3)), (192, (7, 1), 1, (3, 0))],
                           [(c7, 1, 1, 0), (c7, (7, 1), 1, (3, 0)), (c7, (1, 7), 1, (0, 3)),
                            (c7, (7, 1), 1, (3, 0)), (192, (1, 7), 1, (0, 3))]], ("avg", 192))
    x, c = _inception(g, "Mixed_7a", x, c,                                    # InceptionD
                      [[(192, 1, 1, 0), (320, 3, 2, 0)],
                       [(192, 1, 1, 0), (192, (1, 7), 1, (0, 3)), (192, (7, 1), 1, (3, 0)),
                        (192, 3, 2, 0)]], ("max",))
    for i in range(2):                                                       # InceptionE x2
        x, c = _inception(g, f"Mixed_7{'bc'[i]}", x, c,
                          [[(320, 1, 1, 0)],
                           [[(384, 1, 1, 0)], (384, (1, 3), 1, (0, 1)), (384, (3, 1), 1, (1, 0))],
                           [[(448, 1, 1, 0), (384, 3, 1, 1)], (384, (1, 3), 1, (0, 1)),
                            (384, (3, 1), 1, (1, 0))]], ("avg", 192))
    g.block("head")
    x = g.add("avgpool", nn.AdaptiveAvgPool2d((1, 1)), x)
    x = g.add("flatten", Flatten(), x)
    x = g.add("dropout", nn.Dropout(0.5), x)
    g.add("fc", nn.Linear(c, num_classes), x)
    return _finish(g, arch, in_channels, input_hw, num_classes)


# ------------------------------------------------------------------------------------ NASNet-A
class _NasOps:
    """Builder helpers for the NASNet-A cell vocabulary (Zoph et al. 2018, fig. 4): separable
    conv branches (relu - [dw kxk, pw 1x1, bn] x2), 3x3 avg/max pools, and the factorised
    reduction that halves the resolution of the skip input."""

    def __init__(self, g: GraphBuilder):
        self.g = g

    def sep(self, pfx, x, cin, cout, k, stride) -> str:
        g, p = self.g, k // 2
        x = g.add(f"{pfx}.relu", nn.ReLU(inplace=False), x)
        x = g.add(f"{pfx}.dw1", _conv(cin, cin, k, stride, p, groups=cin), x)
        x = g.add(f"{pfx}.pw1", _conv(cin, cout, 1), x)
        x = g.add(f"{pfx}.bn1", nn.BatchNorm2d(cout, eps=1e-3), x)
        x = g.add(f"{pfx}.relu2", nn.ReLU(inplace=False), x)
        x = g.add(f"{pfx}.dw2", _conv(cout, cout, k, 1, p, groups=cout), x)
        x = g.add(f"{pfx}.pw2", _conv(cout, cout, 1), x)
        return g.add(f"{pfx}.bn2", nn.BatchNorm2d(cout, eps=1e-3), x)

    def squeeze(self, pfx, x, cin, cout) -> str:
        g = self.g
        x = g.add(f"{pfx}.relu", nn.ReLU(inplace=False), x)
        x = g.add(f"{pfx}.conv", _conv(cin, cout, 1), x)
        return g.add(f"{pfx}.bn", nn.BatchNorm2d(cout, eps=1e-3), x)

    def factorized_reduce(self, pfx, x, cin, cout) -> str:
        g = self.g
        x = g.add(f"{pfx}.relu", nn.ReLU(inplace=False), x)
        a = g.add(f"{pfx}.path1.pool", nn.AvgPool2d(1, 2), x)
        a = g.add(f"{pfx}.path1.conv", _conv(cin, cout // 2, 1), a)
        b = g.add(f"{pfx}.path2.shift", _ShiftCrop(), x)
        b = g.add(f"{pfx}.path2.pool", nn.AvgPool2d(1, 2), b)
        b = g.add(f"{pfx}.path2.conv", _conv(cin, cout - cout // 2, 1), b)
        x = g.add(f"{pfx}.cat", Concat(), a, b)
        return g.add(f"{pfx}.bn", nn.BatchNorm2d(cout, eps=1e-3), x)

    def normal(self, pfx, x, x_prev, c_x, c_prev, f, reduce_prev: bool) -> Tuple[str, int]:
        """Normal cell; returns (output, channels = 6f)."""
        g = self.g
        g.block(pfx)
        h = self.squeeze(f"{pfx}.cur", x, c_x, f)
        p = (self.factorized_reduce(f"{pfx}.prev", x_prev, c_prev, f) if reduce_prev
             else self.squeeze(f"{pfx}.prev", x_prev, c_prev, f))
        c0 = g.add(f"{pfx}.c0", Add(), self.sep(f"{pfx}.c0l", h, f, f, 5, 1),
                   self.sep(f"{pfx}.c0r", p, f, f, 3, 1))
        c1 = g.add(f"{pfx}.c1", Add(), self.sep(f"{pfx}.c1l", p, f, f, 5, 1),
                   self.sep(f"{pfx}.c1r", p, f, f, 3, 1))
        c2 = g.add(f"{pfx}.c2", Add(), g.add(f"{pfx}.c2l", nn.AvgPool2d(3, 1, 1, count_include_pad=False), h), p)
        c3 = g.add(f"{pfx}.c3", Add(),
                   g.add(f"{pfx}.c3l", nn.AvgPool2d(3, 1, 1, count_include_pad=False), p),
                   g.add(f"{pfx}.c3r", nn.AvgPool2d(3, 1, 1, count_include_pad=False), p))
        c4 = g.add(f"{pfx}.c4", Add(), self.sep(f"{pfx}.c4l", h, f, f, 3, 1), h)
        return g.add(f"{pfx}.cat", Concat(), p, c0, c1, c2, c3, c4), 6 * f

    def reduction(self, pfx, x, x_prev, c_x, c_prev, f) -> Tuple[str, int]:
        """Reduction cell (stride 2); returns (output, channels = 4f)."""
        g = self.g
        g.block(pfx)
        h = self.squeeze(f"{pfx}.cur", x, c_x, f)
        p = self.squeeze(f"{pfx}.prev", x_prev, c_prev, f)
        c0 = g.add(f"{pfx}.c0", Add(), self.sep(f"{pfx}.c0l", h, f, f, 5, 2),
                   self.sep(f"{pfx}.c0r", p, f, f, 7, 2))
        c1 = g.add(f"{pfx}.c1", Add(), g.add(f"{pfx}.c1l", nn.MaxPool2d(3, 2, 1), h),
                   self.sep(f"{pfx}.c1r", p, f, f, 7, 2))
        c2 = g.add(f"{pfx}.c2", Add(),
                   g.add(f"{pfx}.c2l", nn.AvgPool2d(3, 2, 1, count_include_pad=False), h),
                   self.sep(f"{pfx}.c2r", p, f, f, 5, 2))
        c3 = g.add(f"{pfx}.c3", Add(),
                   g.add(f"{pfx}.c3l", nn.AvgPool2d(3, 1, 1, count_include_pad=False), c0), c1)
        c4 = g.add(f"{pfx}.c4", Add(), self.sep(f"{pfx}.c4l", c0, f, f, 3, 1),
                   g.add(f"{pfx}.c4r", nn.MaxPool2d(3, 2, 1), h))
        return g.add(f"{pfx}.cat", Concat(), c1, c2, c3, c4), 4 * f


class _ShiftCrop(nn.Module):
    """Shift by one pixel (drop first row/column, zero-pad at the end) — second path of the
    factorised reduction."""

    def forward(self, x):
        return nn.functional.pad(x[:, :, 1:, 1:], (0, 1, 0, 1))

    def __repr__(self):
        return "ShiftCrop()"


def build_nasnet(arch="nasnetamobile", in_channels=3, num_classes=1000, input_hw=224):
    stem_f, n_cells, filters = {"nasnetamobile": (32, 4, 44), "nasnetalarge": (96, 6, 168)}[arch]
    g = GraphBuilder().block("stem")
    ops = _NasOps(g)
    x0 = g.add("conv0.conv", _conv(in_channels, stem_f, 3, 2, 1), INPUT)
    x0 = g.add("conv0.bn", nn.BatchNorm2d(stem_f, eps=1e-3), x0)
    f0, f1 = filters // 4, filters // 2
    s0, c0 = ops.reduction("cell_stem_0", x0, x0, stem_f, stem_f, f0)
    s1, c1 = ops.reduction("cell_stem_1", s0, s0, c0, c0, f1)
    prev, cprev, cur, ccur = s0, c0, s1, c1
    reduce_prev = True                 # s0 is at twice the resolution of s1
    f = filters
    for stage in range(3):
        for i in range(n_cells):
            out, cout = ops.normal(f"cell_{stage}_{i}", cur, prev, ccur, cprev, f, reduce_prev)
            reduce_prev = False
            prev, cprev, cur, ccur = cur, ccur, out, cout
        if stage < 2:
            f *= 2
            out, cout = ops.reduction(f"reduction_cell_{stage}", cur, prev, ccur, cprev, f)
            prev, cprev, cur, ccur = cur, ccur, out, cout
            reduce_prev = True
    g.block("head")
    x = g.add("relu", nn.ReLU(inplace=False), cur)
    x = g.add("avgpool", nn.AdaptiveAvgPool2d((1, 1)), x)
    x = g.add("flatten", Flatten(), x)
    x = g.add("dropout", nn.Dropout(0.5), x)
    g.add("last_linear", nn.Linear(ccur, num_classes), x)
    return _finish(g, arch, in_channels, input_hw, num_classes)


# ------------------------------------------------------------------------------ AlexNet
def build_alexnet(arch="alexnet", in_channels=3, num_classes=1000, input_hw=224):
    """torchvision's AlexNet ("one weird trick" variant) — the model behind the reference's
    ``alexnet_{4,16}{dp,pipedream}.yml`` driver configurations and profiler
    (``pipedream-fork/profiler/image_classification/main.py -a alexnet``)."""
    g = GraphBuilder().block("features.0")
    x = g.add("features.0", nn.Conv2d(in_channels, 64, kernel_size=11, stride=4, padding=2), INPUT)
    x = g.add("features.1", nn.ReLU(inplace=False), x)
    x = g.add("features.2", nn.MaxPool2d(kernel_size=3, stride=2), x)
    g.block("features.3")
    x = g.add("features.3", nn.Conv2d(64, 192, kernel_size=5, padding=2), x)
    x = g.add("features.4", nn.ReLU(inplace=False), x)
    x = g.add("features.5", nn.MaxPool2d(kernel_size=3, stride=2), x)
    for i, (ci, co) in zip((6, 8, 10), ((192, 384), (384, 256), (256, 256))):
        g.block(f"features.{i}")
        x = g.add(f"features.{i}", nn.Conv2d(ci, co, kernel_size=3, padding=1), x)
        x = g.add(f"features.{i + 1}", nn.ReLU(inplace=False), x)
    x = g.add("features.12", nn.MaxPool2d(kernel_size=3, stride=2), x)
    g.block("head")
    x = g.add("avgpool", nn.AdaptiveAvgPool2d((6, 6)), x)
    x = g.add("flatten", Flatten(), x)
    x = g.add("classifier.0", nn.Dropout(), x)
    x = g.add("classifier.1", nn.Linear(256 * 6 * 6, 4096), x)
    x = g.add("classifier.2", nn.ReLU(inplace=False), x)
    x = g.add("classifier.3", nn.Dropout(), x)
    x = g.add("classifier.4", nn.Linear(4096, 4096), x)
    x = g.add("classifier.5", nn.ReLU(inplace=False), x)
    g.add("classifier.6", nn.Linear(4096, num_classes), x)
    return _finish(g, arch, in_channels, input_hw, num_classes)


EXTRA_BUILDERS = {}
EXTRA_BUILDERS["alexnet"] = build_alexnet
for _a in _DENSENETS:
    EXTRA_BUILDERS[_a] = build_densenet
for _a in ("squeezenet1_0", "squeezenet1_1"):
    EXTRA_BUILDERS[_a] = build_squeezenet
EXTRA_BUILDERS["mobilenet"] = build_mobilenet
for _a in _RESNEXTS:
    EXTRA_BUILDERS[_a] = build_resnext
EXTRA_BUILDERS["inception_v3"] = build_inception_v3
for _a in ("nasnetamobile", "nasnetalarge"):
    EXTRA_BUILDERS[_a] = build_nasnet
EXTRA_ARCHS = tuple(EXTRA_BUILDERS)
