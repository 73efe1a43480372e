"""Profiler: model -> ``profiles/<arch>/graph.txt`` for the partitioner.

Reference flow (``pipedream-fork/profiler/image_classification/main.py:381-409``): one batch
through ``torchsummary`` for shapes/params, 100 instrumented steps for per-layer fwd/bwd times,
an ``Input`` pseudo-layer carrying the data-loading time, then a tracing pass with
``TensorWrapper`` operator overloading to recover the graph edges (which is why the reference
needs a third, tracing-friendly copy of every model).

Here the model IS a layer graph, so edges are known; timings come from the CUDA-event layer
profiler (``utils.layer_profiler``).  Node ids are ``node<k>`` with ``k = layer index + 1``
(``node0`` is ``Input0``), descriptions are ``repr(module)`` (``Add()`` / ``View(-1)`` for the
join and flatten nodes), sizes are fp32 bytes, times are milliseconds — the reference's units.
"""
from __future__ import annotations

import argparse
import os
import time
from typing import Optional

import torch

from ..models import build_model
from ..models.graph import INPUT, LayerGraphModel
from ..utils.layer_profiler import LayerProfile, profile_layers
from .graph import Graph, Node


def graph_from_profile(model: LayerGraphModel, prof: LayerProfile, batch: int,
                       input_ms: float = 0.0) -> Graph:
    g = Graph()
    in_numel = batch
    for s in model.input_shape:
        in_numel *= s
    inp = Node("node0", "Input0", forward_compute_time=input_ms, backward_compute_time=0.0,
               activation_size=float(in_numel * 4), parameter_size=0.0)
    g.add_node(inp)
    ids = {INPUT: inp}
    for i, n in enumerate(model.nodes):
        node = Node("node%d" % (i + 1), prof.descs[i], prof.forward_ms[i], prof.backward_ms[i],
                    prof.activation_bytes[i], prof.parameter_bytes[i])
        g.add_node(node)
        ids[n.name] = node
    for n in model.nodes:
        for src in n.inputs:
            g.add_edge(ids[src], ids[n.name])
    return g


def stage_ids_from_graph(model: LayerGraphModel, gr: Graph):
    """Per-layer stage id read back from a partitioned graph (``gpus=N.txt``)."""
    out = []
    for i in range(len(model.nodes)):
        sid = gr.nodes["node%d" % (i + 1)].stage_id
        out.append(0 if sid is None else sid)
    return out


def profile_model(arch: str, dataset: str, batch_size: int, device=None,
                  dtype: Optional[torch.dtype] = None, trials: int = 5,
                  out_dir: Optional[str] = None, model: Optional[LayerGraphModel] = None,
                  verbose: bool = True) -> Graph:
    model = model if model is not None else build_model(arch, dataset)
    dev = torch.device(device or ("cuda" if torch.cuda.is_available() else "cpu"))
    t0 = time.time()
    sample = torch.randn((batch_size,) + tuple(model.input_shape))
    input_ms = (time.time() - t0) * 1e3     # synthetic "data loading" time of one batch
    prof = profile_layers(model, sample, device=dev, dtype=dtype, trials=trials)
    g = graph_from_profile(model, prof, batch_size, input_ms)
    if out_dir is not None:
        os.makedirs(out_dir, exist_ok=True)
        with open(os.path.join(out_dir, "graph.txt"), "w") as f:
            f.write(str(g))
        g.to_dot(os.path.join(out_dir, "graph"))
    if verbose:
        tf, tb = sum(prof.forward_ms), sum(prof.backward_ms)
        print("profiled %s/%s batch %d on %s: forward %.3f ms, backward %.3f ms, %d layers"
              % (arch, dataset, batch_size, dev, tf, tb, len(model.nodes)))
    return g


def cli(argv=None) -> int:
    ap = argparse.ArgumentParser(description="PipeDream profiler (graph.txt producer)")
    ap.add_argument("-a", "--arch", default="resnet50")
    ap.add_argument("-d", "--dataset", default="imagenet",
                    choices=["mnist", "cifar10", "imagenet", "highres"])
    ap.add_argument("-b", "--batch-size", type=int, default=128)
    ap.add_argument("--profile_directory", default="profiles")
    ap.add_argument("--trials", type=int, default=5)
    ap.add_argument("--device", default=None)
    a = ap.parse_args(argv)
    profile_model(a.arch, a.dataset, a.batch_size, a.device, trials=a.trials,
                  out_dir=os.path.join(a.profile_directory, a.arch))
    return 0


if __name__ == "__main__":
    raise SystemExit(cli())
