"""Profiling, partitioning and configuration files for GNMT — the roles of
``pipedream-fork/profiler/translation/train.py`` (per-layer timing of the seq2seq model into a
``graph.txt``) and of the hand-checked-in ``runtime/translation/models/{gnmt,gnmt_large}/
gpus=N/{mp,dp,hybrid}_conf.json`` + ``stageK.py`` files (M9): here the stage modules are produced
on the fly by ``GNMT.split(cuts)``, so a configuration is just the cut list + the rank maps.

    python -m ddlbench_b200.partition.gnmt_tools --arch gnmt --gpus 4 8 --out configs/gnmt
"""
from __future__ import annotations

import argparse
import json
import os
import time
from typing import Dict, Sequence, Tuple

import torch

from ..models.gnmt import GNMT
from . import optimizer as part_opt


def profile_gnmt(model: GNMT, batch_size: int = 64, seq_len: int = 48, device="cpu",
                 iters: int = 3) -> Dict[str, Tuple[float, float]]:
    """Per-layer (forward ms, backward ms): each layer is timed in isolation on the activations
    the previous layers produced (the reference monkey-patches every module's forward and hooks
    backward; the layer list here is explicit so no patching is needed)."""
    dev = torch.device(device)
    model = model.to(dev).train()
    g = torch.Generator().manual_seed(0)
    src = torch.randint(4, model.vocab_size, (seq_len, batch_size), generator=g).to(dev)
    lens = torch.full((batch_size,), seq_len, dtype=torch.int64, device=dev)
    tgt = torch.randint(4, model.vocab_size, (seq_len, batch_size), generator=g).to(dev)
    env = {"src": src, "src_len": lens, "tgt": tgt}
    times: Dict[str, Tuple[float, float]] = {}

    def sync():
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)
    for name, m, ins, outs in model.spec:
        args = [env[i].detach().requires_grad_(True) if env[i].is_floating_point() else env[i]
                for i in ins]
        f_ms = b_ms = 0.0
        for it in range(iters + 1):
            sync()
            t0 = time.perf_counter()
            y = m(*args)
            sync()
            t1 = time.perf_counter()
            ys = y if isinstance(y, tuple) and not isinstance(m, torch.nn.Embedding) else (y,)
            loss = sum(v.float().sum() for v in ys if torch.is_tensor(v) and v.requires_grad)
            if torch.is_tensor(loss):
                loss.backward()
            sync()
            t2 = time.perf_counter()
            if it > 0:            # first pass is warm-up
                f_ms += (t1 - t0) * 1e3 / iters
                b_ms += (t2 - t1) * 1e3 / iters
            m.zero_grad(set_to_none=True)
        for o, v in zip(outs, ys):
            env[o] = v.detach()
        times[name] = (f_ms, b_ms)
    return times


def annotate_activation_sizes(model: GNMT, graph, batch_size: int, seq_len: int) -> None:
    H, V = model.hidden_size, model.vocab_size
    for node in graph.nodes.values():
        d = node.node_desc
        if d.startswith("Input"):
            node.activation_size = 8.0 * batch_size * seq_len
        elif "EmulatedBiLSTM" in d:
            node.activation_size = 4.0 * batch_size * seq_len * 2 * H
        elif "RecurrentAttention" in d:
            node.activation_size = 4.0 * batch_size * seq_len * 2 * H
        elif "Classifier" in d:
            node.activation_size = 4.0 * batch_size * seq_len * V
        else:
            node.activation_size = 4.0 * batch_size * seq_len * H


def partition(model: GNMT, num_gpus: int, times=None, batch_size: int = 64, seq_len: int = 48,
              bandwidth: float = 450e9, straight: bool = False):
    """-> (cuts over model.spec, replication factor per stage)."""
    g = model.profile_graph(times)
    annotate_activation_sizes(model, g, batch_size, seq_len)
    _, _, reps, gr = part_opt.main([num_gpus], None, [bandwidth], straight_pipeline=straight,
                                   graph=g, out=lambda *a, **k: None, print_configuration=False)
    n_in = len(GNMT.INPUTS)
    stage_of = {}
    for node in gr.nodes.values():
        k = int(node.node_id[4:]) - n_in
        if k >= 0:
            stage_of[k] = node.stage_id or 0
    cuts, prev = [], 0
    for k in range(len(model.spec)):
        s = stage_of.get(k, prev)
        if s < prev:
            s = prev                      # keep stages contiguous in execution order
        if s != prev:
            cuts.append(k)
        prev = s
    cuts.append(len(model.spec))
    reps = list(reps)[:len(cuts)] if reps else [1] * len(cuts)
    while len(reps) < len(cuts):
        reps.append(1)
    return cuts, reps


def write_confs(out_dir: str, arch: str, model: GNMT, cuts: Sequence[int], reps: Sequence[int]) -> str:
    """``<out>/gpus=N/{mp,dp,hybrid}_conf.json`` in the reference's schema
    (``module_to_stage_map`` / ``stage_to_rank_map``) plus ``cuts`` for ``GNMT.split``."""
    n = sum(reps)
    d = os.path.join(out_dir, "gpus=%d" % n)
    os.makedirs(d, exist_ok=True)
    module_to_stage, start = [], 0
    for s, stop in enumerate(cuts):
        module_to_stage += [s] * (stop - start)
        start = stop
    module_to_stage.append(len(cuts) - 1)                      # the criterion module
    rank, s2r = 0, {}
    for s, r in enumerate(reps):
        s2r[str(s)] = list(range(rank, rank + r))
        rank += r
    layers = [name for name, _, _, _ in model.spec]
    json.dump({"module_to_stage_map": module_to_stage, "stage_to_rank_map": s2r, "cuts": list(cuts),
               "layers": layers, "arch": arch}, open(os.path.join(d, "hybrid_conf.json"), "w"), indent=1)
    even = [round((k + 1) * len(layers) / n) for k in range(n)]
    mts, start = [], 0
    for s, stop in enumerate(even):
        mts += [s] * (stop - start)
        start = stop
    mts.append(n - 1)
    json.dump({"module_to_stage_map": mts, "stage_to_rank_map": {str(k): [k] for k in range(n)},
               "cuts": even, "layers": layers, "arch": arch},
              open(os.path.join(d, "mp_conf.json"), "w"), indent=1)
    json.dump({"module_to_stage_map": [0] * (len(layers) + 1),
               "stage_to_rank_map": {"0": list(range(n))}, "cuts": [len(layers)], "layers": layers,
               "arch": arch}, open(os.path.join(d, "dp_conf.json"), "w"), indent=1)
    write_stage_package(d, arch, model, list(cuts), even)
    return d


_PKG = '''"""GENERATED by ddlbench_b200.partition.gnmt_tools — the stage modules of %(arch)s on %(n)d GPUs, in the
calling convention of the reference's generated packages
(pipedream-fork/runtime/translation/models/%(arch)s/gpus=N/__init__.py): ``arch()``,
``model(criterion)`` -> [(stage module, input names, output names), ..., (criterion, [...], ["loss"])],
``full_model()``.  ``conf`` selects the split: "hybrid" (the partitioner's stages, replicated as in
hybrid_conf.json) or "mp" (one stage per GPU, mp_conf.json)."""
from ddlbench_b200.models.gnmt import GNMT

ARCH = "%(arch)s"
NUM_LAYERS = %(layers)d
CUTS = {"hybrid": %(hybrid)r, "mp": %(mp)r}


def arch():
    return ARCH


def full_model(vocab_size=32320, hidden=1024, dropout=0.2):
    return GNMT(vocab_size, hidden, NUM_LAYERS, dropout)


def stages(conf="mp", **kw):
    """The ``GNMTStage`` modules (Stage0 .. StageK-1 of the reference's generated files)."""
    cuts = CUTS[conf]
    return full_model(**kw).split(len(cuts), cuts)


def model(criterion, conf="mp", **kw):
    parts = stages(conf, **kw)
    out = [(p, list(p.input_names), list(p.output_names)) for p in parts]
    return out + [(criterion, [parts[-1].output_names[0], "target"], ["loss"])]
'''


def write_stage_package(d: str, arch: str, model: GNMT, hybrid_cuts, mp_cuts) -> None:
    """``<d>/__init__.py``: importable stage-module package next to the conf files (the reference
    ships generated ``stageK.py`` files; here the stages are slices of one model definition, so the
    package is a thin, parameterised constructor instead of copied layer code)."""
    n_layers = sum(1 for name, _, _, _ in model.spec if name.startswith("enc.rnn"))
    with open(os.path.join(d, "__init__.py"), "w") as f:
        f.write(_PKG % {"arch": arch, "n": int(os.path.basename(d).split("=")[1]), "layers": n_layers,
                        "hybrid": list(hybrid_cuts), "mp": list(mp_cuts)})


def cli(argv=None) -> int:
    ap = argparse.ArgumentParser(description="profile / partition GNMT and write conf files")
    ap.add_argument("--arch", default="gnmt", choices=["gnmt", "gnmt_large"])
    ap.add_argument("--gpus", type=int, nargs="+", default=[4])
    ap.add_argument("--out", default="configs/gnmt")
    ap.add_argument("--vocab", type=int, default=32320)
    ap.add_argument("--hidden", type=int, default=1024)
    ap.add_argument("--batch-size", type=int, default=64)
    ap.add_argument("--seq-len", type=int, default=48)
    ap.add_argument("--measure", action="store_true", help="time layers on this device")
    ap.add_argument("--bandwidth", type=float, default=450e9)
    a = ap.parse_args(argv)
    model = GNMT(a.vocab, a.hidden, 4 if a.arch == "gnmt" else 8)
    times = None
    if a.measure:
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        times = profile_gnmt(model, a.batch_size, a.seq_len, dev)
    else:
        # analytic cost: proportional to parameter count x tokens (LSTM/Linear GEMMs dominate)
        times = {name: (1e-6 * sum(p.numel() for p in m.parameters()) + 0.01,
                        2e-6 * sum(p.numel() for p in m.parameters()) + 0.02)
                 for name, m, _, _ in model.spec}
    os.makedirs(a.out, exist_ok=True)
    with open(os.path.join(a.out, "graph.txt"), "w") as f:
        g = model.profile_graph(times)
        annotate_activation_sizes(model, g, a.batch_size, a.seq_len)
        f.write(str(g))
    for n in a.gpus:
        cuts, reps = partition(model, n, times, a.batch_size, a.seq_len, a.bandwidth)
        d = write_confs(a.out, a.arch, model, cuts, reps)
        print("%s gpus=%d: cuts %s replication %s -> %s" % (a.arch, n, cuts, reps, d))
    return 0


if __name__ == "__main__":
    raise SystemExit(cli())
