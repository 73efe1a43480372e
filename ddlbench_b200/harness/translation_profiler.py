"""The instrumented GNMT trainer of the PipeDream profiler
(``pipedream-fork/profiler/translation/train.py:1-474`` + ``seq2seq/train/trainer.py:204-229``):
train GNMT on one device — mixed-precision optimizer with an fp32 master copy, warm-up LR schedule,
gradient clipping — and, like the reference's ``--profile``-less default, emit the per-layer profile
graph the partitioner consumes once the model is warm.

    python pipedream-fork/profiler/translation/train.py --dataset-dir DATA --epochs 1 --batch-size 64 \\
        --math fp32 --num-layers 4 --profile-directory profiles/ [--steps 50] [--save ckpt/]
"""
from __future__ import annotations

import argparse
import os

import torch

from ..partition import gnmt_tools
from . import translation


def parse(argv=None):
    p = argparse.ArgumentParser(description="GNMT training + layer profile (PipeDream profiler)")
    p.add_argument("--dataset-dir", default=None, help="WMT16 directory (synthetic corpus when absent)")
    p.add_argument("--results-dir", default="results")
    p.add_argument("--save", default=None, help="checkpoint directory")
    p.add_argument("--epochs", type=int, default=1)
    p.add_argument("--batch-size", type=int, default=64)
    p.add_argument("--max-length-train", type=int, default=50)
    p.add_argument("--num-layers", type=int, default=4, help="4 = GNMT-8, 8 = GNMT-16")
    p.add_argument("--hidden-size", type=int, default=1024)
    p.add_argument("--math", default="fp32", choices=["fp32", "fp16", "bf16"])
    p.add_argument("--optimization-config", default="{'optimizer': 'Adam', 'lr': 5e-4}")
    p.add_argument("--lr", type=float, default=None)
    p.add_argument("--grad-clip", type=float, default=5.0)
    p.add_argument("--warmup-steps", type=int, default=200)
    p.add_argument("--smoothing", type=float, default=0.1)
    p.add_argument("--print-freq", type=int, default=10)
    p.add_argument("--steps", type=int, default=None, help="cap the iterations of an epoch")
    p.add_argument("--beam-size", type=int, default=5)
    p.add_argument("--eval-sentences", type=int, default=64)
    p.add_argument("--seed", type=int, default=1)
    p.add_argument("--device", default=None)
    p.add_argument("--profile-directory", "--profile_directory", default="profiles", dest="profile_directory")
    p.add_argument("--profile-seq-len", type=int, default=48)
    p.add_argument("--no-train", action="store_true", help="profile only")
    # pass-through of the partition driver (the previous behaviour of this path)
    p.add_argument("--arch", default=None, choices=[None, "gnmt", "gnmt_large"])
    p.add_argument("--gpus", type=int, nargs="+", default=None)
    p.add_argument("--out", default=None)
    p.add_argument("--measure", action="store_true")
    return p.parse_known_args(argv)


def main(argv=None) -> int:
    a, rest = parse(argv)
    if a.gpus is not None or a.out is not None:            # profile + partition + write confs
        args = []
        for k, v in (("--arch", a.arch), ("--out", a.out)):
            if v:
                args += [k, v]
        if a.gpus:
            args += ["--gpus"] + [str(g) for g in a.gpus]
        if a.measure:
            args.append("--measure")
        return gnmt_tools.cli(args + rest)
    import ast
    oc = ast.literal_eval(a.optimization_config) if a.optimization_config else {}
    lr = a.lr if a.lr is not None else float(oc.get("lr", 5e-4))
    out = {}
    if not a.no_train:
        targs = ["--num-layers", str(a.num_layers), "--hidden", str(a.hidden_size), "-b", str(a.batch_size),
                 "--max-len", str(a.max_length_train), "--epochs", str(a.epochs), "--lr", repr(lr),
                 "--optimizer", str(oc.get("optimizer", "Adam")), "--smoothing", str(a.smoothing),
                 "--grad-clip", str(a.grad_clip), "--warmup-steps", str(a.warmup_steps),
                 "--dtype", "fp32" if a.math == "fp32" else "bf16", "--beam-size", str(a.beam_size),
                 "--eval-sentences", str(a.eval_sentences), "--print-freq", str(a.print_freq),
                 "--seed", str(a.seed)]
        if a.dataset_dir:
            targs += ["--data-dir", a.dataset_dir]
        if a.steps:
            targs += ["--steps", str(a.steps)]
        if a.save:
            targs += ["--checkpoint-dir", a.save]
        if a.device:
            targs += ["--device", a.device]
        out = translation.main(targs)
    # layer profile of the (now warm) model -> graph.txt, the partitioner's input
    from ..models.gnmt import GNMT
    dev = "cuda" if (torch.cuda.is_available() and a.device != "cpu") else "cpu"
    vocab = 32320
    model = GNMT(vocab, a.hidden_size, a.num_layers)
    times = gnmt_tools.profile_gnmt(model, min(a.batch_size, 64), a.profile_seq_len, dev)
    arch = "gnmt" if a.num_layers <= 4 else "gnmt_large"
    d = os.path.join(a.profile_directory, arch)
    os.makedirs(d, exist_ok=True)
    g = model.profile_graph(times)
    gnmt_tools.annotate_activation_sizes(model, g, min(a.batch_size, 64), a.profile_seq_len)
    with open(os.path.join(d, "graph.txt"), "w") as f:
        f.write(str(g))
    fwd = sum(t[0] for t in times.values())
    bwd = sum(t[1] for t in times.values())
    print("Total forward %.3f ms, backward %.3f ms -> %s" % (fwd, bwd, os.path.join(d, "graph.txt")))
    if out:
        print("train_loss %.4f%s" % (out.get("train_loss", float("nan")),
                                     (", BLEU %.2f" % out["bleu"]) if "bleu" in out else ""))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
