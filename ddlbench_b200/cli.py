"""Command-line front end, flag-compatible with the reference's ``run/run/run.sh``
(``-b -f -g -n -p -m -q -s -h``; ``run/run/run.sh:16-47``) but launching LOCAL processes
instead of submitting an sbatch job (the reference is hard-wired to SURF's LISA cluster:
module loads, pip installs, mpirun/horovodrun/ssh fan-out — ``run/run/run_template.sh``).

    python -m ddlbench_b200.cli -b imagenet -f horovod -g 8 -m resnet50
    python -m ddlbench_b200.cli -b cifar10 -f gpipe -g 4 -m vgg16
    python -m ddlbench_b200.cli -b imagenet -f pipedream -g 8 -m resnet50 --stages 4x2

One worker process per GPU is started for horovod / gpipe / pipedream (``--worker`` is the
per-rank entry point, started through ``torch.distributed.run``).
"""
from __future__ import annotations

import argparse
import datetime
import os
import subprocess
import sys
from typing import List

from .config import BENCHMARKS, FRAMEWORKS, RunConfig
from .models.zoo import BENCH_ARCHS

USAGE = """usage: run.sh [-b benchmark] [-f framework] [-g gpus] [-n nodes] [-p loginterval] [-m model] [-q queue] [-s] [-h]
    -b  Benchmark: mnist, cifar10, imagenet, highres (synthetic only), all. Default: mnist
    -f  Framework: pytorch, horovod, gpipe, pipedream, all. Default: pytorch
    -g  Nr. GPUs per node. Default: 1
    -n  Nr. of nodes. Default: 1
    -p  Print statistics every X batches. Default: 25
    -m  Model name: resnet18, resnet50, resnet152, vgg11, vgg16, mobilenetv2, exp2 (resnet50, vgg16, mobilenetv2), all. Default: all
    -q  Queue name (accepted for compatibility; jobs run locally). Default: gpu_short
    -s  Disable synthetic data generation, use real data from $DATADIR (JPEG ImageFolder tree)
    -h  Display help."""


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(add_help=False, description="DDLBench-compatible launcher")
    p.add_argument("-b", dest="benchmark", default="mnist")
    p.add_argument("-f", dest="framework", default="pytorch")
    p.add_argument("-g", dest="gpus", type=int, default=1)
    p.add_argument("-n", dest="nodes", type=int, default=1)
    p.add_argument("-p", dest="loginterval", type=int, default=25)
    p.add_argument("-m", dest="model", default="all")
    p.add_argument("-q", dest="queue", default="gpu_short")
    p.add_argument("-s", dest="real_data", action="store_true")
    p.add_argument("-h", dest="help", action="store_true")
    # extensions (not in run.sh)
    p.add_argument("--epochs", type=int, default=None)
    p.add_argument("--steps", type=int, default=None, help="cap iterations per epoch")
    p.add_argument("--eval-steps", type=int, default=None)
    p.add_argument("--batch-size", type=int, default=None)
    p.add_argument("--microbatches", type=int, default=None)
    p.add_argument("--device", default="auto")
    p.add_argument("--allreduce", default="fused", choices=["fused", "nccl"])
    p.add_argument("--no-graph", action="store_true")
    p.add_argument("--stages", default=None, help="pipedream: '<stages>x<replicas>' e.g. 4x2")
    p.add_argument("--straight", action="store_true", help="pipedream: no replication")
    p.add_argument("--no-input-pipelining", "--no_input_pipelining", action="store_true",
                   dest="no_input_pipelining")
    p.add_argument("--macrobatch", action="store_true")
    # accepted for command-line parity with main_with_runtime.py; see the notes printed at start
    # hyper-parameters of the reference scripts (argparse flags of benchmark/*/*.py and
    # pipedream-fork/runtime/image_classification/main_with_runtime.py:29-104)
    p.add_argument("--lr", "--learning-rate", type=float, default=None, dest="lr")
    p.add_argument("--momentum", type=float, default=None)
    p.add_argument("--wd", "--weight-decay", type=float, default=None, dest="wd")
    p.add_argument("--lr_policy", "--lr-policy", default=None, dest="lr_policy",
                   choices=["step", "polynomial", "exponential_decay"])
    p.add_argument("--lr_warmup", "--lr-warmup", type=int, default=None, dest="lr_warmup",
                   help="pipedream: warm-up epochs")
    p.add_argument("--forward_only", "--forward-only", action="store_true", dest="forward_only",
                   help="pipedream: evaluation passes only")
    p.add_argument("--recompute", action="store_true",
                   help="pipedream: plumbed but a no-op in the reference (runtime.py:72-78) and here")
    p.add_argument("--fp16", action="store_true",
                   help="reference: apex fp16 master-weight training; here the 16-bit compute type is bf16 "
                        "(fp32 master weights + momentum), which --fp16 selects explicitly")
    p.add_argument("--loss_scale", "--loss-scale", type=float, default=1.0, dest="loss_scale",
                   help="pipedream: static loss scale S (loss gradient x S, every stage's optimizer / S)")
    p.add_argument("--config-path", "--config_path", dest="config_path", default=None, help="pipedream: explicit *_conf.json")
    p.add_argument("--checkpoint-dir", "--checkpoint_dir", dest="checkpoint_dir", default=None,
                   help="pipedream: write checkpoint.<stage>.pth.tar here after every epoch")
    p.add_argument("--resume", default=None,
                   help="pipedream: directory with checkpoint.<stage>.pth.tar files to continue from "
                        "(weights, fp32 master, momentum, BatchNorm buffers, epoch)")
    p.add_argument("--workdir", default=None)
    p.add_argument("--out", default="out")
    p.add_argument("--master-port", type=int, default=29400)
    p.add_argument("--worker", action="store_true", help=argparse.SUPPRESS)
    return p


def expand(args) -> List[RunConfig]:
    benches = list(BENCHMARKS if not args.real_data else BENCHMARKS[:3]) if args.benchmark == "all" \
        else [args.benchmark]
    if args.framework == "all":
        fws = list(FRAMEWORKS) if args.nodes == 1 else ["horovod", "pipedream"]
    else:
        fws = [args.framework]
    if args.model == "all":
        models = list(BENCH_ARCHS)
    elif args.model == "exp2":
        models = ["resnet50", "vgg16", "mobilenetv2"]
    else:
        models = [args.model]
    cfgs = []
    for fw in fws:
        for b in benches:
            for m in models:
                if fw == "pipedream" and m == "resnet152" and args.model in ("all", "exp2"):
                    continue   # the reference disables this pair; explicit -m resnet152 runs it here
                extra = {}
                for k in ("stages", "config_path", "checkpoint_dir", "workdir", "resume"):
                    v = getattr(args, k)
                    if v:
                        extra[k] = str(v)
                if args.straight:
                    extra["straight"] = "1"
                if args.no_input_pipelining:
                    extra["no_input_pipelining"] = "1"
                if args.macrobatch:
                    extra["macrobatch"] = "1"
                if args.loss_scale != 1.0:
                    if fw != "pipedream":
                        raise SystemExit("--loss_scale is a pipedream runtime option (main_with_runtime.py); "
                                         "the %s scripts of the reference have no loss scaling" % fw)
                    extra["loss_scale"] = repr(args.loss_scale)
                if args.lr_policy:
                    extra["lr_policy"] = args.lr_policy
                if args.lr_warmup is not None:
                    extra["lr_warmup"] = str(args.lr_warmup)
                if args.forward_only:
                    extra["forward_only"] = "1"
                kw = dict(benchmark=b, framework=fw, arch=m, gpus=args.gpus, nodes=args.nodes,
                          log_interval=args.loginterval, synthetic=not args.real_data,
                          device=args.device, allreduce=args.allreduce,
                          cuda_graph=not args.no_graph, steps_per_epoch=args.steps,
                          eval_steps=args.eval_steps, batch_size=args.batch_size,
                          microbatches=args.microbatches, lr=args.lr, momentum=args.momentum,
                          weight_decay=args.wd, extra=extra)
                if args.epochs is not None:
                    kw["epochs"] = args.epochs
                cfgs.append(RunConfig.from_env(**{k: v for k, v in kw.items() if v is not None}))
    return cfgs


def run_worker(cfg: RunConfig) -> dict:
    """Per-rank entry (inside torchrun for multi-GPU frameworks)."""
    import torch
    from .parallel import dp
    multi = "RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1
    if multi:
        backend = "gloo" if (cfg.device == "cpu" or not torch.cuda.is_available()) else "nccl"
        dp.init(backend=backend)
    if dp.rank() == 0:
        print(cfg.header(), flush=True)
    if cfg.framework in ("pytorch", "horovod"):
        from .harness.loop import run
        out = run(cfg, verbose=(cfg.framework == "horovod" or dp.rank() == 0))
    elif cfg.framework == "gpipe":
        from .harness.pipeline_loop import run_gpipe
        out = run_gpipe(cfg, verbose=True)
    else:
        from .harness.pipeline_loop import run_pipedream
        out = run_pipedream(cfg, verbose=True)
    if multi:
        dp.shutdown()
    return out


def _cfg_to_argv(cfg: RunConfig, args) -> List[str]:
    a = ["-b", cfg.benchmark, "-f", cfg.framework, "-g", str(cfg.gpus), "-n", str(cfg.nodes),
         "-p", str(cfg.log_interval), "-m", cfg.arch, "--device", cfg.device,
         "--allreduce", cfg.allreduce, "--epochs", str(cfg.epochs),
         "--batch-size", str(cfg.batch_size)]
    if cfg.framework == "gpipe":
        a += ["--microbatches", str(cfg.microbatches)]
    if cfg.steps_per_epoch:
        a += ["--steps", str(cfg.steps_per_epoch)]
    if cfg.eval_steps is not None:
        a += ["--eval-steps", str(cfg.eval_steps)]
    if not cfg.cuda_graph:
        a.append("--no-graph")
    if not cfg.synthetic:
        a.append("-s")
    for k, flag in (("stages", "--stages"), ("config_path", "--config-path"),
                    ("checkpoint_dir", "--checkpoint-dir"), ("workdir", "--workdir"),
                    ("resume", "--resume")):
        if k in cfg.extra:
            a += [flag, cfg.extra[k]]
    a += ["--lr", repr(cfg.lr), "--momentum", repr(cfg.momentum), "--wd", repr(cfg.weight_decay)]
    for k, flag in (("lr_policy", "--lr_policy"), ("lr_warmup", "--lr_warmup"), ("loss_scale", "--loss_scale")):
        if k in cfg.extra:
            a += [flag, cfg.extra[k]]
    for k, flag in (("straight", "--straight"), ("no_input_pipelining", "--no-input-pipelining"),
                    ("macrobatch", "--macrobatch"), ("forward_only", "--forward_only")):
        if cfg.extra.get(k) == "1":
            a.append(flag)
    return a


def main(argv=None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    args = build_parser().parse_args(argv)
    if args.help or not argv:
        print(USAGE)
        return 0
    if args.nodes > 1 and args.framework in ("pytorch", "gpipe"):
        print("\n\nOnly Horovod and Pipedream support multi-node execution. Continuing with 1 node.\n\n")
        args.nodes = 1
    if (args.fp16 or args.recompute) and not args.worker:
        print("note: --fp16 selects 16-bit compute with fp32 master weights, which here is bf16 (the "
              "tensor-core kernels are bf16; no fp16 kernels exist) — --loss_scale S is honoured exactly; "
              "--recompute is a no-op as in the reference runtime (runtime.py:72-78)")
    cfgs = expand(args)
    if args.worker:
        assert len(cfgs) == 1
        run_worker(cfgs[0])
        return 0
    stamp = datetime.datetime.now().strftime("%Y-%m-%d_%H-%M-%S")
    out_dir = os.path.join(args.out, stamp)
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, "info.txt"), "w") as f:
        f.write("Benchmark      %s\nFramework      %s\nGPUs / node    %d\nNodes          %d\n"
                "Log interval   %d\nModel name     %s\nQueue          %s\nUse synthetic  %s\n" % (
                    args.benchmark, args.framework, args.gpus, args.nodes, args.loginterval,
                    args.model, args.queue, str(not args.real_data).lower()))
    log_path = os.path.join(out_dir, "slurm.out")
    rc = 0
    with open(log_path, "a") as log:
        for i, cfg in enumerate(cfgs):
            world = 1 if cfg.framework == "pytorch" else cfg.gpus * cfg.nodes
            wargv = _cfg_to_argv(cfg, args) + ["--worker"]
            if world == 1:
                cmd = [sys.executable, "-m", "ddlbench_b200.cli"] + wargv
            else:
                cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                       "--nproc-per-node=%d" % world, "--master-addr", "127.0.0.1",
                       "--master-port", str(args.master_port + i), "-m", "ddlbench_b200.cli"] + wargv
            p = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            for line in p.stdout:
                sys.stdout.write(line)
                log.write(line)
            p.wait()
            rc = rc or p.returncode
    print("output written to %s" % log_path)
    return rc


if __name__ == "__main__":
    raise SystemExit(main())
