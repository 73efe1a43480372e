"""PipeDream runtime entry point at the reference's path with its flags (``--module
models.<arch>.gpus=<N> -b <batch> --config_path <conf.json> --rank R --local_rank L --master_addr A
--distributed_backend gloo|nccl --epochs E --lr ... --no_input_pipelining --macrobatch -s
--num_ranks_in_server K --highres H``).  One process per rank, exactly as ``pipedream_run.sh``
starts them; the stage is sliced from the layer graph according to the configuration file."""
import argparse
import json
import os
import re
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "..")))
from ddlbench_b200.cli import run_worker  # noqa: E402
from ddlbench_b200.config import RunConfig  # noqa: E402


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description="PipeDream runtime (ddlbench_b200)")
    ap.add_argument("--module", "-m", required=True)
    ap.add_argument("-b", "--batch-size", type=int, default=None, dest="batch_size")
    ap.add_argument("--config_path", default=None)
    ap.add_argument("--rank", type=int, default=None)
    ap.add_argument("--local_rank", type=int, default=0)
    ap.add_argument("--master_addr", default="127.0.0.1")
    ap.add_argument("--master_port", default="29500")
    ap.add_argument("--distributed_backend", default="nccl")
    ap.add_argument("--epochs", type=int, default=None)
    ap.add_argument("--lr", "--learning-rate", type=float, default=None, dest="lr")
    ap.add_argument("--momentum", type=float, default=None)
    ap.add_argument("--weight-decay", "--wd", type=float, default=None, dest="wd")
    ap.add_argument("--lr_policy", default=None)
    ap.add_argument("--lr_warmup", action="store_true")
    ap.add_argument("--print-freq", "-p", type=int, default=None, dest="print_freq")
    ap.add_argument("--no_input_pipelining", action="store_true")
    ap.add_argument("--macrobatch", action="store_true")
    ap.add_argument("--forward_only", action="store_true")
    ap.add_argument("--num_minibatches", type=int, default=None)
    ap.add_argument("--checkpoint_dir", default=None)
    ap.add_argument("--highres", type=int, default=0)
    ap.add_argument("--dataset", default=os.environ.get("BENCHMARK", "imagenet"),
                    help="extension: mnist | cifar10 | imagenet | highres (the reference encodes it in --module)")
    ap.add_argument("--data_dir", default=None,
                    help="image trees (<SET>/{train,val}/<class>/*) or the MNIST / CIFAR-10 dataset files; "
                         "without it batches are generated in memory")
    for flag in ("-j", "--workers", "-s", "--synthetic_data", "--num_ranks_in_server",
                 "--eval-batch-size", "-v", "--verbose_frequency", "--loss_scale", "--resume", "--start-epoch"):
        ap.add_argument(flag, default=None)
    ap.add_argument("--fp16", action="store_true")
    ap.add_argument("--recompute", action="store_true")
    a = ap.parse_args(argv)
    m = re.match(r"models\.([A-Za-z0-9_]+)\.gpus=(\d+)", a.module)
    arch = m.group(1) if m else a.module.split(".")[-1]
    world = int(m.group(2)) if m else 1
    if a.config_path:
        world = sum(len(v) for v in json.load(open(a.config_path))["stage_to_rank_map"].values())
    if a.rank is not None:                      # pipedream_run.sh convention: explicit ranks
        os.environ.update(RANK=str(a.rank), LOCAL_RANK=str(a.local_rank), WORLD_SIZE=str(world),
                          MASTER_ADDR=a.master_addr, MASTER_PORT=str(a.master_port))
    dataset = "highres" if a.highres else a.dataset
    for prefix in ("mnist", "cifar10", "highres"):
        if arch.startswith(prefix):
            arch, dataset = arch[len(prefix):].lstrip("_"), prefix
    extra = {}
    if a.config_path:
        extra["config_path"] = a.config_path
    else:
        extra["stages"] = "%dx1" % world
    for flag in ("no_input_pipelining", "macrobatch", "forward_only"):
        if getattr(a, flag):
            extra[flag] = "1"
    if a.lr_policy:
        extra["lr_policy"] = a.lr_policy
    if a.lr_warmup:
        extra["lr_warmup"] = "5"
    if a.checkpoint_dir:
        extra["checkpoint_dir"] = a.checkpoint_dir
    data = {}
    if a.data_dir and os.path.isdir(a.data_dir):
        data = dict(datadir=a.data_dir, synthetic=False)
    kw = dict(benchmark=dataset, framework="pipedream", arch=arch, gpus=world, batch_size=a.batch_size, **data,
              epochs=a.epochs, lr=a.lr, momentum=a.momentum, weight_decay=a.wd, log_interval=a.print_freq,
              steps_per_epoch=a.num_minibatches, extra=extra,
              device=("cpu" if a.distributed_backend == "gloo" else "auto"))
    run_worker(RunConfig.from_env(**{k: v for k, v in kw.items() if v is not None}))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
