"""GNMT runtime entry point at the reference's path with its flags
(pipedream-fork/runtime/translation/main_with_runtime.py: ``--module models.gnmt[_large].gpus=N
--config_path <conf.json> -b <batch> --lr --lr_policy --epochs --max-length-train
--checkpoint_dir --num_minibatches -s --fp16 ...``).  One process per rank as ``pipedream_run.sh``
starts them, or under ``torch.distributed.run``; translated onto
``ddlbench_b200.harness.translation`` (``--stages`` from the configuration file / module name)."""
import argparse
import json
import os
import re
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "..")))
from ddlbench_b200.harness import translation  # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser(description="GNMT PipeDream runtime (ddlbench_b200)")
    ap.add_argument("--module", "-m", required=True)
    ap.add_argument("--data_dir", default=None)
    ap.add_argument("--distributed_backend", default="nccl")
    ap.add_argument("--epochs", type=int, default=8)
    ap.add_argument("-b", "--batch-size", type=int, default=128, dest="batch_size")
    ap.add_argument("--grad-clip", type=float, default=5.0, dest="grad_clip")
    ap.add_argument("--lr", "--learning-rate", type=float, default=None, dest="lr")
    ap.add_argument("--print-freq", "-p", type=int, default=10, dest="print_freq")
    ap.add_argument("--fp16", action="store_true")
    ap.add_argument("--master_addr", default="127.0.0.1")
    ap.add_argument("--master_port", default="29500")
    ap.add_argument("--config_path", default=None)
    ap.add_argument("--rank", type=int, default=None)
    ap.add_argument("--local_rank", type=int, default=0)
    ap.add_argument("--num_minibatches", type=int, default=None)
    ap.add_argument("--checkpoint_dir", default=None)
    ap.add_argument("--max-length-train", type=int, default=50, dest="max_len")
    ap.add_argument("--hidden", type=int, default=1024, help="extension: model width (reference: 1024)")
    for flag in ("-j", "--workers", "--start-epoch", "--eval-batch-size", "--lr_policy", "--momentum",
                 "--weight-decay", "--wd", "--loss_scale", "--resume", "-v", "--verbose_frequency",
                 "--num_ranks_in_server", "--min-length-train"):
        ap.add_argument(flag, default=None)              # accepted; no counterpart needed here
    for flag in ("--no_input_pipelining", "--forward_only", "--checkpoint_dir_not_nfs", "-s", "--synthetic_data",
                 "--no-bucketing", "--recompute", "--macrobatch"):
        ap.add_argument(flag, action="store_true")
    a = ap.parse_args(argv)
    m = re.match(r"models\.(gnmt_large|gnmt)\.gpus=(\d+)", a.module)
    arch, world = (m.group(1), int(m.group(2))) if m else ("gnmt", 1)
    stages = world
    if a.config_path:
        conf = json.load(open(a.config_path))
        stages = len(conf["stage_to_rank_map"])
        world = sum(len(v) for v in conf["stage_to_rank_map"].values())
    if a.rank is not None:
        os.environ.update(RANK=str(a.rank), LOCAL_RANK=str(a.local_rank), WORLD_SIZE=str(world),
                          MASTER_ADDR=a.master_addr, MASTER_PORT=str(a.master_port))
    args = ["--num-layers", "8" if arch == "gnmt_large" else "4", "--hidden", str(a.hidden),
            "-b", str(a.batch_size), "--epochs", str(a.epochs), "--grad-clip", str(a.grad_clip),
            "--print-freq", str(a.print_freq), "--max-len", str(a.max_len), "--stages", str(stages)]
    if a.data_dir and not a.synthetic_data:
        args += ["--data-dir", a.data_dir]
    if a.lr is not None:
        args += ["--lr", str(a.lr)]
    if a.fp16:
        args += ["--dtype", "bf16"]
    if a.num_minibatches is not None:
        args += ["--steps", str(a.num_minibatches)]
    if a.checkpoint_dir:
        args += ["--checkpoint-dir", a.checkpoint_dir]
    if a.distributed_backend == "gloo":
        args += ["--device", "cpu"]
    return translation.main(args)


if __name__ == "__main__":
    main()
