"""Script-level entry points with the command line of the reference's nine harness scripts
(``benchmark/<dataset>/<dataset>_{pytorch,horovod,gpipe}.py``): ``-a <arch>``, ``-s <int>``
(bash convention: 0 = synthetic data, anything else = real data from ``$DATADIR``; for the ImageNet
scripts ``-1`` selects the 512x512 HIGHRES set, ``imagenet_pytorch.py:83-96``), ``--lr``,
``--momentum``, ``--wd`` — and the same environment contract (``BATCH_SIZE, EPOCHS, LOGINTER,
CORES_GPU, DATADIR, MICROBATCHES``).  The thin files under ``benchmark/`` call :func:`main`, so

    python benchmark/imagenet/imagenet_pytorch.py -a resnet50 -s 0
    horovodrun-style:  torchrun --nproc-per-node 8 benchmark/imagenet/imagenet_horovod.py -a resnet50

keep working unchanged."""
from __future__ import annotations

import argparse
from typing import Optional, Sequence

from .config import RunConfig


def parse(dataset: str, argv: Optional[Sequence[str]] = None):
    ap = argparse.ArgumentParser(description="D-DNN %s benchmark" % dataset)
    ap.add_argument("-a", "--arch", default="resnet50")
    ap.add_argument("--lr", "--learning-rate", type=float, default=None, dest="lr")
    ap.add_argument("--momentum", type=float, default=None)
    ap.add_argument("--wd", "--weight-decay", type=float, default=None, dest="weight_decay")
    ap.add_argument("-s", "--synthetic_data", type=int, default=0)
    ap.add_argument("--warmup-epochs", type=float, default=None,
                    help="Horovod scripts: epochs of linear LR warm-up from lr to lr*world (default 1)")
    ap.add_argument("--device", default="auto")
    return ap.parse_args(argv)


def config_from_args(dataset: str, framework: str, a) -> RunConfig:
    benchmark = dataset
    synthetic = a.synthetic_data == 0
    if dataset == "imagenet" and a.synthetic_data == -1:
        benchmark, synthetic = "highres", True          # -1: synthetic HIGHRES directories
    kw = dict(benchmark=benchmark, framework=framework, arch=a.arch, synthetic=synthetic,
              device=a.device, lr=a.lr, momentum=a.momentum, weight_decay=a.weight_decay)
    if a.warmup_epochs is not None:
        kw["extra"] = {"warmup_epochs": str(a.warmup_epochs)}
    return RunConfig.from_env(**{k: v for k, v in kw.items() if v is not None})


def main(dataset: str, framework: str, argv: Optional[Sequence[str]] = None) -> dict:
    from .cli import run_worker
    cfg = config_from_args(dataset, framework, parse(dataset, argv))
    return run_worker(cfg)
