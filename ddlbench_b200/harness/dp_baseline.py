"""The upstream runtime's non-pipelined data-parallel baseline
(``pipedream-fork/runtime/image_classification/main.py:24-379``: ``nn.DataParallel`` /
``DistributedDataParallel`` + apex amp over the generated ``full_model()``, with its in-memory
``SyntheticDataset``, ``--forward_only``, ``--num_minibatches``, ``--resume`` and the
``Epoch: [e][i/n]  Time … Data … Memory … Loss … Prec@1 … Prec@5`` log lines that
``scripts/process_output.py`` parses).

Here the full model is the layer graph of ``--module`` (``models.<arch>.gpus=<N>`` as the code
generator names it, a generated package on ``sys.path`` exposing ``arch()``, or a plain
architecture name) run by the engine: one process per GPU, whole step in a CUDA graph, gradients
reduced by the fused all-reduce + SGD kernel when launched with more than one rank (torchrun /
``--world-size`` + ``--rank``), bf16 compute with fp32 master weights where the reference offers
``--fp16`` through amp.

    python -m ddlbench_b200.harness.dp_baseline -m models.resnet50.gpus=4 -b 64 -s --epochs 1 \\
        --num_minibatches 100
    python pipedream-fork/runtime/image_classification/main.py -m models.vgg16.gpus=4 -b 64 -s   # same thing
"""
from __future__ import annotations

import argparse
import importlib
import os
import re
import shutil
import time
from typing import Dict, List, Optional

import torch
import torch.distributed as dist

from ..data.synthetic import SyntheticImages
from .trainer import Trainer

__all__ = ["main", "parse_args"]


def parse_args(argv: Optional[List[str]] = None):
    ap = argparse.ArgumentParser(description="data-parallel baseline (upstream main.py flags)")
    ap.add_argument("--data_dir", type=str, default=None)
    ap.add_argument("--module", "-m", required=True)
    ap.add_argument("-j", "--workers", default=4, type=int)
    ap.add_argument("--epochs", default=90, type=int)
    ap.add_argument("--start-epoch", default=0, type=int, dest="start_epoch")
    ap.add_argument("-b", "--batch-size", default=256, type=int, dest="batch_size")
    ap.add_argument("--eval-batch-size", default=100, type=int, dest="eval_batch_size")
    ap.add_argument("--lr", "--learning-rate", default=0.1, type=float, dest="lr")
    ap.add_argument("--momentum", default=0.9, type=float)
    ap.add_argument("--weight-decay", "--wd", default=1e-4, type=float, dest="weight_decay")
    ap.add_argument("--print-freq", "-p", default=10, type=int, dest="print_freq")
    ap.add_argument("--fp16", action="store_true", help="bf16 compute + fp32 master weights (the default on GPUs)")
    ap.add_argument("--forward_only", action="store_true")
    ap.add_argument("--num_minibatches", default=None, type=int)
    ap.add_argument("--resume", default="", type=str)
    ap.add_argument("-s", "--synthetic_data", action="store_true")
    ap.add_argument("--world-size", default=None, type=int, dest="world_size")
    ap.add_argument("--rank", default=None, type=int)
    ap.add_argument("--dist-url", default="tcp://127.0.0.1:23456", type=str, dest="dist_url")
    ap.add_argument("--dist-backend", default=None, type=str, dest="dist_backend")
    # extensions
    ap.add_argument("--dataset", default=os.environ.get("BENCHMARK", "imagenet"))
    ap.add_argument("--device", default="auto")
    ap.add_argument("--checkpoint_dir", default=".", help="where checkpoint.pth.tar / model_best.pth.tar go")
    ap.add_argument("--no-graph", action="store_true")
    return ap.parse_args(argv)


def resolve_module(module: str, dataset: str):
    """-> (arch, dataset).  ``models.<arch>.gpus=<N>`` / ``<prefix><arch>`` naming of the code
    generator, an importable generated package with ``arch()``, or an architecture name."""
    m = re.match(r"(?:.*\.)?models\.([A-Za-z0-9_]+)\.gpus=\d+$", module)
    arch = m.group(1) if m else None
    if arch is None:
        try:
            arch = importlib.import_module(module).arch()
        except Exception:  # noqa: BLE001 - not a package: take the last component as the name
            arch = module.split(".")[-1]
    for prefix in ("mnist", "cifar10", "highres"):
        if arch.startswith(prefix) and len(arch) > len(prefix):
            arch, dataset = arch[len(prefix):].lstrip("_"), prefix
    return arch, dataset


class _Meter:
    def __init__(self):
        self.val = self.sum = self.avg = 0.0
        self.count = 0

    def update(self, v: float, n: int = 1):
        self.val = v
        self.sum += v * n
        self.count += n
        self.avg = self.sum / max(self.count, 1)


def _mem():
    if torch.cuda.is_available():
        return torch.cuda.memory_allocated() / 1e9, torch.cuda.memory_reserved() / 1e9
    return 0.0, 0.0


def _train_line(epoch, i, n, bt, dt, loss, top1, top5):
    a, c = _mem()
    return ("Epoch: [{0}][{1}/{2}]\tTime: {3:.3f} ({4:.3f})\tData: {5:.3f} ({6:.3f})\t"
            "Memory: {7:.3f} ({8:.3f})\tLoss: {9:.4f} ({10:.4f})\tPrec@1: {11:.3f} ({12:.3f})\t"
            "Prec@5: {13:.3f} ({14:.3f})").format(epoch, i, n, bt.val, bt.avg, dt.val, dt.avg, a, c, loss.val,
                                                   loss.avg, top1.val, top1.avg, top5.val, top5.avg)


def _test_line(epoch, i, n, bt, loss, top1, top5):
    a, c = _mem()
    return ("Test: [{0}][{1}/{2}]\tTime: {3:.3f} ({4:.3f})\tMemory: {5:.3f} ({6:.3f})\t"
            "Loss: {7:.4f} ({8:.4f})\tPrec@1: {9:.3f} ({10:.3f})\tPrec@5: {11:.3f} ({12:.3f})").format(
        epoch, i, n, bt.val, bt.avg, a, c, loss.val, loss.avg, top1.val, top1.avg, top5.val, top5.avg)


def _sources(a, arch, dataset, device, rank, world):
    if a.synthetic_data or not a.data_dir:
        # the reference's SyntheticDataset: one fixed sample repeated; here distinct random batches
        n = (a.num_minibatches or 100)
        train = SyntheticImages(dataset, a.batch_size, num_batches=min(n, 8), device=device, length=n)
        val = SyntheticImages(dataset, a.batch_size, num_batches=2, device=device, seed=2,
                              length=max(1, n // 10))
        return train, val, False
    from ..config import RunConfig
    from ..data import make_source
    cfg = RunConfig(benchmark=dataset, framework="horovod" if world > 1 else "pytorch", arch=arch,
                    batch_size=a.batch_size, datadir=a.data_dir, synthetic=False, cores_gpu=a.workers)
    return (make_source(cfg, a.batch_size, device, rank, world, "train"),
            make_source(cfg, a.batch_size, device, rank, world, "val"), True)


def _run_epoch(trainer, src, epoch, a, train: bool, say):
    from .. import ops
    bt, dt, loss, top1, top5 = _Meter(), _Meter(), _Meter(), _Meter(), _Meter()
    n = len(src)
    if a.num_minibatches is not None:
        n = min(n, a.num_minibatches)
    t_epoch = time.time()
    end = time.time()
    for i in range(n):
        img, lab = src.batch(i)
        img = img.to(trainer.device, non_blocking=True)
        lab = lab.to(trainer.device, non_blocking=True)
        dt.update(time.time() - end)
        if train:
            stats = trainer.step_u8(img, lab, src.mean, src.std)
        else:
            stats = trainer.eval_step(ops.normalize_u8(img, src.mean, src.std, dtype=trainer.dtype), lab)
        if i % a.print_freq == 0 or i == n - 1:
            s = [float(v) for v in stats.detach().float().cpu()]          # (mean loss, #top1, #top5)
            bs = lab.shape[0]
            loss.update(s[0], bs)
            top1.update(100.0 * s[1] / bs, bs)
            top5.update(100.0 * s[2] / bs, bs)
        bt.update(time.time() - end)
        end = time.time()
        if i % a.print_freq == 0:
            say(_train_line(epoch, i, n, bt, dt, loss, top1, top5) if train
                else _test_line(epoch, i, n, bt, loss, top1, top5))
    if not train:
        say(" * Prec@1 {0:.3f} Prec@5 {1:.3f}".format(top1.avg, top5.avg))
    say("Epoch %d: %.3f seconds" % (epoch, time.time() - t_epoch))
    say("Epoch start time: %.3f, epoch end time: %.3f" % (t_epoch, time.time()))
    return {"loss": loss.avg, "prec1": top1.avg, "prec5": top5.avg, "seconds": time.time() - t_epoch, "steps": n}


def save_checkpoint(state: Dict, is_best: bool, directory: str, filename: str = "checkpoint.pth.tar") -> str:
    os.makedirs(directory, exist_ok=True)
    path = os.path.join(directory, filename)
    tmp = path + ".tmp"
    torch.save(state, tmp)
    os.replace(tmp, path)
    if is_best:
        shutil.copyfile(path, os.path.join(directory, "model_best.pth.tar"))
    return path


def main(argv: Optional[List[str]] = None) -> Dict[str, float]:
    a = parse_args(argv)
    arch, dataset = resolve_module(a.module, a.dataset)
    use_cuda = a.device != "cpu" and torch.cuda.is_available()
    # ranks: torchrun's environment, or the reference's explicit --world-size / --rank + --dist-url
    world = int(os.environ.get("WORLD_SIZE", a.world_size or 1))
    rank = int(os.environ.get("RANK", a.rank or 0))
    own_group = False
    if world > 1 and not dist.is_initialized():
        backend = a.dist_backend or ("nccl" if use_cuda else "gloo")
        if "RANK" in os.environ:
            dist.init_process_group(backend)
        else:
            dist.init_process_group(backend, init_method=a.dist_url, world_size=world, rank=rank)
        own_group = True
    if use_cuda:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank % max(torch.cuda.device_count(), 1))))
    device = torch.device("cuda", torch.cuda.current_device()) if use_cuda else torch.device("cpu")
    say = print if rank == 0 else (lambda *x, **k: None)
    say("=> creating model '%s' (%s), %d process(es)" % (arch, dataset, world))
    trainer = Trainer(arch, dataset, a.batch_size, device=str(device),
                      dtype=torch.bfloat16 if use_cuda else torch.float32,
                      lr=a.lr * 1.0, momentum=a.momentum, weight_decay=a.weight_decay,
                      cuda_graph=use_cuda and not a.no_graph, distributed=world > 1,
                      allreduce="fused" if use_cuda else "nccl", tcgen05=use_cuda)
    best_prec1, start_epoch = 0.0, a.start_epoch
    if a.resume:
        if os.path.isfile(a.resume):
            say("=> loading checkpoint '{}'".format(a.resume))
            ck = torch.load(a.resume, map_location="cpu")
            trainer.ex.load_state_dict(ck["state_dict"])              # weights + BatchNorm buffers
            opt = ck.get("optimizer", {})
            if "master" in opt and "momentum" in opt:
                # full fp32 master + momentum (every rank re-slices its own shard when sharded)
                trainer.ex.store.load_full_state(opt["master"], opt["momentum"])
            start_epoch, best_prec1 = ck["epoch"], ck.get("best_prec1", 0.0)
            say("=> loaded checkpoint '{}' (epoch {})".format(a.resume, ck["epoch"]))
        else:
            say("=> no checkpoint found at '{}'".format(a.resume))
    train_src, val_src, real = _sources(a, arch, dataset, device, rank, world)
    out: Dict[str, float] = {"arch": arch, "dataset": dataset, "world": world}
    for epoch in range(start_epoch, a.epochs):
        trainer.set_lr(a.lr * (0.1 ** (epoch // 30)))           # main.py:352-357
        if real:
            train_src.set_epoch(epoch)
        if a.forward_only:
            out["val"] = _run_epoch(trainer, val_src, epoch, a, False, say)
            continue
        tr = _run_epoch(trainer, train_src, epoch, a, True, say)
        va = _run_epoch(trainer, val_src, epoch, a, False, say)
        out.update(train=tr, val=va, throughput=tr["steps"] * a.batch_size * world / max(tr["seconds"], 1e-9))
        is_best = va["prec1"] > best_prec1
        best_prec1 = max(va["prec1"], best_prec1)
        master, mom = trainer.ex.store.gather_full_state(trainer.group)      # collective when sharded
        if rank == 0:
            save_checkpoint({"epoch": epoch + 1, "arch": arch, "state_dict": trainer.ex.state_dict(),
                             "best_prec1": best_prec1,
                             "optimizer": {"master": master.detach().cpu(), "momentum": mom.detach().cpu()}},
                            is_best, a.checkpoint_dir)
    out["best_prec1"] = best_prec1
    trainer.check()
    if own_group:
        dist.barrier()
        dist.destroy_process_group()
    return out


if __name__ == "__main__":
    main()
