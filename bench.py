#!/usr/bin/env python
"""Headline benchmark: ResNet-50, ImageNet-shaped synthetic input, images/sec.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--config NAME]

N=1 runs in-process; for N>1 launch one rank per GPU with
``python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1
--master-port P bench.py --gpus N --steps K --warmup W`` (RANK / LOCAL_RANK / WORLD_SIZE are
read from the environment).  Rank 0 prints ONE JSON line.

``--config dp`` (default, BASELINE.json config 2): per-GPU batch 32 (weak scaling, global batch
32*N), SGD momentum 0.9, lr 0.0125*N, wd 5e-5 for N>1 (Horovod script) / lr 0.1, wd 1e-4 for N=1
(PyTorch script) — reference ``run/run/run_template.sh:196,223`` — bf16 compute with fp32 master
weights, random-init weights.

* ``value`` — device-timed (CUDA events on the launching stream, barrier + synchronize on both
  sides, MAX over ranks) over EXACTLY K steps of the captured training step with the uint8 batch
  already on the device.  ``long_run`` repeats the same measurement over >= 200 steps (the K-step
  region is ~0.1 s; box-to-box noise is a few %).
* ``e2e`` — the same metric end to end through the public API, LIKE FOR LIKE with the reference
  arm: the SAME synthetic JPEG tree (``baseline/jpeg_tree.py``) is read from disk, decoded and
  augmented inside the timed loop (``data.JpegFolder``; decode worker processes, BENCH_DECODE_PROCS=0
  for the thread pool), every step's uint8 batch is copied from pinned host memory (copy stream) and
  the loss is read back to the host.
  ``e2e_pinned_u8`` is the same loop fed from pre-decoded pinned batches (input pipeline excluded).
* ``extra.torch_bf16_graph`` — stock PyTorch (cuDNN/cuBLAS, torch.optim.SGD) bf16 channels-last
  under a whole-step CUDA graph, measured in the same process: the iso-precision comparator.
* ``checks`` (N>1) — the fused all-reduce kernel (peer loads/stores and NVLS multimem) against
  ``dist.all_reduce`` on a seeded buffer before timing, and the spread of the weights across ranks
  after timing; a failure makes the run exit non-zero.
* ``exposed_comm`` (N>1, opt-in BENCH_EXPOSED_COMM=1, not yet run on hardware) — the step without the
  gradient all-reduce (independent replicas) and the difference to the distributed step.

Pipeline configs (BASELINE.json configs 3-5), same JSON schema:
  ``gpipe_vgg16_c10_4``  VGG-16 / CIFAR-10 shape, GPipe 4 stages, micro-batch 64 x 32 chunks
  ``pd_r50_4x2``         ResNet-50 / ImageNet shape, PipeDream 1F1B 4 stages x 2 replicas, batch 128
  ``pd_r152_8``          ResNet-152 / ImageNet shape, PipeDream 1F1B 8 stages, batch 128
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ARCH, DATASET, PER_GPU_BATCH = "resnet50", "imagenet", 32
METRIC = "images/sec (device-timed, max over ranks) ResNet-50 ImageNet at 1/2/4/8 B200"
LONG_STEPS = 200


# ------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def _time_steps(trainer, src, steps, dev, dist, n):
    """Device time of `steps` captured steps with the batch resident on the device (ms total)."""
    import torch
    mean, std = src.mean, src.std
    torch.cuda.synchronize(dev)
    if n > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    e0.record()
    for i in range(steps):
        x, t = src.batch(i)
        trainer.step_u8(x, t, mean, std)
    e1.record()
    torch.cuda.synchronize(dev)
    if n > 1:
        dist.barrier()
    return e0.elapsed_time(e1)


def _torch_baseline(dev, bs=PER_GPU_BATCH, iters=30):
    """Stock PyTorch bf16 channels-last ResNet-50 step under a CUDA graph (library kernels)."""
    import torch
    import torch.nn.functional as F
    from ddlbench_b200.models import build_model
    torch.backends.cudnn.benchmark = True
    m = build_model(ARCH, DATASET).to(dev).to(torch.bfloat16).to(memory_format=torch.channels_last)
    x = torch.randn(bs, 3, 224, 224, device=dev, dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    t = torch.randint(0, 1000, (bs,), device=dev)
    opt = torch.optim.SGD(m.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4)

    def step():
        opt.zero_grad(set_to_none=True)
        loss = F.cross_entropy(m(x).float(), t)
        loss.backward()
        opt.step()

    s = torch.cuda.Stream(device=dev)
    s.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(s):
        for _ in range(3):
            step()
    torch.cuda.current_stream(dev).wait_stream(s)
    torch.cuda.synchronize(dev)
    g = torch.cuda.CUDAGraph()
    opt.zero_grad(set_to_none=True)
    with torch.cuda.graph(g):
        step()
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / iters
    del g, m, opt
    torch.cuda.empty_cache()
    return {"ms_per_step": round(ms, 4), "value": round(bs / ms * 1e3, 2), "unit": "images/sec",
            "what": "stock PyTorch ops (cuDNN/cuBLAS, torch.optim.SGD), bf16 channels_last, whole-step CUDA "
                    "graph, batch %d, 1 GPU, input resident on the device" % bs}


def _allreduce_checks(trainer, dev, dist, world, rank):
    """Fused all-reduce kernel vs NCCL on a seeded buffer (P2P and NVLS paths)."""
    import torch
    from ddlbench_b200.dist.symm import SymmContext
    from ddlbench_b200.parallel.collectives import SymmAllreduce
    out = {}
    ctx = SymmContext(dev)
    ar = SymmAllreduce(ctx, max_bytes=32 << 20, blocks=32)
    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    for label, n in (("small", 96 * 1024), ("large", 6 << 20)):
        n -= n % (world * 8)
        base = torch.randn(n, generator=g).to(dev)
        ref = base.clone()
        dist.all_reduce(ref)
        ref /= world
        for algo in ("p2p", "auto"):
            x = base.to(torch.bfloat16)
            ar.allreduce_(x, True, algo)
            err = (x.float() - ref).abs().max().item()
            path = "p2p" if algo == "p2p" else ("oneshot" if label == "small" else
                                                ("nvls" if ctx.use_multimem else "twoshot"))
            key = "ar_max_abs_err_%s_%s" % (path, label)
            out[key] = round(err, 5)
    ctx.check()
    out["nvls_multimem"] = bool(ctx.use_multimem)
    # bf16 inputs of unit variance, sum of `world` of them, bf16 output: a few 2^-8 relative
    out["ar_ok"] = all(v < 0.08 for k, v in out.items() if k.startswith("ar_max_abs_err"))
    return out


def _exposed_comm(args, dev, n, src_dev, ms_step, lr, wd, max_over_ranks):
    """Opt-in (BENCH_EXPOSED_COMM=1): the same step WITHOUT the gradient all-reduce — an independent,
    non-distributed replica of the trainer on every rank (its own separate SGD launch included) —
    timed concurrently on all ranks; ``exposed_comm_ms`` = distributed step - this one.  No collective
    is issued inside the measured part and every rank always reaches the final reduction, so a
    failure on one rank cannot strand the others."""
    import torch
    from ddlbench_b200.harness import Trainer
    ms_local, err = float("nan"), None
    local = None
    try:
        local = Trainer(ARCH, DATASET, PER_GPU_BATCH, device=str(dev), dtype=torch.bfloat16, lr=lr,
                        momentum=0.9, weight_decay=wd, cuda_graph=not args.no_graph, distributed=False,
                        tcgen05=not args.no_tcgen05)
        mean, std = src_dev.mean, src_dev.std
        for i in range(max(args.warmup, 3)):
            local.step_u8(*src_dev.batch(i), mean, std)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        e0.record()
        for i in range(args.steps):
            local.step_u8(*src_dev.batch(i), mean, std)
        e1.record()
        torch.cuda.synchronize(dev)
        ms_local = e0.elapsed_time(e1) / args.steps
    except Exception as e:  # noqa: BLE001
        err = "%s: %s" % (type(e).__name__, str(e)[:200])
    finally:
        del local
        torch.cuda.empty_cache()
    failed = max_over_ranks(0.0 if err is None else 1.0, dev)            # every rank gets here
    ms_max = max_over_ranks(ms_local if err is None else -1.0, dev)
    if failed > 0:
        return {"error": err or "failed on another rank"}
    return {"ms_per_step_without_allreduce": round(ms_max, 4),
            "exposed_comm_ms": round(ms_step - ms_max, 4),
            "note": "independent non-distributed replicas (separate SGD launch) timed concurrently on all "
                    "ranks, max over ranks; negative = the fused all-reduce + SGD step is shorter than a "
                    "local step with its own optimizer launch"}


def _weight_spread(trainer, dev, dist):
    import torch
    w = trainer.ex.store.w_flat[trainer.ex.store.latest].float()
    lo, hi = w.clone(), w.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    return float((hi - lo).abs().max().item())


# ------------------------------------------------------------------------------ ours: DP
def run_ours(args) -> dict:
    import torch
    import torch.distributed as dist
    from ddlbench_b200 import ops
    from ddlbench_b200.data.pipeline import DevicePrefetcher
    from ddlbench_b200.data.synthetic import SyntheticImages
    from ddlbench_b200.harness import Trainer, timed_loop
    from ddlbench_b200.harness.loop import max_over_ranks
    from ddlbench_b200.parallel import dp

    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        dp.init()
    rank = dp.rank()
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world != args.gpus and rank == 0:
        print(f"[bench] warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)
    n = world
    lr, wd = (0.1, 1e-4) if n == 1 else (0.0125 * n, 5e-5)
    checks = {}
    if n > 1 and args.allreduce == "fused":
        checks.update(_allreduce_checks(None, dev, dist, n, rank))
    trainer = Trainer(ARCH, DATASET, PER_GPU_BATCH, device=str(dev), dtype=torch.bfloat16, lr=lr,
                      momentum=0.9, weight_decay=wd, cuda_graph=not args.no_graph,
                      distributed=n > 1, allreduce=args.allreduce, tcgen05=not args.no_tcgen05,
                      comm_blocks=args.comm_blocks)
    src_dev = SyntheticImages(DATASET, PER_GPU_BATCH, num_batches=4, device=dev, seed=1 + rank)
    mean, std = src_dev.mean, src_dev.std

    # ---- kernel-only (inputs resident on the device): W warm-up, then EXACTLY K steps
    for i in range(args.warmup):
        x, t = src_dev.batch(i)
        trainer.step_u8(x, t, mean, std)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ops.launches
    ms = _time_steps(trainer, src_dev, args.steps, dev, dist, n)
    eager_launches = ops.launches - launches0
    ms_step = max_over_ranks(ms, dev) / args.steps
    value = PER_GPU_BATCH * n / (ms_step / 1e3)
    trainer.check()
    # the same, over a region long enough to average out scheduling noise
    long_steps = max(LONG_STEPS, args.steps)
    ms_long = max_over_ranks(_time_steps(trainer, src_dev, long_steps, dev, dist, n), dev) / long_steps
    clocks = sampler.stop() if rank == 0 else {}
    trainer.check()

    # ---- end to end, pre-decoded pinned uint8 batches (input pipeline excluded)
    src_host = SyntheticImages(DATASET, PER_GPU_BATCH, num_batches=8, device="cpu", pinned=True,
                               seed=11 + rank)
    pre = DevicePrefetcher(src_host, dev)
    ms_pin, losses = timed_loop(trainer, pre, args.steps, max(args.warmup, 3), read_loss=True)
    ms_pin = max_over_ranks(ms_pin, dev)
    trainer.check()

    # ---- end to end through the JPEG tree (the like-for-like number against the reference arm)
    e2e = _e2e_jpeg(args, trainer, dev, rank, n, dist, timed_loop, max_over_ranks, DevicePrefetcher)
    e2e_nvjpeg = None
    if os.environ.get("BENCH_NVJPEG", "0") == "1":
        e2e_nvjpeg = _e2e_jpeg(args, trainer, dev, rank, n, dist, timed_loop, max_over_ranks, DevicePrefetcher,
                               gpu_decode=True)
    if n > 1:
        checks["weight_spread"] = _weight_spread(trainer, dev, dist)
        checks["weights_ok"] = checks["weight_spread"] == 0.0

    exposed = None
    if n > 1 and os.environ.get("BENCH_EXPOSED_COMM", "0") == "1":
        exposed = _exposed_comm(args, dev, n, src_dev, ms_step, lr, wd, max_over_ranks)

    extra = {}
    if rank == 0 and not args.no_torch_baseline:
        try:
            extra["torch_bf16_graph"] = _torch_baseline(dev)
            extra["vs_torch_bf16_graph_n1"] = None if n > 1 else round(
                value / extra["torch_bf16_graph"]["value"], 3)
        except Exception as e:  # noqa: BLE001 - informational, never breaks the measurement
            extra["torch_bf16_graph"] = {"error": str(e)[:200]}
    if n > 1:
        dist.barrier()

    mem_gb = torch.cuda.max_memory_allocated(dev) / 1e9
    out = {
        "metric": METRIC, "value": round(value, 2), "unit": "images/sec", "n_gpus": n,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_step, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic (random uint8 images of the ImageNet shape, random labels, "
                "random-init weights); e2e: synthetic JPEG tree decoded inside the loop",
        "impl": "ours",
        "config": {"model": ARCH, "global_batch": PER_GPU_BATCH * n, "per_gpu_batch": PER_GPU_BATCH,
                   "seq_len": None, "input": "3x224x224",
                   "parallelism": ("single" if n == 1 else f"dp{n}"),
                   "optimizer": f"SGD momentum 0.9 lr {lr} wd {wd}, fp32 master + bf16 weights",
                   "allreduce": ("none" if n == 1 else trainer.allreduce),
                   "cuda_graph": trainer.cuda_graph, "tcgen05": not args.no_tcgen05,
                   "conv_kernels": "hand-written tcgen05 implicit GEMM (fprop, dgrad, wgrad)"
                                   if ops.IGEMM and not args.no_tcgen05 else "library",
                   "l2": "no flush: per-step working set (weights+grads+master+activations, "
                         ">1 GB) exceeds the 126 MB L2 and inputs cycle over distinct batches",
                   "timing": "CUDA events on the launching stream, barrier+synchronize both "
                             "sides, max over ranks",
                   "switches": {k: v for k, v in sorted(os.environ.items()) if k.startswith("DDLB_")}},
        "clocks": clocks,
        "long_run": {"steps": long_steps, "ms_per_step": round(ms_long, 4),
                     "value": round(PER_GPU_BATCH * n / (ms_long / 1e3), 2)},
        "e2e": e2e,
        "e2e_pinned_u8": {"value": round(PER_GPU_BATCH * n / (ms_pin / 1e3), 2), "unit": "images/sec",
                          "ms_per_step": round(ms_pin, 4), "h2d_bytes_per_step": pre.h2d_bytes,
                          "d2h_bytes_per_step": 12, "last_loss": (losses[-1] if losses else None)},
        "gpu_launches": int(trainer.launches_per_step * args.steps + eager_launches),
        "launches_per_step_in_graph": int(trainer.launches_per_step),
        "peak_mem_gb": round(mem_gb, 3),
        "extra": extra,
    }
    if checks:
        out["checks"] = checks
    if exposed is not None:
        out["exposed_comm"] = exposed
    if e2e_nvjpeg is not None:
        out["e2e_nvjpeg"] = e2e_nvjpeg
    if rank == 0:
        try:   # informational; must never break the measurement
            from ddlbench_b200.models import build_model
            from ddlbench_b200.utils.cost import roofline
            out["roofline"] = roofline(build_model(ARCH, DATASET), PER_GPU_BATCH, ms_step)
        except Exception as e:  # pragma: no cover
            out["roofline"] = {"error": str(e)[:200]}
    failed = bool(checks) and not (checks.get("ar_ok", True) and checks.get("weights_ok", True))
    if n > 1:
        dist.barrier()
        dp.shutdown()
    if failed:
        if rank == 0:
            print(json.dumps(out), flush=True)
            print("[bench] multi-GPU correctness check FAILED: %s" % checks, file=sys.stderr)
        raise SystemExit(3)
    return out if rank == 0 else {}


def _e2e_jpeg(args, trainer, dev, rank, n, dist, timed_loop, max_over_ranks, DevicePrefetcher, gpu_decode=False):
    """Train from a synthetic JPEG ImageFolder tree: decode + augmentation + H2D + loss read-back
    inside the timed region (what the reference's own metric contains)."""
    import torch
    from ddlbench_b200.data.imagefolder import JpegFolder
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    from jpeg_tree import make_tree
    steps = max(args.steps, 50)
    warm = max(args.warmup, 5)
    tmp = os.path.join(tempfile.gettempdir(), "ddlb_jpeg_%s" % os.environ.get("MASTER_PORT", str(os.getpid())))
    ready = os.path.join(tmp, ".ready")
    t0 = time.time()
    if rank == 0:
        shutil.rmtree(tmp, ignore_errors=True)
        os.makedirs(tmp, exist_ok=True)
        make_tree(tmp, (steps + warm) * PER_GPU_BATCH * n, PER_GPU_BATCH * n)
        open(ready, "w").write("ok")
    else:
        while not os.path.exists(ready):
            time.sleep(0.1)
    gen_s = time.time() - t0
    try:
        cores = len(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        cores = os.cpu_count() or 8
    workers = max(2, min(int(os.environ.get("BENCH_DECODE_THREADS", "32")), cores // n))
    # decode in worker PROCESSES (data/decode_worker.py): the thread pool is GIL-bound near 900-2000
    # images/s whatever the core count; BENCH_DECODE_PROCS=0 keeps the threads.  The reference arm gets
    # up to 24 DataLoader worker processes per GPU (baseline/run_reference.py::_cores); we take the
    # cores that are there, minus two for the training loop and the copy thread
    procs = int(os.environ.get("BENCH_DECODE_PROCS", str(max(0, min(64, cores // n - 2)))))
    if gpu_decode or procs < 6:            # a handful of processes is no better than the thread pool
        procs = 0
    try:
        src = JpegFolder(tmp, DATASET, PER_GPU_BATCH, split="train", shuffle=True, seed=3, rank=rank,
                         world=n, workers=workers, pinned=True, prefetch=8 if procs else 6,
                         gpu_decode=gpu_decode, device=dev, decode_procs=procs)
        pre = DevicePrefetcher(src, dev)
        ms, losses = timed_loop(trainer, pre, steps, warm, read_loss=True)
        ms = max_over_ranks(ms, dev)
        src.close()
        out = {"value": round(PER_GPU_BATCH * n / (ms / 1e3), 2), "unit": "images/sec",
               "ms_per_step": round(ms, 4), "steps": steps, "warmup": warm,
               "h2d_bytes_per_step": pre.h2d_bytes, "d2h_bytes_per_step": 12,
               "last_loss": (losses[-1] if losses else None),
               "input": "synthetic JPEG tree (baseline/jpeg_tree.py, the one the reference arm reads), "
                        "decoded + RandomResizedCrop/flip by data.JpegFolder on %s inside "
                        "the timed loop; uint8 batch H2D from pinned memory every step; loss read back" % (
                            ("%d worker processes" % procs) if (procs and src._procs_error is None)
                            else ("%d host threads" % workers)),
               "decode_threads": workers, "decode_procs": (procs if src._procs_error is None else 0),
               "decode_procs_error": src._procs_error, "host_cores": cores,
               "datagen_seconds": round(gen_s, 1),
               "decode": "nvJPEG on the device (torchvision.io.decode_jpeg)" if gpu_decode else "host (PIL/libjpeg)"}
    except Exception as e:  # noqa: BLE001
        out = {"value": None, "error": "%s: %s" % (type(e).__name__, str(e)[:200])}
    if n > 1:
        dist.barrier()
    if rank == 0:
        shutil.rmtree(tmp, ignore_errors=True)
    return out


# ------------------------------------------------------------------------------ ours: pipelines
PIPE = {
    # name: (framework, arch, benchmark, gpus, batch, extra, metric, parallelism)
    "gpipe_vgg16_c10_4": ("gpipe", "vgg16", "cifar10", 4, 64, {},
                          "images/sec (device-timed, max over ranks) VGG-16 CIFAR-10 GPipe 4-stage on 4 B200",
                          "gpipe 4 stages, micro-batch 64 x 32 chunks"),
    "pd_r50_4": ("pipedream", "resnet50", "imagenet", 4, 128, {"stages": "4x1"},
                 "images/sec (device-timed, max over ranks) ResNet-50 ImageNet PipeDream 1F1B 4-stage on 4 B200",
                 "pipedream 1F1B, 4 stages, weight stashing"),
    "pd_r50_4x2": ("pipedream", "resnet50", "imagenet", 8, 128, {"stages": "4x2"},
                   "images/sec (device-timed, max over ranks) ResNet-50 ImageNet PipeDream 1F1B 4x2 on 8 B200",
                   "pipedream 1F1B, 4 stages x 2 replicas, weight stashing, fused replica reduction"),
    "pd_r152_8": ("pipedream", "resnet152", "imagenet", 8, 128, {"stages": "8x1"},
                  "images/sec (device-timed, max over ranks) ResNet-152 ImageNet PipeDream 1F1B 8-stage on 8 B200",
                  "pipedream 1F1B, 8 stages, weight stashing"),
}


def run_pipeline(args) -> dict:
    import torch
    from ddlbench_b200 import ops
    from ddlbench_b200.config import RunConfig
    from ddlbench_b200.parallel import dp
    fw, arch, bench, gpus, batch, extra, metric, par = PIPE[args.config]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != gpus:
        return {"impl": "ours", "config": {"name": args.config},
                "unavailable": "config %s needs %d ranks (torchrun --nproc-per-node %d), got %d" % (
                    args.config, gpus, gpus, world)}
    dp.init()
    rank = dp.rank()
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    steps = max(args.steps, 3)
    cfg = RunConfig(benchmark=bench, framework=fw, arch=arch, gpus=gpus, epochs=2, batch_size=batch,
                    steps_per_epoch=steps, eval_steps=1, log_interval=10 ** 9, device="cuda",
                    extra=dict(extra))
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ops.launches
    if fw == "gpipe":
        from ddlbench_b200.harness.pipeline_loop import run_gpipe
        res = run_gpipe(cfg, verbose=False)
        per_step = batch * cfg.microbatches
    else:
        from ddlbench_b200.harness.pipeline_loop import run_pipedream
        res = run_pipedream(cfg, verbose=False)
        per_step = batch * int(extra["stages"].split("x")[1])
    clocks = sampler.stop() if rank == 0 else {}
    tput = res["device_throughputs"][-1]          # epoch 0 = warm-up (same number of steps), epoch 1 timed
    out = {
        "metric": metric, "value": round(tput, 2), "unit": "images/sec", "n_gpus": gpus,
        "steps": steps, "warmup": steps, "ms_per_step": round(per_step / tput * 1e3, 4),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic (random images of the data set's shape, random labels, random-init weights)",
        "impl": "ours",
        "config": {"name": args.config, "model": arch, "global_batch": per_step, "seq_len": None,
                   "parallelism": par, "dataset_shape": bench,
                   "timing": "CUDA events around the epoch's minibatch loop on every rank (pipeline fill and "
                             "drain included), max over ranks; epoch 0 (same length) is the warm-up",
                   "l2": "no flush: activations of the in-flight minibatches exceed the L2",
                   "epoch_throughputs": [round(v, 1) for v in res["device_throughputs"]],
                   "switches": {k: v for k, v in sorted(os.environ.items()) if k.startswith("DDLB_")}},
        "clocks": clocks,
        "e2e": {"value": round(res["throughput"], 2), "unit": "images/sec",
                "note": "host wall clock of the epochs (first, cold epoch included) as the reference's "
                        "scripts report it; inputs are generated on the device in this configuration",
                "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 12},
        "gpu_launches": int(ops.launches - launches0),
        "peak_mem_gb": round(res.get("peak_mem_gb", 0.0), 3),
    }
    dp.shutdown()
    return out if rank == 0 else {}


# ------------------------------------------------------------------------------ reference
def run_reference(args) -> dict:
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import run_reference
    if args.config != "dp":
        return {"impl": "reference",
                "unavailable": "the reference's torchgpipe / pipedream-fork runtimes need torch 1.1-1.5 builds "
                               "(pre_hook.patch, torchgpipe 0.0.5) that cannot be installed offline"}
    return run_reference.main(args.gpus, args.steps, args.warmup, METRIC)


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="dp", choices=["dp"] + sorted(PIPE))
    ap.add_argument("--allreduce", default="fused", choices=["fused", "nccl"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-tcgen05", action="store_true")
    ap.add_argument("--no-torch-baseline", action="store_true")
    ap.add_argument("--comm-blocks", type=int, default=32)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        out = run_reference(args)
    elif args.config == "dp":
        out = run_ours(args)
    else:
        out = run_pipeline(args)
    if out:
        print(json.dumps(out), flush=True)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
