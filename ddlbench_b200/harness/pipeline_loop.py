"""Epoch loops of the `gpipe` and `pipedream` benchmarks (one process per GPU, launched by
``ddlbench_b200.cli`` / torchrun).

gpipe: reference ``benchmark/imagenet/imagenet_gpipe.py`` — BATCH_SIZE is the micro-batch, the
mini-batch is BATCH_SIZE x MICROBATCHES, ``balance_by_time`` picks the partition sizes, one
SGD step per mini-batch; same log lines as the single-GPU script.

pipedream: reference flow ``run/run/run_template.sh:399-558`` — profile, partition
(hierarchical optimizer), generate the model package + confs, then run
``main_with_runtime.py``; here rank 0 does the first three steps in-process and broadcasts the
stage assignment, and every rank runs its ``StageRuntime``.  Log lines follow
``main_with_runtime.py:458-500,552-566``.
"""
from __future__ import annotations

import os
import time
from typing import Dict, Optional

import torch
import torch.distributed as dist

from ..config import RunConfig
from ..data.synthetic import SyntheticImages
from ..models import build_model
from . import logfmt
from .loop import StepTimer, max_over_ranks
from .lr import pipedream_lr


def _device(cfg: RunConfig) -> torch.device:
    if cfg.device != "cpu" and torch.cuda.is_available():
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def _float_batch(src: SyntheticImages, i: int, device, dtype):
    from .. import ops
    img, lab = src.batch(i)
    x = ops.normalize_u8(img.to(device, non_blocking=True), src.mean, src.std, dtype=dtype)
    return x, lab.to(device, non_blocking=True)


# ------------------------------------------------------------------------------- GPipe
def run_gpipe(cfg: RunConfig, verbose: bool = True) -> Dict[str, float]:
    from ..parallel.gpipe import GPipe, balance_by_size, balance_by_time
    say = print if verbose else (lambda *a, **k: None)
    device = _device(cfg)
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    parts = max(world, 1) if world > 1 else max(cfg.gpus, 1)
    say(logfmt.creating_model(cfg.arch))
    torch.manual_seed(cfg.seed)
    model = build_model(cfg.arch, cfg.benchmark)
    micro, chunks = cfg.batch_size, cfg.microbatches
    batch = micro * chunks
    dtype = torch.bfloat16 if (device.type == "cuda" and cfg.dtype == "bf16") else torch.float32
    # balance on rank 0 (device-timed), shared with everyone
    if rank == 0:
        try:
            sample = torch.randn((micro,) + tuple(model.input_shape))
            balance = balance_by_time(parts, model, sample, device=device, dtype=dtype, trials=2)
        except Exception as e:  # noqa: BLE001
            say("balance_by_time failed (%s); falling back to balance_by_size" % e)
            balance = balance_by_size(parts, model)
    else:
        balance = None
    if world > 1:
        obj = [balance]
        dist.broadcast_object_list(obj, src=0)
        balance = obj[0]
    say("balance: %s" % balance)
    devices = None if world > 1 else [device] * parts
    pipe = GPipe(model, balance, chunks=chunks, devices=devices, dtype=dtype, lr=cfg.lr,
                 momentum=cfg.momentum, weight_decay=cfg.weight_decay, tcgen05=cfg.fused)
    pipe.set_micro_batch(micro)
    first, last = rank == 0, rank == parts - 1 or world == 1
    from ..data import make_source
    src = make_source(cfg, batch, device, 0, 1, "train", num_batches=2)
    # real-data mode validates on the val split (the synthetic source has none)
    val_src = src if not hasattr(src, "set_epoch") else make_source(cfg, batch, device, 0, 1, "val", num_batches=2)
    steps = len(src)
    if cfg.steps_per_epoch:
        steps = min(steps, cfg.steps_per_epoch)
    throughputs, elapsed, dev_tputs = [], [], []
    vacc = 0.0
    for epoch in range(cfg.epochs):
        if hasattr(src, "set_epoch"):
            src.set_epoch(epoch)
        timer = StepTimer(device)
        tick = time.time()
        timer.start()
        trained, loss_sum = 0, torch.zeros(1, device=device)
        for i in range(steps):
            x, t = _float_batch(src, i, device, dtype)
            st = pipe.train_step(x if (first or world == 1) else None, t if last else None)
            trained += batch
            if st is not None:
                loss_sum += st[0].detach() * batch
            if i % cfg.log_interval == 0:
                say(logfmt.train_progress(epoch + 1, cfg.epochs, i / steps * 100,
                                          trained / max(time.time() - tick, 1e-9), device))
        ms = timer.stop()
        tock = time.time()
        ms_max = max_over_ranks(ms, device)
        thr = trained / max(tock - tick, 1e-9)
        dev_tputs.append(trained / (ms_max / 1e3))
        vloss = vacc = 0.0
        nval = cfg.eval_steps if cfg.eval_steps is not None else max(1, steps // 10)
        n, correct, vl = 0, 0.0, 0.0
        pipe.set_micro_batch(batch)
        for i in range(nval):
            x, t = _float_batch(val_src, i, device, dtype)
            st = pipe.eval_step(x if (first or world == 1) else None, t if last else None)
            if st is not None:
                vl += float(st[0]) * batch
                correct += float(st[1])
            n += batch
        pipe.set_micro_batch(micro)
        vloss, vacc = vl / max(n, 1), correct / max(n, 1)
        if last:
            say(logfmt.epoch_summary(epoch + 1, cfg.epochs, float(loss_sum.item()) / max(trained, 1),
                                     thr, vloss, vacc))
        say(logfmt.device_timed(dev_tputs[-1], ms_max / max(steps, 1), parts))
        throughputs.append(thr)
        elapsed.append(tock - tick)
    k = max(len(throughputs), 1)
    out = {"throughput": sum(throughputs) / k, "sec_per_epoch": sum(elapsed) / k,
           "device_throughput": sum(dev_tputs) / k, "device_throughputs": list(dev_tputs),
           "valid_accuracy": vacc,
           "peak_mem_gb": logfmt.memory_gb(device if device.type == "cuda" else None)[0]}
    say(logfmt.final_summary(vacc, out["throughput"], out["sec_per_epoch"]))
    return out


# ------------------------------------------------------------------------------- PipeDream
def plan_pipedream(cfg: RunConfig, model, world: int, device, dtype, out_dir: Optional[str],
                   straight: bool = False, say=print):
    """profile -> partition -> (codegen) on the calling rank.  Returns (stage id per layer,
    replicas per stage)."""
    from ..partition import codegen, optimizer
    from ..partition.profiler import profile_model, stage_ids_from_graph
    prof_dir = os.path.join(out_dir, "profiles", cfg.arch) if out_dir else None
    g = profile_model(cfg.arch, cfg.benchmark, cfg.batch_size, device=device, dtype=dtype,
                      trials=2, out_dir=prof_dir, model=model, verbose=False)
    bw = float(cfg.extra.get("bandwidth", 9e11 if device.type == "cuda" else 1e9))
    mem = float(cfg.extra.get("memory", 180e9))
    part_dir = os.path.join(out_dir, "partitioned", cfg.arch) if out_dir else None
    pp, dp_t, reps, gr = optimizer.main([world], None, [bw], memory_size=mem,
                                        straight_pipeline=straight, use_memory_constraint=True,
                                        output_directory=part_dir, graph=g, out=say)
    if out_dir:
        codegen.generate(gr, cfg.arch, os.path.join(out_dir, "models", cfg.arch, "gpus=%d" % world),
                         stage_to_num_ranks=reps)
    # replay the runtime's 1F1B schedule with the profiled stage costs: the partitioner's
    # objective is a steady-state bound, this is what the schedule can actually deliver
    try:
        from ..partition.simulate import simulate_1f1b
        per_stage: Dict[int, list] = {}
        for node in gr.nodes.values():
            t = per_stage.setdefault(node.stage_id or 0, [0.0, 0.0])
            t[0] += node.forward_compute_time
            t[1] += node.backward_compute_time
        order = sorted(per_stage)
        sim = simulate_1f1b([per_stage[k][0] for k in order], [per_stage[k][1] for k in order],
                            replicas=list(reps)[:len(order)], minibatches=64)
        print("Simulated 1F1B steady state: %.3f ms per minibatch (bound %.3f ms), worst bubble %.0f%%"
              % (sim["steady_ms_per_minibatch"], sim["bound_ms"],
                 100.0 * max(sim["bubble_fraction"].values())), flush=True)
    except Exception as e:  # noqa: BLE001 - advisory only
        say("schedule simulation skipped: %s" % e)
    return stage_ids_from_graph(model, gr), reps


def run_pipedream(cfg: RunConfig, verbose: bool = True) -> Dict[str, float]:
    from ..parallel.pipedream.config import PipelineConfig
    from ..parallel.pipedream.runtime import StageRuntime, reorder_by_stage
    say = print if verbose else (lambda *a, **k: None)
    device = _device(cfg)
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    dtype = torch.bfloat16 if (device.type == "cuda" and cfg.dtype == "bf16") else torch.float32
    say(logfmt.creating_model(cfg.arch))
    torch.manual_seed(cfg.seed)
    model = build_model(cfg.arch, cfg.benchmark)
    conf_path = cfg.extra.get("config_path")
    stage_ids = reps = None
    if rank == 0:
        if cfg.extra.get("stages"):
            # explicit request: N equal-time stages x given replication ("4x2" = 4 stages, 2 replicas)
            ns, nr = (int(v) for v in str(cfg.extra["stages"]).split("x"))
            from ..parallel.gpipe import balance_by_size, balance_by_time
            # stages of equal TIME (profiled forward+backward of every layer on this device with a
            # minibatch of the run's shape); parameter bytes only as the fallback — they put most of
            # a ResNet's compute (the early, wide-activation layers) into the first stage
            try:
                if os.environ.get("DDLB_STAGE_BALANCE", "time") == "size":
                    raise RuntimeError("DDLB_STAGE_BALANCE=size")      # reproducible split (tests, resume)
                sample = torch.randn((min(cfg.batch_size, 32),) + tuple(model.input_shape))
                sizes = balance_by_time(ns, model, sample, device=device, dtype=dtype, trials=2)
            except Exception as e:  # noqa: BLE001
                say("balance_by_time failed (%s); falling back to balance_by_size" % e)
                sizes = balance_by_size(ns, model)
            stage_ids = [s for s, n in enumerate(sizes) for _ in range(n)]
            reps = [nr] * ns
        else:
            out_dir = cfg.extra.get("workdir")
            stage_ids, reps = plan_pipedream(cfg, model, world, device, dtype, out_dir,
                                             straight=cfg.extra.get("straight") == "1",
                                             say=say if cfg.extra.get("verbose_partition") else (lambda *a: None))
    if world > 1:
        obj = [stage_ids, reps]
        dist.broadcast_object_list(obj, src=0)
        stage_ids, reps = obj
    pcfg = PipelineConfig.from_json(conf_path) if conf_path else PipelineConfig.hybrid(reps)
    assert pcfg.num_ranks == world, f"configuration needs {pcfg.num_ranks} ranks, world is {world}"
    say("stage replication: %s" % reps)
    model, cuts = reorder_by_stage(model, stage_ids)
    rt = StageRuntime(model, cuts, pcfg, cfg.batch_size, device=device, dtype=dtype, lr=cfg.lr,
                      momentum=cfg.momentum, weight_decay=cfg.weight_decay,
                      no_input_pipelining=cfg.extra.get("no_input_pipelining") == "1",
                      macrobatch=cfg.extra.get("macrobatch") == "1", tcgen05=cfg.fused,
                      verbose=verbose, loss_scale=float(cfg.extra.get("loss_scale", 1.0)))
    topo = rt.topo
    from ..data import make_source
    src = make_source(cfg, cfg.batch_size, device, topo.rank_in_stage, pcfg.replicas(0), "train",
                      num_batches=2)
    # validation reads the val split in real-data mode (the synthetic source has no splits)
    val_src = src if not hasattr(src, "set_epoch") else \
        make_source(cfg, cfg.batch_size, device, topo.rank_in_stage, pcfg.replicas(0), "val", num_batches=2)
    start_epoch = 0
    if cfg.extra.get("resume"):
        from ..utils.checkpoint import load_stage_checkpoint
        payload = load_stage_checkpoint(cfg.extra["resume"], topo.stage, rt)
        if payload is None:
            raise FileNotFoundError("--resume: no checkpoint.%d.pth.tar under %s" % (topo.stage, cfg.extra["resume"]))
        start_epoch = int(payload.get("epoch", 0))
        say("=> loaded checkpoint '%s' (epoch %d)" % (cfg.extra["resume"], start_epoch))
    loader_size = len(src) if hasattr(src, "set_epoch") else len(src) // pcfg.replicas(0)
    if cfg.steps_per_epoch:
        loader_size = min(loader_size, cfg.steps_per_epoch)
    n = pcfg.num_iterations(topo.stage, loader_size)
    n_first = pcfg.num_iterations(0, loader_size)
    say("Letting in %d warm-up minibatches" % rt.warmup)
    say("Running training for %d minibatches" % n)
    epoch_times, dev_tputs = [], []
    meters = {"loss": 0.0, "top1": 0.0, "top5": 0.0, "count": 0}
    for epoch in range(start_epoch, cfg.epochs):
        if hasattr(src, "set_epoch"):
            src.set_epoch(epoch)               # a new shuffle every epoch (DistributedSampler.set_epoch)

        def loader():
            for i in range(n):
                yield _float_batch(src, i, device, dtype)

        t_epoch = time.time()
        last_t = [time.time()]
        timer = StepTimer(device)
        timer.start()

        def on_fwd(i, stats):
            rt.set_lr(pipedream_lr(cfg.lr, cfg.extra.get("lr_policy", "step"), epoch, cfg.epochs,
                                   i, n, int(cfg.extra.get("lr_warmup", 0))))
            if i % cfg.log_interval != 0:
                return
            mem_a, mem_c, _ = logfmt.memory_gb(device if device.type == "cuda" else None)
            if topo.is_last and stats is not None:
                s = stats.tolist()
                now = time.time()
                meters["count"] += 1
                meters["loss"] += s[0]
                meters["top1"] += 100.0 * s[1] / cfg.batch_size
                meters["top5"] += 100.0 * s[2] / cfg.batch_size
                c = meters["count"]
                hr = (now - t_epoch) / 3600.0
                say(logfmt.pd_train(epoch, i, n, now - last_t[0], (now - t_epoch) / (i + 1), hr,
                                    mem_a, mem_c, s[0], meters["loss"] / c,
                                    100.0 * s[1] / cfg.batch_size, meters["top1"] / c,
                                    100.0 * s[2] / cfg.batch_size, meters["top5"] / c))
                last_t[0] = now
            else:
                say(logfmt.pd_train_nonlast(epoch, i, n, mem_a, mem_c))

        trace_dir = os.environ.get("DDLB_TIMELINE_DIR")
        if trace_dir:
            rt.forward_stats.reset_stats()
            rt.backward_stats.reset_stats()
            rt.forward_stats.spans, rt.backward_stats.spans = [], []
            rt.forward_stats.keep_spans = rt.backward_stats.keep_spans = True
        if cfg.extra.get("forward_only") == "1":
            # --forward_only (main_with_runtime.py:87-88): the timed region is a forward-only pass
            rt.run_eval(loader() if topo.is_first else None, n)
        else:
            rt.run_epoch(loader() if topo.is_first else None, n, on_forward=on_fwd)
        ms = timer.stop()
        if trace_dir:
            # one Chrome-trace file per rank: load them together to see the 1F1B interleaving
            import json
            from ..parallel.pipedream.stats import pipeline_trace
            os.makedirs(trace_dir, exist_ok=True)
            with open(os.path.join(trace_dir, "pipeline_e%d_rank%d.json" % (epoch, topo.rank)), "w") as f:
                json.dump(pipeline_trace(rt.forward_stats, rt.backward_stats, topo.rank, topo.stage), f)
            rt.forward_stats.keep_spans = rt.backward_stats.keep_spans = False
        sec = time.time() - t_epoch
        say(logfmt.pd_epoch_time(epoch, sec))
        ms_max = max_over_ranks(ms, device)
        images = n_first * pcfg.replicas(0) * cfg.batch_size
        dev_tputs.append(images / (ms_max / 1e3))
        say(logfmt.device_timed(dev_tputs[-1], ms_max / max(n_first, 1), world))
        epoch_times.append(sec)
        # validation (forward-only pipeline, no ack tokens)
        nval_first = cfg.eval_steps if cfg.eval_steps is not None else max(1, loader_size // 10)
        nv = pcfg.num_iterations(topo.stage, nval_first)
        acc = {"n": 0, "l": 0.0, "p1": 0.0, "p5": 0.0}

        def on_eval(i, stats):
            if stats is not None:
                s = stats.tolist()
                acc["n"] += 1
                acc["l"] += s[0]
                acc["p1"] += 100.0 * s[1] / cfg.batch_size
                acc["p5"] += 100.0 * s[2] / cfg.batch_size

        def vloader():
            for i in range(nv):
                yield _float_batch(val_src, i, device, dtype)

        rt.run_eval(vloader() if topo.is_first else None, nv, on_forward=on_eval)
        if topo.is_last and acc["n"]:
            say(logfmt.pd_prec(acc["p1"] / acc["n"], acc["p5"] / acc["n"]))
        ckpt_dir = cfg.extra.get("checkpoint_dir")
        if ckpt_dir:
            from ..utils.checkpoint import save_stage_checkpoint
            # collective over the stage's replicas (sharded optimizer state is gathered); replica 0 writes
            save_stage_checkpoint(ckpt_dir, topo.stage, epoch + 1, cfg.arch, rt,
                                  write=topo.rank_in_stage == 0)
    k = max(len(epoch_times), 1)
    images = n_first * pcfg.replicas(0) * cfg.batch_size
    return {"throughput": images / (sum(epoch_times) / k), "sec_per_epoch": sum(epoch_times) / k,
            "device_throughput": sum(dev_tputs) / k, "device_throughputs": list(dev_tputs),
            "peak_mem_gb": logfmt.memory_gb(device if device.type == "cuda" else None)[0]}
