"""Reference arm of bench.py: runs the UNMODIFIED reference harness scripts from
``baseline/_ref`` (a copy of /root/reference made by ``baseline/install_ref.py``; the reference
has no setup.py/pyproject so ``pip install`` cannot install it) through their own entry points:

  N = 1 : ``benchmark/imagenet/imagenet_pytorch.py -a resnet50 -s 0``
  N > 1 : ``benchmark/imagenet/imagenet_horovod.py -a resnet50 -s 0`` — one process per GPU

with the reference's environment contract (DATADIR/EPOCHS/BATCH_SIZE=32/LOGINTER/CORES_GPU,
``run/run/run_template.sh:70-73,196,223``) on a synthetic JPEG ImageFolder tree of the layout
``benchmark/generate_synthetic_data.py`` writes.  The metric is the reference's own:
``samples/sec = images / wall seconds of the training loop of one epoch`` (DataLoader JPEG
decode + augmentation + synchronous H2D inside the timed region), last epoch of EPOCHS=2 (the
first epoch is the warm-up / cuDNN autotune).  For Horovod the script reports per-rank numbers;
the aggregate is their sum.

Horovod itself is not installable offline (no wheel, no MPI): for N > 1 the script imports
``horovod.torch`` from ``baseline/hvd_shim`` — a stand-alone torch.distributed(NCCL)
implementation of the six calls the script makes, none of it from ddlbench_b200 — and is
launched with torchrun instead of horovodrun.  This is declared in the JSON line.
"""
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "_ref")
BATCH = 32


def _gen_tree(datadir, n_train, n_val, classes=8):
    from jpeg_tree import make_tree
    make_tree(datadir, n_train, n_val, classes=classes)


MIN_ITERS = 200      # iterations per epoch: long enough that DataLoader worker start-up is noise


def _cores(gpus):
    # the reference's run.sh gives 12 or 24 CPU cores per GPU depending on the partition
    # (run/run/run.sh:65-69); use what this box has, capped at the larger of the two
    return max(2, min(24, (os.cpu_count() or 8) // max(gpus, 1)))


def _parse(out):
    per_epoch = {}
    for m in re.finditer(r"(\d+)/(\d+) epoch \| train loss:\S+ ([\d.]+) samples/sec", out):
        per_epoch.setdefault(int(m.group(1)), []).append(float(m.group(3)))
    return per_epoch


def _result(metric, gpus, steps, value, per_epoch, script, cores, gen_s, launcher):
    # the reference's own number IS end to end: host wall clock around DataLoader decode +
    # augmentation + synchronous H2D of fp32 NCHW tensors + step + loss.item()
    h2d = BATCH * 3 * 224 * 224 * 4 + BATCH * 8
    return {
        "metric": metric, "value": round(value, 2), "unit": "images/sec", "n_gpus": gpus,
        "steps": steps, "warmup": steps, "ms_per_step": round(BATCH * gpus / value * 1e3, 4),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32",
        "data": "synthetic JPEG ImageFolder tree (reference layout, baseline/jpeg_tree.py), decoded "
                "by the reference's DataLoader",
        "impl": "reference",
        "config": {"model": "resnet50", "global_batch": BATCH * gpus, "per_gpu_batch": BATCH,
                   "parallelism": "single" if gpus == 1 else "dp%d (horovod API)" % gpus,
                   "script": os.path.relpath(script, HERE),
                   "timing": "the reference's own: host wall clock of the epoch's training "
                             "loop (includes DataLoader JPEG decode + H2D), epoch 2 of 2; "
                             "`steps` = iterations of that epoch, `warmup` = iterations of the "
                             "discarded first epoch",
                   "cores_gpu": cores,
                   "horovod": None if gpus == 1 else
                   "baseline/hvd_shim (torch.distributed NCCL stand-in; Horovod is not "
                   "installable offline), launched with %s instead of horovodrun" % launcher},
        "e2e": {"value": round(value, 2), "unit": "images/sec", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": 4,
                "note": "identical to `value`: the reference's metric is its end-to-end loop"},
        "gpu_launches": None,
        "epoch_throughputs": {str(k): v for k, v in sorted(per_epoch.items())},
        "datagen_seconds": round(gen_s, 1),
    }


def _run_inprocess(script, gpus, steps, metric):
    """bench.py was itself launched by torchrun with one process per GPU: run the reference
    script inside each of those processes (the shim's hvd.init() joins the torchrun world)."""
    import io
    import runpy
    rank = int(os.environ["RANK"])
    tmp = os.path.join(tempfile.gettempdir(), "ddlb_ref_%s" % os.environ.get("MASTER_PORT", "0"))
    ready = os.path.join(tmp, ".ready")
    t0 = time.time()
    if rank == 0:
        shutil.rmtree(tmp, ignore_errors=True)
        os.makedirs(tmp, exist_ok=True)
        _gen_tree(tmp, steps * BATCH * gpus, BATCH * gpus)
        open(ready, "w").write("ok")
    else:
        while not os.path.exists(ready):
            time.sleep(0.2)
    gen_s = time.time() - t0
    cores = _cores(gpus)
    os.environ.update(DATADIR=tmp, EPOCHS="2", BATCH_SIZE=str(BATCH),
                      LOGINTER=str(max(steps // 2, 1)), CORES_GPU=str(cores))
    sys.path.insert(0, os.path.join(HERE, "hvd_shim"))
    sys.path.insert(0, os.path.dirname(script))
    argv, sys.argv = sys.argv, [script, "-a", "resnet50", "-s", "0"]
    buf = io.StringIO()
    real = sys.stdout

    class Tee:
        def write(self, s):
            buf.write(s)
            return real.write(s) if rank == 0 and os.environ.get("REF_VERBOSE") else len(s)

        def flush(self):
            real.flush()

    sys.stdout = Tee()
    try:
        cwd = os.getcwd()
        os.chdir(os.path.dirname(script))
        runpy.run_path(script, run_name="__main__")
    finally:
        sys.stdout = real
        sys.argv = argv
        os.chdir(cwd)
    import torch.distributed as dist
    per_epoch = _parse(buf.getvalue())
    mine = per_epoch[max(per_epoch)][0] if per_epoch else 0.0
    vals = [None] * gpus
    dist.all_gather_object(vals, (mine, {k: v[0] for k, v in per_epoch.items()}))
    dist.barrier()
    if rank == 0:
        shutil.rmtree(tmp, ignore_errors=True)
    dist.destroy_process_group()
    if rank != 0:
        return {}
    merged = {}
    for _, pe in vals:
        for k, v in pe.items():
            merged.setdefault(k, []).append(v)
    return _result(metric, gpus, steps, sum(v for v, _ in vals), merged, script, cores, gen_s,
                   "torchrun (in the bench's own ranks)")


def main(gpus, steps, warmup, metric):
    script = os.path.join(REF, "benchmark", "imagenet",
                          "imagenet_pytorch.py" if gpus == 1 else "imagenet_horovod.py")
    if not os.path.isfile(script):
        return {"impl": "reference", "unavailable": "baseline/_ref is missing (run python baseline/install_ref.py)"}
    try:
        import torch
        import torchvision  # noqa: F401
        if not torch.cuda.is_available() or torch.cuda.device_count() < gpus:
            return {"impl": "reference", "unavailable": "needs %d CUDA device(s)" % gpus}
    except Exception as e:  # noqa: BLE001
        return {"impl": "reference", "unavailable": "torch/torchvision import failed: %s" % e}
    steps = max(steps, warmup, MIN_ITERS)
    sys.path.insert(0, HERE)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and "RANK" in os.environ:
        try:
            return _run_inprocess(script, int(os.environ["WORLD_SIZE"]), steps, metric)
        except BaseException as e:  # noqa: BLE001 - the arm must print a line and exit 0
            if int(os.environ.get("RANK", "0")) != 0:
                return {}
            return {"impl": "reference",
                    "unavailable": "reference horovod script failed: %s: %s" % (
                        type(e).__name__, str(e)[:300].replace("\n", " | "))}
    tmp = tempfile.mkdtemp(prefix="ddlb_ref_")
    try:
        t0 = time.time()
        _gen_tree(tmp, steps * BATCH * gpus, BATCH * gpus)
        gen_s = time.time() - t0
        cores = _cores(gpus)
        env = dict(os.environ, DATADIR=tmp, EPOCHS="2", BATCH_SIZE=str(BATCH),
                   LOGINTER=str(max(steps // 2, 1)), CORES_GPU=str(cores))
        if gpus == 1:
            cmd = [sys.executable, script, "-a", "resnet50", "-s", "0"]
        else:
            env["PYTHONPATH"] = os.path.join(HERE, "hvd_shim") + os.pathsep + env.get("PYTHONPATH", "")
            port = 29000 + (os.getpid() % 500)
            cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                   "--nproc-per-node=%d" % gpus, "--master-addr", "127.0.0.1",
                   "--master-port", str(port), script, "-a", "resnet50", "-s", "0"]
        p = subprocess.run(cmd, env=env, cwd=os.path.dirname(script), capture_output=True,
                           text=True, timeout=3000)
        out = p.stdout
        if p.returncode != 0:
            return {"impl": "reference",
                    "unavailable": "reference script failed (rc=%d): %s" % (
                        p.returncode, (p.stderr or out)[-400:].replace("\n", " | "))}
        per_epoch = _parse(out)
        last = max(per_epoch)
        vals = per_epoch[last]
        value = sum(vals) if gpus > 1 else vals[0]
        if gpus > 1 and len(vals) != gpus:
            value = sum(vals) / len(vals) * gpus
        return _result(metric, gpus, steps, value, per_epoch, script, cores, gen_s, "torchrun")
    except subprocess.TimeoutExpired:
        return {"impl": "reference", "unavailable": "reference script timed out"}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    print(json.dumps(main(int(sys.argv[1]) if len(sys.argv) > 1 else 1, 20, 5, "images/sec")))
