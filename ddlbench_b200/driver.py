"""YAML-driven experiment launcher — the role of ``pipedream-fork/runtime/driver.py``.

Same configuration keys (``driver.py:13-40``): required ``log_directory, module, data_dir,
container, machines, model_type, distributed_backend``; optional ``batch_size, learning_rate,
learning_rate_policy, weight_decay, epochs, config_file, print_frequency, no_input_pipelining,
verbose_frequency, lr_warmup, synthetic_data, recompute, macrobatch``.  ``machines`` is the
reference's ``['host:gpu', ...]`` list: its length is the number of workers.

Differences by design: workers are local processes started through ``torch.distributed.run``
(one per entry of ``machines``; the host part is recorded but all GPUs of an NVSwitch box are
local), there is no docker/ssh layer, and ``module`` (``models.<arch>.gpus=<N>``) selects the
architecture — stage modules are sliced from the layer graph instead of being imported.
The launcher writes ``<log_directory>/<timestamp>/{config.yml, machinefile, output.log}`` and
records the git commit like the reference (``driver.py:73-100``).
"""
from __future__ import annotations

import argparse
import datetime
import os
import re
import subprocess
import sys
from typing import Dict, List

import yaml

REQUIRED = ("log_directory", "module", "data_dir", "container", "machines", "model_type",
            "distributed_backend")
OPTIONAL = ("batch_size", "learning_rate", "learning_rate_policy", "weight_decay", "epochs",
            "config_file", "print_frequency", "no_input_pipelining", "verbose_frequency",
            "lr_warmup", "synthetic_data", "recompute", "macrobatch")


def load(path: str) -> Dict:
    with open(path) as f:
        conf = yaml.safe_load(f)
    missing = [k for k in REQUIRED if k not in conf]
    if missing:
        raise ValueError("missing configuration fields: %s" % missing)
    return conf


def arch_of(module: str) -> str:
    m = re.match(r"models\.([A-Za-z0-9_]+)\.gpus=(\d+)", module)
    return m.group(1) if m else module.split(".")[-1]


def build_command(conf: Dict, resume: str = "", dataset: str = "imagenet", port: int = 29450) -> List[str]:
    n = len(conf["machines"])
    arch = arch_of(conf["module"])
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % n,
           "--master-addr", "127.0.0.1", "--master-port", str(port), "-m", "ddlbench_b200.cli",
           "-b", dataset, "-f", "pipedream", "-g", str(n), "-m", arch, "--worker",
           "-p", str(conf.get("print_frequency", 100))]
    if conf.get("batch_size"):
        cmd += ["--batch-size", str(conf["batch_size"])]
    if conf.get("epochs"):
        cmd += ["--epochs", str(conf["epochs"])]
    if conf.get("config_file"):
        cmd += ["--config-path", str(conf["config_file"])]
    if conf.get("no_input_pipelining"):
        cmd.append("--no-input-pipelining")
    if conf.get("macrobatch"):
        cmd.append("--macrobatch")
    if conf.get("distributed_backend") == "gloo":
        cmd += ["--device", "cpu"]
    if resume:
        cmd += ["--checkpoint-dir", resume]
    return cmd


def main(argv=None) -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--config_file", required=True, help="Path to configuration file")
    ap.add_argument("--resume", default="", help="checkpoint directory to resume from")
    ap.add_argument("--dataset", default="imagenet")
    ap.add_argument("--dry_run", action="store_true", help="print the command, do not launch")
    ap.add_argument("--quiet", action="store_true")
    a = ap.parse_args(argv)
    conf = load(a.config_file)
    out_dir = os.path.join(conf["log_directory"],
                           datetime.datetime.now().isoformat().split(".")[0].replace(":", "-"))
    cmd = build_command(conf, a.resume, a.dataset)
    if a.dry_run:
        print(" ".join(cmd))
        return 0
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, "config.yml"), "w") as f:
        yaml.safe_dump(conf, f)
    with open(os.path.join(out_dir, "machinefile"), "w") as f:
        f.write("\n".join(str(m) for m in conf["machines"]) + "\n")
    try:
        commit = subprocess.run(["git", "rev-parse", "HEAD"], capture_output=True, text=True).stdout.strip()
    except Exception:  # noqa: BLE001
        commit = "unknown"
    env = dict(os.environ)
    for k, name in (("learning_rate", "DDLB_LR"), ("weight_decay", "DDLB_WD")):
        if conf.get(k) is not None:
            env[name] = str(conf[k])
    with open(os.path.join(out_dir, "output.log"), "w") as log:
        log.write("git commit: %s\ncommand: %s\n" % (commit, " ".join(cmd)))
        log.flush()
        p = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
        for line in p.stdout:
            log.write(line)
            if not a.quiet:
                sys.stdout.write(line)
        p.wait()
    print("output in %s" % out_dir)
    return p.returncode


if __name__ == "__main__":
    raise SystemExit(main())
