"""YAML-driven experiment launcher — the role of ``pipedream-fork/runtime/driver.py``.

Same configuration keys (``driver.py:13-40``): required ``log_directory, module, data_dir,
container, machines, model_type, distributed_backend``; optional ``batch_size, learning_rate,
learning_rate_policy, weight_decay, epochs, config_file, print_frequency, no_input_pipelining,
verbose_frequency, lr_warmup, synthetic_data, recompute, macrobatch``.  ``machines`` is the
reference's ``['host:gpu', ...]`` list: its length is the number of workers.

Differences by design: when every entry of ``machines`` names the same host (the normal case:
all GPUs of an NVSwitch box) the workers are local processes started through
``torch.distributed.run``; with several hosts one ``ddlbench_b200.launch`` per host is started
over ``ssh -n`` (the reference's fan-out, ``driver.py:313-356``), all pointing at the first host as
rendezvous.  ``--docker`` runs every node's launcher inside the configuration's ``container``
(one container per node, ``docker run --gpus all --net=host --ipc=host --shm-size 16g`` with the
repository, the data directory, the log directory and ``--mount_directories`` mounted — the
reference's ``--launch_single_container`` mode, ``driver.py:276-318``, with ``nvidia-docker``
replaced by ``--gpus all``).  ``module`` (``models.<arch>.gpus=<N>``) selects the architecture —
stage modules are sliced from the layer graph instead of being imported; ``model_type:
translation`` starts the GNMT harness.
The launcher writes ``<log_directory>/<timestamp>/{config.yml, machinefile, output.log,
command_history.txt, git_info.log}`` (commit + working-tree diff) like the reference
(``driver.py:170-200``).
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
    if conf.get("stages"):                      # e.g. "1x16": data parallel through the stage runtime
        cmd += ["--stages", str(conf["stages"])]
    if conf.get("weight_decay") is not None:
        cmd += ["--wd", str(conf["weight_decay"])]
    if conf.get("learning_rate") is not None:
        cmd += ["--lr", str(conf["learning_rate"])]
    if conf.get("learning_rate_policy"):
        cmd += ["--lr_policy", str(conf["learning_rate_policy"])]
    if conf.get("no_input_pipelining"):
        cmd.append("--no-input-pipelining")
    if conf.get("macrobatch"):
        cmd.append("--macrobatch")
    if conf.get("distributed_backend") == "gloo":
        cmd += ["--device", "cpu"]
    if resume:
        # continue FROM the checkpoints and keep writing them to the same place (reference:
        # main_with_runtime.py --resume <prefix> + --checkpoint_dir)
        cmd += ["--resume", resume, "--checkpoint-dir", resume]
    return cmd


def hosts_of(conf: Dict) -> "Dict[str, int]":
    """host -> number of workers, in first-appearance order (entries are ``host:gpu``)."""
    out: Dict[str, int] = {}
    for m in conf["machines"]:
        h = str(m).split(":")[0]
        out[h] = out.get(h, 0) + 1
    return out


def worker_args(conf: Dict, resume: str = "", dataset: str = "imagenet") -> List[str]:
    """Module + arguments every worker runs (after the launcher part of the command)."""
    n = len(conf["machines"])
    if conf.get("model_type") == "translation":
        args = ["ddlbench_b200.harness.translation", "--stages", str(n)]
        if "gnmt_large" in conf["module"]:
            args += ["--num-layers", "8"]
        for key, flag in (("batch_size", "--batch-size"), ("epochs", "--epochs"),
                          ("learning_rate", "--lr"), ("print_frequency", "--print-freq")):
            if conf.get(key):
                args += [flag, str(conf[key])]
        if conf.get("data_dir") and conf["data_dir"] != "synthetic":
            args += ["--data-dir", str(conf["data_dir"])]
        if resume:
            args += ["--resume", resume, "--checkpoint-dir", resume]
        return args
    return build_command(conf, resume, dataset)[build_command(conf, resume, dataset).index("ddlbench_b200.cli"):]


def docker_prefix(conf: Dict, repo: str, mounts=()) -> List[str]:
    """``docker run`` argv (up to and including ``/bin/bash -c``) for one node's container."""
    dirs = [repo]
    for d in [conf.get("data_dir"), conf.get("log_directory")] + list(mounts or ()):
        if d and d != "synthetic" and os.path.isabs(str(d)) and str(d) not in dirs:
            dirs.append(str(d))
    cmd = ["docker", "run", "--rm", "--gpus", "all", "--net=host", "--ipc=host", "--shm-size", "16g"]
    for d in dirs:
        cmd += ["-v", "%s:%s" % (d, d)]
    return cmd + ["-w", repo, str(conf["container"]), "/bin/bash", "-c"]


def build_node_commands(conf: Dict, resume: str = "", dataset: str = "imagenet",
                        port: int = 29450, repo: str = "", docker: bool = False,
                        mounts=()) -> "List[tuple]":
    """[(host, argv)] — one entry for a single-host run, one ``ssh`` command per host otherwise;
    ``docker``: each node's launcher runs inside ``conf['container']``."""
    hosts = hosts_of(conf)
    wargs = worker_args(conf, resume, dataset)
    names = list(hosts)
    repo = repo or os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if docker and len(names) == 1:
        inner = ("python -m torch.distributed.run --nnodes=1 --nproc-per-node=%d --master-addr 127.0.0.1 "
                 "--master-port %d -m %s" % (hosts[names[0]], port, " ".join(wargs)))
        return [(names[0], docker_prefix(conf, repo, mounts) + [inner])]
    if len(names) == 1:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
               "--nproc-per-node=%d" % hosts[names[0]], "--master-addr", "127.0.0.1",
               "--master-port", str(port), "-m"] + wargs
        return [(names[0], cmd)]
    per = set(hosts.values())
    if len(per) != 1:
        raise ValueError("every host must run the same number of workers, got %s" % hosts)
    if "-g" in wargs:                       # cli workers: GPUs per node + number of nodes
        wargs[wargs.index("-g") + 1] = str(hosts[names[0]])
        wargs += ["-n", str(len(names))]
    out = []
    for i, h in enumerate(names):
        inner = ("cd %s && python -m ddlbench_b200.launch --nnodes %d --node_rank %d "
                 "--nproc_per_node %d --master_addr %s --master_port %d --no-rank-args "
                 "--pid-file driver_node%d.pids -m %s" % (
                     repo, len(names), i, hosts[h], names[0], port, i, " ".join(wargs)))
        if docker:
            import shlex
            inner = " ".join(shlex.quote(c) for c in docker_prefix(conf, repo, mounts) + [inner])
        out.append((h, ["ssh", "-n", h, "-o", "StrictHostKeyChecking=no", inner]))
    return out


def main(argv=None) -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--config_file", required=True, help="Path to configuration file")
    ap.add_argument("--resume", default="", help="checkpoint directory to resume from")
    ap.add_argument("--dataset", default="imagenet")
    ap.add_argument("--dry_run", action="store_true", help="print the command, do not launch")
    ap.add_argument("--quiet", action="store_true")
    ap.add_argument("--docker", "--launch_single_container", action="store_true", dest="docker",
                    help="run each node's launcher inside the configuration's container")
    ap.add_argument("--mount_directories", nargs="*", default=[], help="extra directories to mount (docker)")
    a = ap.parse_args(argv)
    conf = load(a.config_file)
    out_dir = os.path.join(conf["log_directory"],
                           datetime.datetime.now().isoformat().split(".")[0].replace(":", "-"))
    nodes = build_node_commands(conf, a.resume, a.dataset, docker=a.docker, mounts=a.mount_directories)
    cmd = nodes[0][1]
    if a.dry_run:
        for _, c in nodes:
            print(" ".join(c))
        return 0
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, "config.yml"), "w") as f:
        yaml.safe_dump(conf, f)
    with open(os.path.join(out_dir, "machinefile"), "w") as f:
        f.write("\n".join(str(m) for m in conf["machines"]) + "\n")
    with open(os.path.join(out_dir, "command_history.txt"), "w") as f:
        for _, c in nodes:
            f.write(" ".join(c) + "\n")
    try:
        commit = subprocess.run(["git", "rev-parse", "HEAD"], capture_output=True, text=True).stdout.strip()
        diff = subprocess.run(["git", "diff"], capture_output=True, text=True).stdout
    except Exception:  # noqa: BLE001
        commit, diff = "unknown", ""
    with open(os.path.join(out_dir, "git_info.log"), "w") as f:
        f.write(commit + "\n" + diff)
    env = dict(os.environ)
    for k, name in (("learning_rate", "DDLB_LR"), ("weight_decay", "DDLB_WD")):
        if conf.get(k) is not None:
            env[name] = str(conf[k])
    with open(os.path.join(out_dir, "output.log"), "w") as log:
        log.write("git commit: %s\ncommand: %s\n" % (commit, " ".join(cmd)))
        log.flush()
        others = [subprocess.Popen(c, stdout=open(os.path.join(out_dir, "output_%s.log" % h), "w"),
                                   stderr=subprocess.STDOUT, env=env) for h, c in nodes[1:]]
        p = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
        for line in p.stdout:
            log.write(line)
            if not a.quiet:
                sys.stdout.write(line)
        p.wait()
        for o in others:
            o.wait()
    print("output in %s" % out_dir)
    return p.returncode


if __name__ == "__main__":
    raise SystemExit(main())
