"""Profiler entry point with the reference's path and flags (pipedream-fork/profiler/
image_classification/main.py: ``-a <arch> -b <batch> --data_dir --workers --epochs --print-freq
--synthetic_data --highres --profile_directory``).  ``-a`` accepts the reference's names:
``mnist_<arch>``, ``cifar10_<arch>``, ``highres_<arch>`` or a plain (ImageNet) architecture.
Writes ``<profile_directory>/<arch>/graph.txt`` through ddlbench_b200.partition.profiler."""
import argparse
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "..")))
from ddlbench_b200.partition.profiler import profile_model  # noqa: E402


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description="PipeDream profiler (ddlbench_b200)")
    ap.add_argument("-a", "--arch", default="resnet18")
    ap.add_argument("-b", "--batch-size", type=int, default=256)
    ap.add_argument("--profile_directory", default="profiles/")
    ap.add_argument("--highres", type=int, default=0)
    for flag in ("--data_dir", "-j", "--workers", "--epochs", "--print-freq", "-p", "-s",
                 "--synthetic_data", "--num_minibatches", "--lr", "--momentum", "--weight-decay"):
        ap.add_argument(flag, default=None)            # accepted, not needed by this profiler
    ap.add_argument("--forward_only", action="store_true")
    ap.add_argument("-v", "--verbose", action="store_true")
    a = ap.parse_args(argv)
    name, dataset = a.arch, ("highres" if a.highres else "imagenet")
    for prefix in ("mnist", "cifar10", "highres"):
        if name.startswith(prefix + "_"):
            name, dataset = name[len(prefix) + 1:], prefix
    profile_model(name, dataset, a.batch_size, out_dir=os.path.join(a.profile_directory, a.arch),
                  verbose=True)
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
