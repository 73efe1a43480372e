"""Inference partitioner at the reference's path; see ddlbench_b200/partition/inference_optimizer.py."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")))
from ddlbench_b200.partition.inference_optimizer import cli  # noqa: E402

if __name__ == "__main__":
    raise SystemExit(cli())
