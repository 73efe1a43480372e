"""Code-generator entry point at the reference's path; same flags (see
ddlbench_b200/partition/codegen.py)."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")))
from ddlbench_b200.partition.codegen import cli  # noqa: E402

if __name__ == "__main__":
    raise SystemExit(cli())
