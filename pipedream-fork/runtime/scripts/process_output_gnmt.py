"""GNMT log post-processing at the reference's path; see ddlbench_b200/utils/process_output.py."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "..")))
from ddlbench_b200.utils.process_output import main  # noqa: E402

if __name__ == "__main__":
    raise SystemExit(main(["--gnmt"] + sys.argv[1:]))
