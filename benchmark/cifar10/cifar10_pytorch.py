"""cifar10 / pytorch: same command line and environment contract as the reference script of this
name; the work is done by ddlbench_b200 (see ddlbench_b200/scripts.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from ddlbench_b200.scripts import main  # noqa: E402

if __name__ == "__main__":
    main("cifar10", "pytorch")
