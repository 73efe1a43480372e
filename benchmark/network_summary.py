"""Layer summaries of every model x data-set shape (reference: benchmark/network_summary.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(os.path.dirname(__file__))))
from ddlbench_b200.utils.summary import network_summary  # noqa: E402

if __name__ == "__main__":
    network_summary()
