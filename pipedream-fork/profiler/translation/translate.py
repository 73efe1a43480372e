"""GNMT translate tool at the reference's path (profiler/translation/translate.py); see
ddlbench_b200/harness/translate.py."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "..")))
from ddlbench_b200.harness.translate import main  # noqa: E402

if __name__ == "__main__":
    main()
