"""GNMT trainer + layer profiler at the reference's path (profiler/translation/train.py): trains,
then writes the profile graph; with --gpus/--out it profiles, partitions and writes conf files.
See ddlbench_b200/harness/translation_profiler.py."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "..")))
from ddlbench_b200.harness.translation_profiler import main  # noqa: E402

if __name__ == "__main__":
    raise SystemExit(main())
