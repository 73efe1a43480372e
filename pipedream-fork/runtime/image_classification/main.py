"""The upstream non-pipelined data-parallel baseline at the reference's path with its flags
(``-m models.<arch>.gpus=<N> -b <batch> -s --epochs E --num_minibatches K --forward_only --resume
PATH --fp16 --world-size W --dist-url URL``).  See ``ddlbench_b200/harness/dp_baseline.py``."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "..")))
from ddlbench_b200.harness.dp_baseline import main  # noqa: E402

if __name__ == "__main__":
    main()
