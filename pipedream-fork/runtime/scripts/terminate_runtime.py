"""Stop the processes of a run at the reference's path: takes the PID file written by launch.py --pid-file (the reference pkill -f's every python on the machine list; exact PIDs here)."""
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "..")))
from ddlbench_b200.launch import main  # noqa: E402

if __name__ == "__main__":
    raise SystemExit(main(["--terminate"] + sys.argv[1:]))
