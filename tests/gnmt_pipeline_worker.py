"""torchrun worker: 1F1B GNMT pipeline on CPU/gloo; last stage writes its loss curve."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ddlbench_b200.harness import translation  # noqa: E402

out_dir = sys.argv[1]
stages = int(os.environ["WORLD_SIZE"])
res = translation.main(["--data-dir", os.path.join(out_dir, "data"), "--stages", str(stages), "--hidden", "32",
                        "--num-layers", "2", "--batch-size", "16", "--max-len", "16", "--steps", "30",
                        "--epochs", "2", "--lr", "0.01", "--device", "cpu", "--dropout", "0.0",
                        "--synthetic-sentences", "600", "--synthetic-vocab", "40", "--print-freq", "5",
                        "--checkpoint-dir", os.path.join(out_dir, "ckpt")])
if int(os.environ["RANK"]) == stages - 1:
    json.dump(res, open(os.path.join(out_dir, "result.json"), "w"))
