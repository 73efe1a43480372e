"""Shortest possible hardware check of the end-to-end JPEG loop with decode worker processes
(bench.py::_e2e_jpeg): python tools/e2e_procs_check.py [steps]"""
import os, sys, time, types
t00 = time.time()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from ddlbench_b200.harness import Trainer
from ddlbench_b200.harness.loop import timed_loop, max_over_ranks
from ddlbench_b200.data.pipeline import DevicePrefetcher
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
dev = torch.device("cuda", 0)
tr = Trainer(bench.ARCH, bench.DATASET, bench.PER_GPU_BATCH, device=str(dev), dtype=torch.bfloat16, cuda_graph=True, tcgen05=True)
args = types.SimpleNamespace(steps=steps, warmup=5)
print("setup %.1fs" % (time.time() - t00), flush=True)
for mode in (("procs", None), ("threads", "0")):
    if mode[1] is not None:
        os.environ["BENCH_DECODE_PROCS"] = mode[1]
    t0 = time.time()
    out = bench._e2e_jpeg(args, tr, dev, 0, 1, None, timed_loop, max_over_ranks, DevicePrefetcher)
    print(mode[0], {k: out.get(k) for k in ("value", "ms_per_step", "decode_procs", "decode_procs_error", "host_cores", "decode_threads", "error")},
          "%.1fs" % (time.time() - t0), flush=True)
