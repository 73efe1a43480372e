"""In-graph kernel timeline of the ResNet-50 training step (CUPTI through torch.profiler: real
concurrency and gaps, unlike ncu which serialises).  Writes gpurun_out/timeline_<tag>.json:
per-kernel totals, per-stream busy time and the ordered kernel list of one replay."""
import json, os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from ddlbench_b200.harness import Trainer
from ddlbench_b200.data.synthetic import SyntheticImages
tag = sys.argv[1] if len(sys.argv) > 1 else "default"
bs = 32
tr = Trainer("resnet50", "imagenet", batch_size=bs, device="cuda", dtype=torch.bfloat16, cuda_graph=True, tcgen05=True)
src = SyntheticImages("imagenet", bs, num_batches=2, device="cuda")
x, t = src.batch(0)
for _ in range(6):
    tr.step_u8(x, t, src.mean, src.std)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        tr.step_u8(x, t, src.mean, src.std)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.name != "Memset (Device)" or (e.device_type == torch.autograd.DeviceType.CUDA)]
rows = []
for e in evs:
    tr_ = e.time_range
    rows.append((tr_.start, tr_.end, e.name[:90], getattr(e, "device_index", 0), getattr(e, "stream", None) if hasattr(e, "stream") else None))
rows.sort()
# split into replays by the u8 normalise kernel
starts = [i for i, r in enumerate(rows) if "u8_to_bf16" in r[2]]
if len(starts) >= 2:
    one = rows[starts[-2]:starts[-1]]
else:
    one = rows
t0 = one[0][0]
span = max(r[1] for r in one) - t0
agg = collections.defaultdict(lambda: [0, 0.0])
for s, e, n, *_ in one:
    k = n.split("<")[0].split("(")[0][:60]
    agg[k][0] += 1
    agg[k][1] += e - s
# union busy time (any kernel running) and sum
ivs = sorted((s, e) for s, e, *_ in one)
busy, cur_s, cur_e = 0.0, None, None
for s, e in ivs:
    if cur_e is None or s > cur_e:
        if cur_e is not None:
            busy += cur_e - cur_s
        cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
busy += (cur_e - cur_s) if cur_e is not None else 0
out = {"tag": tag, "span_us": span, "union_busy_us": busy, "sum_kernel_us": sum(e - s for s, e, *_ in one),
       "launches": len(one),
       "by_kernel": sorted(([k, v[0], round(v[1], 1)] for k, v in agg.items()), key=lambda r: -r[2]),
       "timeline": [[round(s - t0, 1), round(e - s, 1), n] for s, e, n, *_ in one]}
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/timeline_%s.json" % tag, "w"))
print("span %.1f us  union busy %.1f  sum of kernels %.1f  launches %d" % (span, busy, out["sum_kernel_us"], len(one)))
for k, c, v in out["by_kernel"][:25]:
    print("%8.1f us x%-4d %s" % (v, c, k))
