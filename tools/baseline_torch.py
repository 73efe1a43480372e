"""Library baseline for context: the same ResNet-50 step with stock PyTorch ops (cuDNN/cuBLAS,
   torch.optim.SGD), bf16 channels_last, eager and whole-step CUDA graph."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from ddlbench_b200.models import build_model
torch.backends.cudnn.benchmark = True
bs = int(os.environ.get("BS", "32"))
def timeit(fn, n=30, w=5):
    for _ in range(w): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
res = {}
for mode in ("bf16_cl", "fp32_tf32"):
    m = build_model("resnet50", "imagenet").cuda()
    if mode == "bf16_cl":
        m = m.to(torch.bfloat16).to(memory_format=torch.channels_last)
        x = torch.randn(bs, 3, 224, 224, device="cuda", dtype=torch.bfloat16).contiguous(memory_format=torch.channels_last)
    else:
        x = torch.randn(bs, 3, 224, 224, device="cuda")
    t = torch.randint(0, 1000, (bs,), device="cuda")
    opt = torch.optim.SGD(m.parameters(), lr=0.1, momentum=0.9, weight_decay=1e-4)
    def step():
        opt.zero_grad(set_to_none=True)
        loss = F.cross_entropy(m(x).float(), t)
        loss.backward()
        opt.step()
        return loss
    res[mode + "_eager_ms"] = timeit(step)
    try:
        s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3): step()
        torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        opt.zero_grad(set_to_none=True)
        with torch.cuda.graph(g):
            loss = F.cross_entropy(m(x).float(), t); loss.backward(); opt.step()
        res[mode + "_graph_ms"] = timeit(g.replay)
    except Exception as e:
        res[mode + "_graph_ms"] = "failed: %s" % str(e)[:200]
for k, v in list(res.items()):
    if isinstance(v, float): res[k.replace("_ms", "_img_s")] = bs / v * 1e3
print(json.dumps(res))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/baseline_torch_bs%d.json" % bs, "w"), indent=1)
