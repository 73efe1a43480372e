"""One CTA-pair GEMM launch (8192^3) for an ncu capture: python tools/ncu_gemm_pair.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ddlbench_b200 import ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
a = (torch.randn(n, n, device="cuda")).to(torch.bfloat16)
w = (torch.randn(n, n, device="cuda") * n ** -0.5).to(torch.bfloat16)
out = torch.empty(n, n, device="cuda", dtype=torch.bfloat16)
for _ in range(2):
    ops.gemm_pair(a, w, out=out)
torch.cuda.synchronize()
torch.cuda.profiler.start()
ops.gemm_pair(a, w, out=out)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
