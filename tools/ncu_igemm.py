"""A few representative igemm launches for an ncu capture (python tools/ncu_igemm.py [names...])."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from ddlbench_b200 import ops
import igemm_check as C
want = sys.argv[1:] or ["l1.conv3", "l4.b1.conv2", "l3.conv3"]
cfgs = [c for c in C.RESNET50 if c[0] in want]
for name, n, ci, co, hw, k, s, p in cfgs:
    x = C._rand((n, ci, hw, hw)).contiguous(memory_format=C.CL)
    w = C._rand((co, ci, k, k), 0.05).contiguous(memory_format=C.CL)
    sums = torch.zeros(ops.BN_REPLICAS * 2 * co, device="cuda")
    stat = torch.zeros(2 * co, device="cuda")
    tk = torch.zeros((co + 63) // 64, dtype=torch.int32, device="cuda")
    fin = (tk, stat, None, None, 1e-5, 0.1)
    for _ in range(3):
        y = ops.conv_fprop(x, w, s, [p, p], col_sums=sums, bn_finalize=fin)
    dy = torch.randn_like(y)
    gw = torch.zeros((co, k, k, ci), device="cuda", dtype=torch.bfloat16).permute(0, 3, 1, 2)
    for _ in range(3):
        dx = ops.conv_dgrad(dy, w, x.shape, s, [p, p])
        ops.conv_wgrad_into(dy, x, gw, s, [p, p], False)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    y = ops.conv_fprop(x, w, s, [p, p], col_sums=sums, bn_finalize=fin)
    dx = ops.conv_dgrad(dy, w, x.shape, s, [p, p])
    ops.conv_wgrad_into(dy, x, gw, s, [p, p], False)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("done", name)
