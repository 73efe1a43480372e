"""Per-CTA stage timestamps of single weight-gradient launches (see WG_STAMP in gemm_wgrad_tcgen05.cu).
python tools/wgrad_trace.py [layer names...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from ddlbench_b200 import ops
import igemm_check as C
want = sys.argv[1:] or ["l4.b1.conv2", "l4.b1.conv1", "l3.b1.conv1", "l3.b1.conv2"]
for name, n, ci, co, hw, k, s, p in [c for c in C.RESNET50 if c[0] in want]:
    x = C._rand((n, ci, hw, hw)).contiguous(memory_format=C.CL)
    oh = (hw + 2 * p - k) // s + 1
    dy = C._rand((n, co, oh, oh)).contiguous(memory_format=C.CL)
    gw = torch.zeros((co, k, k, ci), device="cuda", dtype=torch.bfloat16).permute(0, 3, 1, 2)
    for _ in range(3):
        ops.conv_wgrad_into(dy, x, gw, s, [p, p], False)
    buf = torch.zeros(8192 * 16, dtype=torch.int64, device="cuda")
    ops.ext().wgrad_set_trace(buf)
    torch.cuda.synchronize()
    ops.conv_wgrad_into(dy, x, gw, s, [p, p], False)
    torch.cuda.synchronize()
    ops.ext().wgrad_set_trace(None)
    t = buf.view(-1, 16).cpu()
    t = t[t[:, 0] > 0]
    t0 = int(t[:, 0].min())
    print("%s wgrad: %d CTAs" % (name, t.shape[0]))
    names = ["entry", "prologue", "acc ready", "epilogue done"] + ["kb %d issued" % (4 * i) for i in range(9)] + \
        ["tmem->smem done", "after bar", "acc ready (warp 3)"]
    for kk in range(16):
        col = t[:, kk]
        m = col > 0
        if not m.any():
            continue
        rel = (col[m] - t0).float() / 1e3
        print("   %-16s n=%4d first %6.2f  mean %6.2f  last %6.2f us" % (names[kk], int(m.sum()), rel.min(), rel.mean(), rel.max()))
