"""Per-CTA stage timestamps (globaltimer) of single igemm launches: where the time of the small-M /
long-K layers goes.  python tools/igemm_trace.py [layer names...]   (env DDLB_IG_SPLIT=cluster|slots|off)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from ddlbench_b200 import ops
import igemm_check as C
NAMES = ["entry", "prologue+pdl", "acc ready", "cl.bar1", "scatter", "cl.bar2", "staged+store", "stats flushed",
         "ticket/finalize", "exit"]
want = sys.argv[1:] or ["l4.b1.conv1", "l3.b1.conv1", "l3.b1.conv2"]
cfgs = [c for c in C.RESNET50 if c[0] in want]
for name, n, ci, co, hw, k, s, p in cfgs:
    x = C._rand((n, ci, hw, hw)).contiguous(memory_format=C.CL)
    w = C._rand((co, ci, k, k), 0.05).contiguous(memory_format=C.CL)
    sums = torch.zeros(ops.BN_REPLICAS * 2 * co, device="cuda")
    stat = torch.zeros(2 * co, device="cuda")
    tk = torch.zeros((co + 63) // 64, dtype=torch.int32, device="cuda")
    fin = (tk, stat, None, None, 1e-5, 0.1)
    for which in ("fprop+stats", "fprop", "dgrad"):
        def run():
            if which == "fprop+stats":
                return ops.conv_fprop(x, w, s, [p, p], col_sums=sums, bn_finalize=fin)
            if which == "fprop":
                return ops.conv_fprop(x, w, s, [p, p])
            return ops.conv_dgrad(dy, w, x.shape, s, [p, p])
        if which == "dgrad":
            dy = torch.randn_like(y)
        for _ in range(3):
            y = run()
        buf = torch.zeros(4096 * 16, dtype=torch.int64, device="cuda")
        ops.ext().igemm_set_trace(buf)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        ops.ext().igemm_set_trace(None)
        t = buf.view(-1, 16).cpu()
        t = t[t[:, 0] > 0]
        t0 = int(t[:, 0].min())
        print("%s %s: %d CTAs, event time %.1f us" % (name, which, t.shape[0], e0.elapsed_time(e1) * 1e3))
        prev = None
        for kk in range(10):
            col = t[:, kk]
            m = col > 0
            if not m.any():
                continue
            rel = (col[m] - t0).float() / 1e3
            line = "   %-16s first %6.2f  mean %6.2f  last %6.2f us" % (NAMES[kk], rel.min(), rel.mean(), rel.max())
            print(line)
