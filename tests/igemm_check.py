"""Numerics (and, with BENCH=1, timings against cuDNN) of the persistent tcgen05 implicit-GEMM family
— ``ops.conv_fprop`` / ``ops.conv_dgrad`` / ``ops.conv_wgrad_into`` / ``ops.gemm_ig`` /
``ops.wgrad_gemm_into`` — against fp32 ``F.conv2d`` + autograd on every convolution shape of
ResNet-50 (batch 32, ImageNet), VGG-16 at 224x224, MobileNetV2-style odd channel counts and the FC
layers.  Run as a script (``tests/test_kernels_gpu.py`` runs it in a child process with a timeout so a
protocol bug cannot take the session down).  Exit code 0 = everything within tolerance.

    python tests/igemm_check.py            # numerics
    BENCH=1 OUT=gpurun_out/conv_vs_cudnn.json python tests/igemm_check.py
"""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ddlbench_b200 import ops  # noqa: E402

DEV = "cuda"
BF = torch.bfloat16
CL = torch.channels_last
aten = torch.ops.aten

# (name, batch, Cin, Cout, H(=W) input, k, stride, pad)
RESNET50 = [
    ("l1.conv1", 32, 64, 64, 56, 1, 1, 0), ("l1.conv2", 32, 64, 64, 56, 3, 1, 1),
    ("l1.conv3", 32, 64, 256, 56, 1, 1, 0), ("l1.b1.conv1", 32, 256, 64, 56, 1, 1, 0),
    ("l2.conv1", 32, 256, 128, 56, 1, 1, 0), ("l2.conv2.s2", 32, 128, 128, 56, 3, 2, 1),
    ("l2.conv3", 32, 128, 512, 28, 1, 1, 0), ("l2.down.s2", 32, 256, 512, 56, 1, 2, 0),
    ("l2.b1.conv1", 32, 512, 128, 28, 1, 1, 0), ("l2.b1.conv2", 32, 128, 128, 28, 3, 1, 1),
    ("l3.conv1", 32, 512, 256, 28, 1, 1, 0), ("l3.conv2.s2", 32, 256, 256, 28, 3, 2, 1),
    ("l3.conv3", 32, 256, 1024, 14, 1, 1, 0), ("l3.down.s2", 32, 512, 1024, 28, 1, 2, 0),
    ("l3.b1.conv1", 32, 1024, 256, 14, 1, 1, 0), ("l3.b1.conv2", 32, 256, 256, 14, 3, 1, 1),
    ("l4.conv1", 32, 1024, 512, 14, 1, 1, 0), ("l4.conv2.s2", 32, 512, 512, 14, 3, 2, 1),
    ("l4.conv3", 32, 512, 2048, 7, 1, 1, 0), ("l4.down.s2", 32, 1024, 2048, 14, 1, 2, 0),
    ("l4.b1.conv1", 32, 2048, 512, 7, 1, 1, 0), ("l4.b1.conv2", 32, 512, 512, 7, 3, 1, 1),
]
VGG224 = [
    ("vgg.64@224", 4, 64, 64, 224, 3, 1, 1), ("vgg.128@112", 4, 64, 128, 112, 3, 1, 1),
    ("vgg.256@56", 4, 128, 256, 56, 3, 1, 1), ("vgg.512@28", 4, 256, 512, 28, 3, 1, 1),
    ("vgg.512@14", 4, 512, 512, 14, 3, 1, 1),
]
ODD = [
    ("mbv2.24->144", 5, 24, 144, 19, 1, 1, 0), ("mbv2.144->24", 5, 144, 24, 19, 1, 1, 0),
    ("mbv2.320->1280", 3, 320, 1280, 7, 1, 1, 0), ("odd.5x5.s2", 3, 40, 72, 17, 5, 2, 2),
    ("odd.3x3.pad0", 2, 64, 64, 10, 3, 1, 0), ("odd.7x7", 2, 16, 32, 20, 7, 1, 3),
    ("odd.3x3.s2.oddHW", 3, 64, 128, 15, 3, 2, 1), ("tiny.1px", 1, 64, 64, 1, 3, 1, 1),
]


def _rand(shape, scale=1.0):
    return (torch.randn(*shape, device=DEV) * scale).to(BF)


def _err(a, b):
    a, b = a.float(), b.float()
    return (a - b).abs().max().item(), b.abs().max().item()


def _time(fn, iters=20, warm=3):
    """us per call with the host launch cost removed: `iters` calls captured in one CUDA graph."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(g):
            for _ in range(iters):
                fn()
    except Exception:  # noqa: BLE001 - not capturable: time eagerly
        g = None
        torch.cuda.synchronize()
    run = g.replay if g is not None else (lambda: [fn() for _ in range(iters)])
    run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        run()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (3 * iters) * 1e3      # us


def check_conv(name, n, ci, co, hw, k, s, p, bias_relu=False, bench=False):
    torch.manual_seed(hash(name) % 1000)
    x = _rand((n, ci, hw, hw)).contiguous(memory_format=CL)
    w = _rand((co, ci, k, k), (2.0 / (ci * k * k)) ** 0.5).contiguous(memory_format=CL)
    bias = _rand((co,), 0.5) if bias_relu else None
    assert ops.igemm_conv_supported(x, w.shape, [s, s], [p, p], [1, 1], 1, co), name
    # ---- forward (+ fused statistics and finalize)
    sums = torch.zeros(ops.BN_REPLICAS * 2 * co, device=DEV)
    stat = torch.zeros(2 * co, device=DEV)
    rm, rv = torch.zeros(co, device=DEV), torch.ones(co, device=DEV)
    ticket = torch.zeros((co + 63) // 64, dtype=torch.int32, device=DEV)
    fin = None if bias_relu else (ticket, stat, rm, rv, 1e-5, 0.1)
    y = None
    for _ in range(2):        # twice: workspace + ticket must self-reset
        y = ops.conv_fprop(x, w, s, [p, p], col_sums=None if bias_relu else sums, bias=bias,
                           relu=bias_relu, bn_finalize=fin)
    assert y is not None, name + ": fprop declined"
    y_ref = F.conv2d(x.float(), w.float(), bias.float() if bias_relu else None, s, p)
    if bias_relu:
        y_ref = y_ref.relu()
    e, m = _err(y, y_ref)
    ok = e <= 0.012 * max(m, 1.0)
    msgs = ["fprop %.4f/%.2f" % (e, m)]
    if not bias_relu:
        yb = y.float()
        mean_ref = yb.mean((0, 2, 3))
        var_ref = yb.var((0, 2, 3), unbiased=False)
        e1 = (stat[:co] - mean_ref).abs().max().item()
        e2 = (stat[co:] - torch.rsqrt(var_ref + 1e-5)).abs().max().item() / max(stat[co:].abs().max().item(), 1e-6)
        ok = ok and e1 < 2e-3 * max(1.0, mean_ref.abs().max().item()) and e2 < 5e-3
        ok = ok and float(sums.abs().sum()) == 0.0 and int(ticket.abs().sum()) == 0
        # two calls with momentum 0.1 starting from 0: rm = (1 - 0.9^2) * mean
        e3 = (rm - 0.19 * mean_ref).abs().max().item()
        ok = ok and e3 < 2e-3 * max(1.0, mean_ref.abs().max().item())
        msgs.append("stat %.1e %.1e rm %.1e" % (e1, e2, e3))
    # ---- data gradient
    dy = _rand(tuple(y.shape)).contiguous(memory_format=CL)
    dx = ops.conv_dgrad(dy, w, x.shape, s, [p, p])
    assert dx is not None, name + ": dgrad declined"
    dx_ref, dw_ref, _ = aten.convolution_backward(dy.float(), x.float(), w.float(), None, [s, s], [p, p], [1, 1],
                                                  False, [0, 0], 1, [True, True, False])
    e, m = _err(dx, dx_ref)
    ok = ok and e <= 0.012 * max(m, 1.0)
    msgs.append("dgrad %.4f/%.2f" % (e, m))
    # ---- weight gradient (bf16 KRSC view, as the parameter store hands it over), then accumulate
    gw = torch.full((co, k, k, ci), 7.0, device=DEV, dtype=BF).permute(0, 3, 1, 2)
    assert ops.conv_wgrad_into(dy, x, gw, s, [p, p], False), name + ": wgrad declined"
    e, m = _err(gw, dw_ref)
    ok = ok and e <= 0.012 * max(m, 1.0)
    msgs.append("wgrad %.4f/%.2f" % (e, m))
    assert ops.conv_wgrad_into(dy, x, gw, s, [p, p], True)
    e, m = _err(gw, 2 * dw_ref)
    ok = ok and e <= 0.02 * max(m, 1.0)
    g32 = torch.zeros((co, k, k, ci), device=DEV).permute(0, 3, 1, 2)
    assert ops.conv_wgrad_into(dy, x, g32, s, [p, p], False)
    e, m = _err(g32, dw_ref)
    ok = ok and e <= 0.004 * max(m, 1.0)
    sk = ops._ig_ws(gw.device)                      # split-K tickets of fprop / dgrad reset themselves
    ok = ok and int(sk[1].abs().sum()) == 0
    print("%-18s n=%d %d->%d %dx%d k%d s%d: %s  %s" % (name, n, ci, co, hw, hw, k, s, "  ".join(msgs),
                                                     "OK" if ok else "FAIL"), flush=True)
    row = None
    if bench:
        args = (None, [s, s], [p, p], [1, 1], False, [0, 0], 1)
        t = {
            "fprop_us": _time(lambda: ops.conv_fprop(x, w, s, [p, p], col_sums=sums, bn_finalize=fin)),
            "fprop_cudnn_us": _time(lambda: aten.convolution(x, w, *args)),
            "dgrad_us": _time(lambda: ops.conv_dgrad(dy, w, x.shape, s, [p, p])),
            "dgrad_cudnn_us": _time(lambda: aten.convolution_backward(dy, x, w, *args, [True, False, False])),
            "wgrad_us": _time(lambda: ops.conv_wgrad_into(dy, x, gw, s, [p, p], False)),
            "wgrad_cudnn_us": _time(lambda: aten.convolution_backward(dy, x, w, *args, [False, True, False])),
        }
        oh = y.shape[2]
        flops = 2.0 * n * oh * oh * co * ci * k * k
        row = {"layer": name, "n": n, "cin": ci, "cout": co, "hw": hw, "k": k, "stride": s,
               **{kk: round(v, 2) for kk, v in t.items()},
               "fprop_tflops": round(flops / t["fprop_us"] / 1e6, 1),
               "fprop_vs_cudnn": round(t["fprop_cudnn_us"] / t["fprop_us"], 3),
               "dgrad_vs_cudnn": round(t["dgrad_cudnn_us"] / t["dgrad_us"], 3),
               "wgrad_vs_cudnn": round(t["wgrad_cudnn_us"] / t["wgrad_us"], 3)}
        print("    " + json.dumps(row), flush=True)
    return ok, row


def check_linear(m, k, n):
    torch.manual_seed(m + k + n)
    a, w, b = _rand((m, k)), _rand((n, k), k ** -0.5), _rand((n,), 0.5)
    y = ops.gemm_ig(a, w, bias=b, relu=True)
    y_ref = (a.float() @ w.float().t() + b.float()).relu()
    e, mx = _err(y, y_ref)
    ok = y is not None and e <= 0.012 * max(mx, 1.0)
    dy = _rand((m, n))
    dx = ops.gemm_ig(dy, w, b_mn=True)
    e2, mx2 = _err(dx, dy.float() @ w.float())
    ok = ok and e2 <= 0.012 * max(mx2, 1.0)
    g = torch.zeros((n, k), device=DEV, dtype=BF)
    assert ops.wgrad_gemm_into(dy, a, g, False)
    e3, mx3 = _err(g, dy.float().t() @ a.float())
    ok = ok and e3 <= 0.012 * max(mx3, 1.0)
    print("linear m=%d k=%d n=%d: fwd %.4f/%.2f dgrad %.4f/%.2f wgrad %.4f/%.2f %s" % (
        m, k, n, e, mx, e2, mx2, e3, mx3, "OK" if ok else "FAIL"), flush=True)
    return ok


def check_gemm_stats(m, k, n):
    """Flat GEMM with statistics + finalize (the stem's im2col GEMM goes through this)."""
    torch.manual_seed(m)
    a, w = _rand((m, k)), _rand((n, k), k ** -0.5)
    sums = torch.zeros(ops.BN_REPLICAS * 2 * n, device=DEV)
    stat = torch.zeros(2 * n, device=DEV)
    ticket = torch.zeros((n + 63) // 64, dtype=torch.int32, device=DEV)
    y = ops.gemm_ig(a, w, col_sums=sums, bn_finalize=(ticket, stat, None, None, 1e-5, 0.1))
    yb = y.float()
    e, mx = _err(y, a.float() @ w.float().t())
    e1 = (stat[:n] - yb.mean(0)).abs().max().item()
    ok = e <= 0.012 * max(mx, 1.0) and e1 < 2e-3 and float(sums.abs().sum()) == 0.0
    print("gemm+stats m=%d k=%d n=%d: %.4f/%.2f mean %.1e %s" % (m, k, n, e, mx, e1, "OK" if ok else "FAIL"), flush=True)
    return ok


def main():
    bench = os.environ.get("BENCH", "0") == "1"
    quick = os.environ.get("QUICK", "0") == "1"
    only = [t for t in os.environ.get("ONLY", "").split(",") if t]      # layer-name substrings
    ok, rows = True, []
    groups = [(RESNET50, False), (ODD, False), (VGG224, True)]
    for cfgs, br in groups:
        for c in (cfgs[:4] if quick else cfgs):
            if only and not any(t in c[0] for t in only):
                continue
            o, row = check_conv(*c, bias_relu=br, bench=bench and cfgs is RESNET50)
            ok = ok and o
            if row:
                rows.append(row)
    for shp in ([] if only else [(32, 2048, 1000), (8, 25088, 4096), (32, 4096, 1000), (5, 64, 24)]):
        ok = check_linear(*shp) and ok
    for shp in ([] if only else [(32 * 112 * 112, 152, 64), (1000, 72, 200)]):
        ok = check_gemm_stats(*shp) and ok
    if bench and rows:
        tot = {k: round(sum(r[k] for r in rows), 1) for k in rows[0] if k.endswith("_us")}
        out = {"note": "ResNet-50 batch 32, per-layer, 20 back-to-back launches captured in one CUDA graph (no host launch "
                       "cost on either side), replayed 3x after warm-up (L2-warm), CUDA events; "
                       "fprop includes the fused BatchNorm statistics + finalize, cuDNN's does not",
               "sum_us": tot, "layers": rows}
        path = os.environ.get("OUT", "gpurun_out/conv_vs_cudnn.json")
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        with open(path, "w") as f:
            json.dump(out, f, indent=1)
        print("SUM", json.dumps(tot), flush=True)
    print("igemm_check", "OK" if ok else "FAIL", flush=True)
    return 0 if ok else 1


if __name__ == "__main__":
    raise SystemExit(main())
