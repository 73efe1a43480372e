"""Stand-alone correctness + timing check of the tcgen05 implicit-GEMM convolution (run under a
timeout by tests/test_kernels_gpu.py when DDLB_TEST_CONV_TCGEN05=1, or directly under gpurun)."""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ddlbench_b200 import ops  # noqa: E402


def check(n, ci, co, hw, k, stats=False, bias=False, relu=False, seed=0, stride=1):
    torch.manual_seed(seed)
    x = (torch.randn(n, ci, hw, hw, device="cuda") * 0.5).to(torch.bfloat16).contiguous(
        memory_format=torch.channels_last)
    w = (torch.randn(co, ci, k, k, device="cuda") * (1.0 / (k * k * ci) ** 0.5)).to(
        torch.bfloat16).contiguous(memory_format=torch.channels_last)
    b = torch.randn(co, device="cuda").to(torch.bfloat16) if bias else None
    sums = torch.zeros(ops.BN_REPLICAS * 2 * co, device="cuda") if stats else None
    assert ops.conv_tn_supported(x, w, [stride, stride], [k // 2, k // 2], [1, 1], 1), (n, ci, co, hw, k)
    y = ops.conv_tn(x, w, col_sums=sums, bias=b, relu=relu, stride=stride)
    torch.cuda.synchronize()
    if y is None:
        print("conv declined", (n, ci, co, hw, k))
        return False
    ref = F.conv2d(x.float(), w.float(), b.float() if bias else None, stride=stride, padding=k // 2)
    if relu:
        ref = torch.relu(ref)
    err = (y.float() - ref).abs().max().item()
    tol = 0.02 * max(ref.abs().max().item(), 1.0)
    ok = err <= tol and y.is_contiguous(memory_format=torch.channels_last)
    if stats:
        yf = y.float()
        tot = sums.view(ops.BN_REPLICAS, 2, co).sum(0)
        s_ref, q_ref = yf.sum((0, 2, 3)), (yf * yf).sum((0, 2, 3))
        ok = ok and (tot[0] - s_ref).abs().max().item() <= 1e-2 * max(s_ref.abs().max().item(), 1.0) + 0.5 \
            and (tot[1] - q_ref).abs().max().item() <= 1e-2 * q_ref.abs().max().item() + 0.5
    print(f"conv n={n} ci={ci} co={co} hw={hw} k={k} stride={stride} stats={stats} bias={bias} relu={relu}: "
          f"max_err={err:.4f} tol={tol:.4f} {'OK' if ok else 'FAIL'}", flush=True)
    return ok


def check_wgrad(K, M, N, splits=0, seed=0):
    """D[M, N] = A[K, M]^T B[K, N] (MN-major operands, split-K fp32 atomics)."""
    torch.manual_seed(seed)
    a = (torch.randn(K, M, device="cuda") * 0.5).to(torch.bfloat16)
    b = (torch.randn(K, N, device="cuda") * 0.5).to(torch.bfloat16)
    assert ops.gemm_nt_mn_supported(a, b)
    d = ops.gemm_nt_mn(a, b, splits=splits)
    d = ops.gemm_nt_mn(a, b, out=d, splits=splits)          # accumulates: 2x
    torch.cuda.synchronize()
    ref = 2.0 * (a.float().t() @ b.float())
    err = (d - ref).abs().max().item()
    tol = 0.02 * max(ref.abs().max().item(), 1.0)
    ok = err <= tol
    print(f"wgrad K={K} M={M} N={N} splits={splits}: max_err={err:.4f} tol={tol:.4f} "
          f"{'OK' if ok else 'FAIL'} (desc variant {os.environ.get('DDLB_WGRAD_DESC_VARIANT', '0')})", flush=True)
    return ok


def check_conv_wgrad(n, ci, co, hw, k, stride=1, seed=0):
    torch.manual_seed(seed)
    pad = k // 2
    x = (torch.randn(n, ci, hw, hw, device="cuda") * 0.5).to(torch.bfloat16).contiguous(
        memory_format=torch.channels_last)
    ohw = (hw + 2 * pad - k) // stride + 1
    dy = (torch.randn(n, co, ohw, ohw, device="cuda") * 0.5).to(torch.bfloat16).contiguous(
        memory_format=torch.channels_last)
    dw = ops.conv_wgrad(dy, x, (co, ci, k, k), stride)
    torch.cuda.synchronize()
    if dw is None:
        print("conv wgrad declined", (n, ci, co, hw, k, stride))
        return False
    w = torch.zeros(co, ci, k, k, device="cuda", requires_grad=True)
    F.conv2d(x.float(), w, stride=stride, padding=pad).backward(dy.float())
    err = (dw - w.grad).abs().max().item()
    tol = 0.02 * max(w.grad.abs().max().item(), 1.0)
    ok = err <= tol
    print(f"conv wgrad n={n} ci={ci} co={co} hw={hw} k={k} stride={stride}: max_err={err:.4f} "
          f"tol={tol:.4f} {'OK' if ok else 'FAIL'}", flush=True)
    return ok


def check_bnfold(n, k, co, hw, act=1, res=False, seed=0):
    """D = act(BN(c) [+ res]) @ W^T with y / mask written by the same kernel, vs the two-step path."""
    from ddlbench_b200.ops import reference as ref
    torch.manual_seed(seed)
    c = (torch.randn(n, k, hw, hw, device="cuda") * 1.5 + 0.3).to(torch.bfloat16).contiguous(
        memory_format=torch.channels_last)
    r = (torch.randn(n, k, hw, hw, device="cuda")).to(torch.bfloat16).contiguous(
        memory_format=torch.channels_last) if res else None
    gamma = (torch.rand(k, device="cuda") + 0.5).to(torch.bfloat16)
    beta = (torch.randn(k, device="cuda") * 0.2).to(torch.bfloat16)
    w = (torch.randn(co, k, device="cuda") * (1.0 / k ** 0.5)).to(torch.bfloat16)
    stat = ref.bn_stat(c, 1e-5).contiguous()
    y = torch.empty_like(c)
    mask = torch.zeros(c.numel() // 8, dtype=torch.uint8, device="cuda")
    sums = torch.zeros(ops.BN_REPLICAS * 2 * co, device="cuda")
    fold = ops.BnFold(c, r, stat, gamma, beta, act, y, mask)
    before = ops.launches
    d = ops.gemm_tn_bnfold(fold, w, col_sums=sums)
    torch.cuda.synchronize()
    fused = ops.launches - before == 1
    y_ref = ref.bn_apply(c, r, stat, gamma, beta, act)
    d_ref = y_ref.permute(0, 2, 3, 1).reshape(-1, k).float() @ w.float().t()
    err_y = (y.float() - y_ref.float()).abs().max().item()
    err_d = (d.float() - d_ref).abs().max().item()
    yf = y.permute(0, 2, 3, 1).reshape(-1).float()
    bits = ((mask.view(-1, 1) >> torch.arange(8, device="cuda", dtype=torch.uint8)) & 1).bool().view(-1)
    expect = (yf > 0) & (yf < 6) if act == 2 else (yf > 0 if act == 1 else torch.ones_like(yf, dtype=torch.bool))
    tot = sums.view(ops.BN_REPLICAS, 2, co).sum(0)
    ok = fused and err_y <= 0.03 * max(1.0, y_ref.float().abs().max().item()) and \
        err_d <= 0.03 * max(1.0, d_ref.abs().max().item()) and bool(torch.equal(bits, expect)) and \
        (tot[0] - d.float().sum(0)).abs().max().item() <= 1e-2 * max(d.float().sum(0).abs().max().item(), 1.0) + 0.5
    print(f"bnfold n={n} k={k} co={co} hw={hw} act={act} res={res}: fused={fused} err_y={err_y:.4f} "
          f"err_d={err_d:.4f} {'OK' if ok else 'FAIL'}", flush=True)
    return ok


def check_bnbwd(n, ci, co, hw, act=1, res=False, two=False, seed=0):
    """BN backward folded into the 1x1 data-gradient GEMM (gemm_bnbwd_tcgen05.cu) vs the unfused
    device path (bn_bwd_reduce + bn_bwd_apply, then dc @ W): dx, dc, dz, dgamma, dbeta."""
    from ddlbench_b200.ops import reference as ref
    torch.manual_seed(seed)
    cl = torch.channels_last
    rnd = lambda *s_: torch.randn(*s_, device="cuda")                                  # noqa: E731
    c = (rnd(n, co, hw, hw) * 1.5 + 0.3).to(torch.bfloat16).contiguous(memory_format=cl)
    gamma = (torch.rand(co, device="cuda") + 0.5)
    beta = rnd(co) * 0.2
    w = (rnd(co, ci) * (1.0 / co ** 0.5)).to(torch.bfloat16)
    stat = ref.bn_stat(c, 1e-5).contiguous()
    r = rnd(n, co, hw, hw).to(torch.bfloat16).contiguous(memory_format=cl) if res else None
    ws = ops.BnWorkspace(co, "cuda")
    mask = torch.zeros(c.numel() // 8, dtype=torch.uint8, device="cuda") if act else None
    y, _ = ops.bn_forward(c, r, gamma, beta, None, None, 1e-5, 0.1, act, ws, mask_out=mask)
    dy = rnd(n, co, hw, hw).to(torch.bfloat16).contiguous(memory_format=cl)
    dy2 = rnd(n, co, hw, hw).to(torch.bfloat16).contiguous(memory_format=cl) if two else None
    dg0, db0, dg1, db1 = (torch.zeros(co, device="cuda") for _ in range(4))
    dc_ref, dz_ref = ops.bn_backward(dy, dy2, y, c, stat, gamma, act, res, dg0, db0, False, ws, mask=mask)
    dx_ref = dc_ref.permute(0, 2, 3, 1).reshape(-1, co).float() @ w.float()
    wt = w.t().contiguous()
    if not ops.bn_bwd_fold_supported(c, dy, dy2, act, mask, wt):
        print(f"bnbwd n={n} ci={ci} co={co} hw={hw}: shape not supported (skipped)", flush=True)
        return True
    before = ops.launches
    dc, dz, dx = ops.bn_backward_dgrad_fold(dy, dy2, c, stat, gamma, act, res, dg1, db1, False, ws, mask, wt)
    torch.cuda.synchronize()
    fused = ops.launches - before == 2

    def err(a, b):
        return (a.float() - b.float()).abs().max().item() / max(1.0, b.float().abs().max().item())

    e = {"dx": err(dx, dx_ref), "dc": err(dc, dc_ref), "dgamma": err(dg1, dg0), "dbeta": err(db1, db0)}
    if res:
        e["dz"] = err(dz, dz_ref)
    ok = fused and all(v <= 0.03 for v in e.values())
    print(f"bnbwd n={n} ci={ci} co={co} hw={hw} act={act} res={res} two={two}: fused={fused} "
          + " ".join(f"{k}={v:.4f}" for k, v in e.items()) + (" OK" if ok else " FAIL"), flush=True)
    return ok


def bench(n, ci, co, hw, k=3, iters=20):
    x = torch.randn(n, ci, hw, hw, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    w = torch.randn(co, ci, k, k, device="cuda").to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    out = {}
    for name, fn in (("tcgen05", lambda: ops.conv_tn(x, w)),
                     ("cudnn", lambda: F.conv2d(x, w, padding=k // 2))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        out[name + "_us"] = round(e0.elapsed_time(e1) / iters * 1e3, 1)
    out["shape"] = [n, ci, co, hw, k]
    out["tflops_tcgen05"] = round(2.0 * n * hw * hw * co * ci * k * k / out["tcgen05_us"] / 1e6, 1)
    return out


def main():
    ok = True
    # ResNet-50 3x3 layers at the benchmark batch, a padded-edge case per box shape, 5x5
    for args in [(32, 64, 64, 56, 3), (32, 128, 128, 28, 3), (32, 256, 256, 14, 3), (32, 512, 512, 7, 3),
                 (3, 64, 64, 9, 3), (5, 128, 192, 13, 3), (2, 64, 128, 10, 5), (1, 64, 64, 1, 3)]:
        ok &= check(*args)
    ok &= check(32, 256, 256, 14, 3, stats=True)
    # strided: ResNet-50's 3x3/s2 and 1x1/s2 (downsample) layers, odd sizes
    for args in [(32, 128, 128, 56, 3), (32, 256, 512, 56, 1), (4, 64, 64, 15, 3), (3, 128, 64, 9, 1)]:
        ok &= check(*args, stride=2)
    ok &= check(8, 256, 256, 28, 3, stats=True, stride=2)
    ok &= check(7, 64, 64, 11, 3, stats=True, bias=True, relu=True)
    if os.environ.get("CHECK_BNFOLD", "1") == "1":
        for args in [(32, 64, 256, 56, 1, False), (32, 256, 64, 56, 1, True), (32, 512, 2048, 7, 1, False),
                     (5, 128, 64, 9, 2, True), (3, 64, 136, 11, 0, False)]:
            ok &= check_bnfold(*args)
    if os.environ.get("CHECK_BNBWD", "1") == "1":
        for args in [(32, 64, 256, 56, 1, True, True), (32, 256, 64, 56, 1, False, False),
                     (32, 2048, 512, 7, 1, False, False), (5, 64, 128, 9, 2, True, False),
                     (3, 136, 64, 11, 0, False, True)]:
            ok &= check_bnbwd(*args)
    if os.environ.get("CHECK_WGRAD", "1") == "1":
        for args in [(64, 128, 64), (128, 128, 128), (6272, 256, 1024), (100352, 64, 256), (1000, 72, 200),
                     (200, 136, 64)]:
            ok &= check_wgrad(*args)
        ok &= check_wgrad(4096, 256, 128, splits=1)
        for args in [(32, 64, 64, 56, 3), (32, 256, 256, 14, 3), (4, 128, 128, 15, 3, 2), (4, 256, 512, 14, 1, 2),
                     (3, 64, 128, 9, 3)]:
            ok &= check_conv_wgrad(*args)
    if ok and os.environ.get("BENCH", "1") == "1":
        res = [bench(32, 64, 64, 56), bench(32, 128, 128, 28), bench(32, 256, 256, 14), bench(32, 512, 512, 7)]
        for r in res:
            print(json.dumps(r))
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(res, open("gpurun_out/conv_bench.json", "w"), indent=1)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
