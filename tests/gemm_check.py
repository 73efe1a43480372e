"""Stand-alone correctness + throughput check of the tcgen05 GEMM (run under a timeout by
tests/test_kernels_gpu.py, and directly under gpurun for profiling)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ddlbench_b200 import ops  # noqa: E402


def check(M, N, K, stats=False, bias=False, relu=False, seed=0):
    torch.manual_seed(seed)
    a = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
    b = (torch.randn(N, K, device="cuda") * 0.5).to(torch.bfloat16)
    bias_t = torch.randn(N, device="cuda").to(torch.bfloat16) if bias else None
    sums = torch.zeros(ops.BN_REPLICAS * 2 * N, device="cuda") if stats else None
    assert ops.gemm_tn_supported(a, b), (M, N, K)
    d = ops.gemm_tn(a, b, col_sums=sums, bias=bias_t, relu=relu)
    torch.cuda.synchronize()
    if stats:
        # same GEMM with the fused BatchNorm finalize (last CTA): (mean, invstd) + running stats
        sums2 = torch.zeros_like(sums)
        ticket = torch.zeros(1, dtype=torch.int32, device="cuda")
        stat = torch.empty(2 * N, device="cuda")
        rm, rv = torch.zeros(N, device="cuda"), torch.ones(N, device="cuda")
        for _ in range(2):
            d2 = ops.gemm_tn(a, b, col_sums=sums2, bn_finalize=(ticket, stat, rm, rv, 1e-5, 0.1))
        torch.cuda.synchronize()
        df2 = d2.float()
        mean_ref, var_ref = df2.mean(0), df2.var(0, unbiased=False)
        fin_ok = (stat[:N] - mean_ref).abs().max().item() < 2e-2 * max(1.0, mean_ref.abs().max().item()) \
            and ((stat[N:] - torch.rsqrt(var_ref + 1e-5)).abs() / torch.rsqrt(var_ref + 1e-5)).max().item() < 2e-2 \
            and sums2.abs().max().item() == 0.0 and int(ticket.item()) == 0 \
            and (rm - 0.19 * mean_ref).abs().max().item() < 2e-2 * max(1.0, mean_ref.abs().max().item())
        if not fin_ok:
            print("  fused BN finalize FAILED", (stat[:N] - mean_ref).abs().max().item(), sums2.abs().max().item(), int(ticket.item()))
            return False
    ref = a.float() @ b.float().t()
    if bias:
        ref = ref + bias_t.float()
    if relu:
        ref = torch.relu(ref)
    err = (d.float() - ref).abs().max().item()
    tol = 0.02 * max(ref.abs().max().item(), 1.0)
    ok = err <= tol
    if stats:
        df = d.float()
        tot = sums.view(ops.BN_REPLICAS, 2, N).sum(0)
        s_err = (tot[0] - df.sum(0)).abs().max().item()
        q_err = (tot[1] - (df * df).sum(0)).abs().max().item()
        ok = ok and s_err <= 1e-2 * max(df.sum(0).abs().max().item(), 1.0) + 0.5 \
            and q_err <= 1e-2 * (df * df).sum(0).abs().max().item() + 0.5
    print(f"gemm M={M} N={N} K={K} stats={stats} bias={bias} relu={relu}: max_err={err:.4f} "
          f"tol={tol:.4f} {'OK' if ok else 'FAIL'}", flush=True)
    return ok


def bench(M, N, K, iters=20):
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    b = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    def timeit(fn):
        for _ in range(3):
            fn()
        ts = []
        for _ in range(iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        return ts[len(ts) // 2]
    t_mine = timeit(lambda: ops.gemm_tn(a, b, out=out))
    t_lib = timeit(lambda: torch.matmul(a, b.t(), out=out))
    fl = 2.0 * M * N * K
    by = 2.0 * (M * K + N * K + M * N)
    return {"M": M, "N": N, "K": K, "tcgen05_ms": round(t_mine, 4), "cublas_ms": round(t_lib, 4),
            "tcgen05_tflops": round(fl / t_mine / 1e9, 1), "cublas_tflops": round(fl / t_lib / 1e9, 1),
            "tcgen05_gbs": round(by / t_mine / 1e6, 1)}


def main():
    ok = True
    shapes = [(128, 128, 64), (128, 64, 64), (256, 128, 128), (128, 128, 512), (1000, 200, 72),
              (100352, 64, 64), (100352, 256, 64), (25088, 512, 128), (6272, 1024, 256),
              (1568, 2048, 512), (1568, 512, 2048), (32, 1000, 2048), (333, 1000, 2048)]
    for (M, N, K) in shapes:
        ok &= check(M, N, K)
    ok &= check(4096, 256, 64, stats=True)
    ok &= check(1568, 2048, 512, stats=True)
    ok &= check(777, 136, 200, stats=True)
    ok &= check(512, 4096, 1024, bias=True, relu=True)
    ok &= check(64, 1000, 2048, bias=True)
    if not ok:
        print("GEMM CHECK FAILED")
        return 1
    if os.environ.get("GEMM_BENCH", "1") == "1":
        rows = [bench(*s) for s in [(100352, 64, 64), (100352, 256, 64), (100352, 64, 256),
                                    (25088, 512, 128), (25088, 128, 512), (6272, 1024, 256),
                                    (6272, 256, 1024), (1568, 2048, 512), (1568, 512, 2048),
                                    (8192, 8192, 8192)]]
        for r in rows:
            print(json.dumps(r), flush=True)
        os.makedirs("gpurun_out", exist_ok=True)
        with open("gpurun_out/gemm_bench.json", "w") as f:
            json.dump(rows, f, indent=1)
    print("GEMM CHECK OK")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
