"""Numerics and timings of the GEMM cores — the CTA-pair ``cta_group::2`` kernel (``ops.gemm_pair``,
csrc/gemm2cta_tcgen05.cu) and the single-CTA persistent kernel (``ops.gemm_ig``) — against an fp32
reference and cuBLAS (``torch.matmul``).  Run as a script in a child process (a protocol bug traps
after ~2 s instead of hanging, the parent adds a timeout on top).

    python tests/gemm_pair_check.py
    BENCH=1 OUT=gpurun_out/gemm_tcgen05_vs_cublas.json python tests/gemm_pair_check.py
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ddlbench_b200 import ops  # noqa: E402

DEV = "cuda"
BF = torch.bfloat16

# the nine ResNet-50 1x1-convolution GEMMs of the round-1 table + the square one
BENCH_SHAPES = [(100352, 64, 64), (100352, 256, 64), (100352, 64, 256), (25088, 512, 128),
                (25088, 128, 512), (6272, 1024, 256), (6272, 256, 1024), (1568, 2048, 512),
                (1568, 512, 2048), (8192, 8192, 8192), (4096, 4096, 4096), (16384, 4096, 1024)]
# numerics: full tiles, ragged M / N / K tails, one-tile problems, both BLOCK_N variants, bias + ReLU
CHECK_SHAPES = [(256, 256, 64), (512, 512, 512), (1000, 328, 200), (4096, 1024, 1024), (300, 128, 136),
                (32, 64, 64), (2048, 72, 512), (777, 4096, 264), (8192, 256, 2048)]


def _rand(shape, scale=1.0):
    return (torch.randn(*shape, device=DEV) * scale).to(BF)


def _time(fn, iters, reps=3):
    """us per call: `iters` calls captured in one CUDA graph, replayed `reps` times, CUDA events."""
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * iters) * 1e3


def check(m, n, k, bias_relu):
    torch.manual_seed(m + n + k)
    a, w = _rand((m, k)), _rand((n, k), k ** -0.5)
    b = _rand((n,), 0.5) if bias_relu else None
    y = ops.gemm_pair(a, w, bias=b, relu=bias_relu)
    if y is None:
        print("pair m=%d n=%d k=%d: DECLINED" % (m, n, k), flush=True)
        return False
    ref = a.float() @ w.float().t()
    if bias_relu:
        ref = (ref + b.float()).relu()
    e = (y.float() - ref).abs().max().item()
    mx = ref.abs().max().item()
    ok = e <= 0.012 * max(mx, 1.0)
    print("pair m=%d n=%d k=%d bias_relu=%d: err %.4f / max %.2f %s" % (m, n, k, bias_relu, e, mx,
                                                                       "OK" if ok else "FAIL"), flush=True)
    return ok


def bench(m, n, k):
    a, w = _rand((m, k)), _rand((n, k), k ** -0.5)
    out = torch.empty((m, n), device=DEV, dtype=BF)
    flops = 2.0 * m * n * k
    iters = 20 if flops < 1e11 else 5
    row = {"M": m, "N": n, "K": k}
    saved = ops.GEMM_PAIR
    try:
        ops.GEMM_PAIR = False
        t_ig = _time(lambda: ops.gemm_ig(a, w), iters)
    finally:
        ops.GEMM_PAIR = saved
    t_pair = _time(lambda: ops.gemm_pair(a, w, out=out), iters) if ops.gemm_pair(a, w, out=out) is not None else None
    wt = w.t()
    t_cublas = _time(lambda: torch.matmul(a, wt, out=out), iters)
    row["single_cta_us"] = round(t_ig, 2)
    row["pair_us"] = round(t_pair, 2) if t_pair else None
    row["cublas_us"] = round(t_cublas, 2)
    best = min(t for t in (t_ig, t_pair) if t)
    row["best"] = "pair" if t_pair and t_pair <= t_ig else "single_cta"
    row["dispatch"] = "pair" if (m >= ops.GEMM_PAIR_MIN[0] and n >= ops.GEMM_PAIR_MIN[1] and
                                 k >= ops.GEMM_PAIR_MIN[2]) else "single_cta"
    t_disp = t_pair if row["dispatch"] == "pair" and t_pair else t_ig
    row["dispatch_us"] = round(t_disp, 2)
    row["tcgen05_tflops"] = round(flops / t_disp / 1e6, 1)
    row["best_tflops"] = round(flops / best / 1e6, 1)
    row["cublas_tflops"] = round(flops / t_cublas / 1e6, 1)
    row["dispatch_vs_cublas"] = round(t_cublas / t_disp, 3)
    row["best_vs_cublas"] = round(t_cublas / best, 3)
    print("    " + json.dumps(row), flush=True)
    return row


def main():
    ok = True
    for i, s in enumerate(CHECK_SHAPES):
        ok = check(*s, bias_relu=bool(i % 2)) and ok
    # dispatch: a big stats-free GEMM through gemm_ig lands on the pair kernel and matches
    a, w = _rand((4096, 1024)), _rand((512, 1024), 1024 ** -0.5)
    y = ops.gemm_ig(a, w)
    e = (y.float() - a.float() @ w.float().t()).abs().max().item()
    ok = ok and e < 0.1
    print("dispatch via gemm_ig: err %.4f" % e, flush=True)
    if os.environ.get("BENCH", "0") == "1":
        rows = [bench(*s) for s in BENCH_SHAPES]
        out = {"note": "D[M,N] = A[M,K] B[N,K]^T bf16; per shape 20 (5 for the big ones) back-to-back launches "
                       "captured in one CUDA graph, replayed 3x after warm-up, CUDA events; `single_cta` = "
                       "igemm_tcgen05.cu flat-GEMM path, `pair` = gemm2cta_tcgen05.cu (cta_group::2, 256x256 "
                       "tiles), `dispatch` = what ops.gemm_ig picks; cuBLAS = torch.matmul(out=)",
               "rows": rows}
        path = os.environ.get("OUT", "gpurun_out/gemm_tcgen05_vs_cublas.json")
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        with open(path, "w") as f:
            json.dump(out, f, indent=1)
    print("gemm_pair_check", "OK" if ok else "FAIL", flush=True)
    return 0 if ok else 1


if __name__ == "__main__":
    raise SystemExit(main())
