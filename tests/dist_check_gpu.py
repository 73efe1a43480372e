"""Multi-GPU checks of the hand-written collectives (run with torchrun on >= 2 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/dist_check_gpu.py

1. generic all-reduce (one-shot / two-shot p2p / NVLS multimem) vs torch.distributed (NCCL)
2. fused all-reduce + SGD kernel vs "all_reduce, then SGD on the full fp32 master"
3. data-parallel Trainer: fused path vs NCCL-baseline path produce the same weights, and all
   ranks hold identical weights
4. all-reduce bandwidth sweep (bus GB/s) for the roofline report
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ddlbench_b200 import ops  # noqa: E402
from ddlbench_b200.dist.symm import SymmContext  # noqa: E402
from ddlbench_b200.harness import Trainer  # noqa: E402
from ddlbench_b200.parallel import dp  # noqa: E402
from ddlbench_b200.parallel.collectives import SymmAllreduce  # noqa: E402


def log(*a):
    if dist.get_rank() == 0:
        print(*a, flush=True)


def check_generic(ctx):
    ok = True
    ar = SymmAllreduce(ctx)
    rank, world = ctx.rank, ctx.world
    for dtype in (torch.float32, torch.bfloat16):
        for n in (8, 1000, 65536 + 24, 3_000_001):
            for algo in ("oneshot", "p2p", "auto"):
                if algo == "oneshot" and n > 200000:
                    continue
                torch.manual_seed(100 + rank)
                x = torch.randn(n, device="cuda").to(dtype)
                ref = x.float().clone()
                dist.all_reduce(ref)
                ref /= world
                y = ar.allreduce_(x.clone(), average=True, algo=algo)
                torch.cuda.synchronize()
                tol = 2e-2 if dtype == torch.bfloat16 else 1e-5
                err = (y.float() - ref).abs().max().item()
                good = err <= tol * max(1.0, ref.abs().max().item())
                ok &= good
                if not good or rank == 0 and n == 3_000_001:
                    print(f"[rank {rank}] allreduce {dtype} n={n} algo={algo} err={err:.3e} "
                          f"{'OK' if good else 'FAIL'}", flush=True)
    ctx.check()
    return ok, ar


def check_fused_sgd(ctx, use_mm):
    """Bucket of n elements; every rank has its own bf16 grads; expected result = SGD with the
    averaged gradient applied to a replicated fp32 master."""
    rank, world = ctx.rank, ctx.world
    n = world * 8 * 5000
    torch.manual_seed(7)
    w0 = torch.randn(n, device="cuda")
    gsym, wsym = ctx.alloc(n, torch.bfloat16), ctx.alloc(n, torch.bfloat16)
    wsym.tensor.copy_(w0.to(torch.bfloat16))
    sl = n // world
    master = w0[rank * sl:(rank + 1) * sl].clone()
    mom = torch.zeros(sl, device="cuda")
    hyper = torch.tensor([0.1, 0.9, 1e-4, 1.0 / world], device="cuda")
    ref_master, ref_mom = w0.clone(), torch.zeros(n, device="cuda")
    ok = True
    for step in range(3):
        torch.manual_seed(1000 * step + rank)
        g = torch.randn(n, device="cuda").to(torch.bfloat16)
        gsym.tensor.copy_(g)
        gsum = g.float().clone()
        dist.all_reduce(gsum)
        gk = gsum / world + 1e-4 * ref_master
        ref_mom.mul_(0.9).add_(gk)
        ref_master.add_(ref_mom, alpha=-0.1)
        torch.cuda.synchronize()
        dist.barrier()
        ops.ext().ar_sgd_twoshot(gsym.ptrs, wsym.ptrs, ctx.flags.ptrs, gsym.mc_ptr, wsym.mc_ptr,
                                 master, mom, hyper, ctx.epoch, ctx.status, 0, sl, rank, 16, use_mm)
        torch.cuda.synchronize()
        dist.barrier()
        err_w = (wsym.tensor.float() - ref_master).abs().max().item()
        err_m = (master - ref_master[rank * sl:(rank + 1) * sl]).abs().max().item()
        good = err_w < 3e-2 and err_m < 2e-2
        ok &= good
        if rank == 0 or not good:
            print(f"[rank {rank}] fused allreduce+SGD (multimem={use_mm}) step {step}: "
                  f"weight err {err_w:.3e} master err {err_m:.3e} {'OK' if good else 'FAIL'}",
                  flush=True)
    ctx.check()
    return ok


def check_trainer():
    rank = dist.get_rank()
    res = {}
    for mode in ("nccl", "fused"):
        tr = Trainer("resnet18", "cifar10", batch_size=16, device=f"cuda:{torch.cuda.current_device()}",
                     dtype=torch.bfloat16, lr=0.05, momentum=0.9, weight_decay=1e-4, cuda_graph=True,
                     distributed=True, allreduce=mode, seed=1, bucket_elems=1 << 20)
        g = torch.Generator().manual_seed(50 + rank)
        x = torch.randn(16, 3, 32, 32, generator=g).cuda()
        t = torch.randint(0, 10, (16,), generator=g).cuda()
        losses = [float(tr.step(x, t)[0]) for _ in range(8)]
        torch.cuda.synchronize()
        tr.check()
        w = tr.ex.store.w_flat[0].float().clone()
        wmax = w.clone()
        wmin = w.clone()
        dist.all_reduce(wmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(wmin, op=dist.ReduceOp.MIN)
        res[mode] = (losses, w, (wmax - wmin).abs().max().item())
        del tr
    spread_ok = res["fused"][2] == 0.0 and res["nccl"][2] < 1e-6
    l_f, l_n = res["fused"][0], res["nccl"][0]
    learn_ok = l_f[-1] < l_f[0] and l_n[-1] < l_n[0]
    log("trainer DP: fused losses", [round(v, 3) for v in l_f])
    log("trainer DP: nccl  losses", [round(v, 3) for v in l_n])
    log("trainer DP: weight spread across ranks fused=%g nccl=%g" % (res["fused"][2], res["nccl"][2]))
    return spread_ok and learn_ok


def sweep(ar, ctx):
    rows = []
    world = ctx.world
    for nbytes in (64 << 10, 1 << 20, 16 << 20, 64 << 20):
        n = nbytes // 2
        x = torch.randn(n, device="cuda").to(torch.bfloat16)
        for name, fn in (("fused_p2p", lambda: ar.allreduce_(x, True, "p2p")),
                         ("fused_auto", lambda: ar.allreduce_(x, True, "auto")),
                         ("nccl", lambda: dist.all_reduce(x))):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1) / 10], device="cuda")
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            bus = nbytes * 2 * (world - 1) / world / (ms.item() * 1e-3) / 1e9
            rows.append({"bytes": nbytes, "impl": name, "ms": round(ms.item(), 4),
                         "bus_GBps": round(bus, 1)})
    if ctx.rank == 0:
        for r in rows:
            print(json.dumps(r), flush=True)
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(rows, open(f"gpurun_out/allreduce_sweep_n{world}.json", "w"), indent=1)


def main():
    dp.init()
    rank = dist.get_rank()
    dev = torch.device("cuda", torch.cuda.current_device())
    ok = True
    ctx = SymmContext(dev)
    log(f"symmetric memory provider: {ctx.flags.provider}, multicast ptr: "
        f"{'yes' if ctx.flags.mc_ptr else 'no'}")
    g_ok, ar = check_generic(ctx)
    ok &= g_ok
    log("generic allreduce:", "OK" if g_ok else "FAIL")
    ctx2 = SymmContext(dev)
    f_ok = check_fused_sgd(ctx2, use_mm=False)
    if ctx2.flags.mc_ptr:
        f_ok &= check_fused_sgd(ctx2, use_mm=True)
    ok &= f_ok
    log("fused allreduce+SGD:", "OK" if f_ok else "FAIL")
    t_ok = check_trainer()
    ok &= t_ok
    log("DP trainer:", "OK" if t_ok else "FAIL")
    if os.environ.get("SWEEP", "1") == "1":
        sweep(ar, ctx)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    log("DIST CHECK", "OK" if flag.item() == 1 else "FAILED")
    dist.barrier()
    dp.shutdown()
    return 0 if flag.item() == 1 else 1


if __name__ == "__main__":
    raise SystemExit(main())
