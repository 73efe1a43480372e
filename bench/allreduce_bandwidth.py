"""All-reduce bandwidth sweep (bus GB/s with the 2(n-1)/n factor) — the micro-benchmark of
``pipedream-fork/runtime/tests/communication/all_to_all.py:13-59``, comparing NCCL with the
hand-written symmetric-memory all-reduce (peer loads/stores and NVLS multimem) and the FUSED
all-reduce+SGD kernel.  Reports the achieved fraction of the NVLink roofline (bytes that must
cross the link per GPU / 770 GB/s measured peer-copy bandwidth).

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 bench/allreduce_bandwidth.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ddlbench_b200 import ops  # noqa: E402
from ddlbench_b200.dist.symm import SymmContext  # noqa: E402
from ddlbench_b200.parallel import dp  # noqa: E402
from ddlbench_b200.parallel.collectives import SymmAllreduce  # noqa: E402

LINK_GBPS = 770.0


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / iters], device="cuda")
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return ms.item()


def main():
    dp.init()
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cuda", torch.cuda.current_device())
    ctx = SymmContext(dev)
    ar = SymmAllreduce(ctx, max_bytes=256 << 20, blocks=int(os.environ.get("BLOCKS", "48")))
    ctx2 = SymmContext(dev)
    rows = []
    for nbytes in (8 << 10, 64 << 10, 1 << 20, 8 << 20, 51 << 20, 102 << 20, 256 << 20):
        n = nbytes // 2
        n -= n % (world * 8)
        x = torch.randn(n, device=dev).to(torch.bfloat16)
        res = {"bytes": n * 2, "world": world}
        for name, fn in (("nccl", lambda: dist.all_reduce(x)),
                         ("symm_p2p", lambda: ar.allreduce_(x, True, "p2p")),
                         ("symm_auto", lambda: ar.allreduce_(x, True, "auto"))):
            ms = timed(fn)
            res[name + "_ms"] = round(ms, 4)
            res[name + "_busGBps"] = round(n * 2 * 2 * (world - 1) / world / ms / 1e6, 1)
        # fused all-reduce + SGD on a bucket of the same size (in place on symmetric buffers)
        g, w = ctx2.alloc(n, torch.bfloat16), ctx2.alloc(n, torch.bfloat16)
        sl = n // world
        master = torch.zeros(sl, device=dev)
        mom = torch.zeros(sl, device=dev)
        hyper = torch.tensor([0.1, 0.9, 1e-4, 1.0 / world], device=dev)
        for mm in (False, True):
            if mm and not g.mc_ptr:
                continue
            ms = timed(lambda: ops.ext().ar_sgd_twoshot(
                g.ptrs, w.ptrs, ctx2.flags.ptrs, g.mc_ptr, w.mc_ptr, master, mom, hyper, ctx2.epoch,
                ctx2.status, 0, sl, rank, int(os.environ.get("BLOCKS", "48")), mm))
            key = "fused_sgd_%s" % ("nvls" if mm else "p2p")
            res[key + "_ms"] = round(ms, 4)
            # bytes per direction per GPU: (W-1)/W of the bucket in (grads) and out (weights)
            wire = n * 2 * (world - 1) / world
            res[key + "_roofline_frac"] = round((wire / (LINK_GBPS * 1e6)) / ms, 3)
        ctx2.check()
        rows.append(res)
        if rank == 0:
            print(json.dumps(res), flush=True)
    if rank == 0:
        os.makedirs("gpurun_out", exist_ok=True)
        json.dump(rows, open("gpurun_out/allreduce_bandwidth_n%d.json" % world, "w"), indent=1)
    dist.barrier()
    dp.shutdown()


if __name__ == "__main__":
    main()
